"""oracle shim (TEST INFRASTRUCTURE ONLY): the reference's RGB loss module imports
`ms_ssim`; never called on the VAEformer inference path."""


def ms_ssim(*a, **k):
    raise NotImplementedError("pytorch_msssim is not available (oracle shim)")
