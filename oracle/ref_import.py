"""Import the REAL reference (read-only mount) through the oracle shims.

TEST INFRASTRUCTURE ONLY -- used by tools/make_golden.py and by tests that pin the
oracle restatement against the reference when /root/reference is present (i.e. in the
build container; never on the GPU box, where the mount does not exist).

Recipe: shims on sys.path, compiled `compressai.ans/_CXX` from
oracle/_ref, import `cra5.models.compressai.zoo` FIRST (circular import otherwise,
zoo/image.py:41), force ATTENTION_MODE='math' (flash_attn is importable here but needs CUDA).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CRA5_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "cra5")) and os.path.isdir(os.path.join(HERE, "_ref", "compressai"))


_cached = None


def load():
    """Returns a namespace with the reference's classes/functions."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference mount or oracle/_ref missing; run oracle/build_ref.py in the build container")
    for p in (os.path.join(HERE, "shims"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import types
    from cra5.models.compressai.zoo import vaeformer_pretrained  # noqa: must be first
    from cra5.models.vaeformer import vit_nlc
    from cra5.models.vaeformer.vaeformer import VAEformer
    from cra5.models.compressai.entropy_models import EntropyBottleneck, GaussianConditional
    from cra5.models.compressai.entropy_models.entropy_models import pmf_to_quantized_cdf
    from compressai import ans, _CXX

    vit_nlc.ATTENTION_MODE = "math"
    ns = types.SimpleNamespace(
        vaeformer_pretrained=vaeformer_pretrained, VAEformer=VAEformer, vit_nlc=vit_nlc,
        EntropyBottleneck=EntropyBottleneck, GaussianConditional=GaussianConditional,
        pmf_to_quantized_cdf=pmf_to_quantized_cdf, ans=ans, _CXX=_CXX)
    _cached = ns
    return ns
