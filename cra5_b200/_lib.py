"""ctypes binding of libcra5b200.so. There is NO fallback: if the library is missing the import fails loudly."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
VARIANT = ""   # build variants (cra5_b200/build.py VARIANTS) -- none at present
LIB_PATH = os.path.join(_HERE, "lib", f"libcra5b200{'_' + VARIANT if VARIANT else ''}.so")


class Cra5Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[cra5_b200 status {code}] {msg}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension has not been built. Run `python -m cra5_b200.build` "
            "(needs nvcc; cross-compiles for sm_100a). cra5_b200 has no CPU fallback.")
    return ctypes.CDLL(LIB_PATH)


lib = _load()
lib.cra5_last_error.restype = ctypes.c_char_p


class Cra5ValueError(ValueError):
    """bad argument, missing CDF tables, or malformed bitstream -- the reference raises ValueError for the first two
    (zoo/image.py:279-290, entropy_models.py:218-256) and has undefined behaviour for the third"""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def check(status):
    if status != 0:
        msg = lib.cra5_last_error().decode("utf-8", "replace")
        if status in (1, 3, 4):
            raise Cra5ValueError(status, msg)
        raise Cra5Error(status, msg)


def ptr(t):
    """device/host pointer of a torch tensor (or None) as c_void_p"""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)
