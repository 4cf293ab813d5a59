"""`compressai` as the reference consumes it at run time (entropy_models.py:42-43, 53-70, 84-92): the entropy-coder
registry plus the two native modules `ans` and `_CXX` -- here bound to libcra5b200.so (see dropin/README.md)."""
_entropy_coder = "ans"
_available_entropy_coders = [_entropy_coder]


def available_entropy_coders():
    return _available_entropy_coders


def get_entropy_coder():
    return _entropy_coder


def set_entropy_coder(entropy_coder):
    global _entropy_coder
    if entropy_coder not in _available_entropy_coders:
        raise ValueError(f'Invalid entropy coder "{entropy_coder}", choose from ({", ".join(_available_entropy_coders)}).')
    _entropy_coder = entropy_coder
