"""cra5/models/vaeformer: `VAEformer` (vaeformer.py:70)"""
from cra5_b200.vaeformer import VAEformer

__all__ = ["VAEformer"]
