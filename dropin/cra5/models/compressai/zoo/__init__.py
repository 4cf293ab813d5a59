"""cra5/models/compressai/zoo: only the VAEformer entry of the reference's model zoo is on the hot path
(zoo/image.py:302-324)."""
from cra5_b200.zoo import _load_model, cfgs, model_urls, vaeformer_pretrained

__all__ = ["vaeformer_pretrained", "cfgs", "model_urls"]
