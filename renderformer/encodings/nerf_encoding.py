"""renderformer.encodings.nerf_encoding (reference: encodings/nerf_encoding.py:25-84)."""
from renderformer_b200.modules import NeRFEncoding

__all__ = ["NeRFEncoding"]
