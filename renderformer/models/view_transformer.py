"""renderformer.models.view_transformer (reference: models/view_transformer.py:12-127) on the sm_100a kernels."""
from renderformer_b200.modules import ViewTransformer

__all__ = ["ViewTransformer"]
