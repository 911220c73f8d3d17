"""renderformer.layers.dpt (reference: layers/dpt.py:57-273) on the sm_100a kernels."""
from renderformer_b200.modules import DPTHead, FeatureFusionBlock, ResidualConvUnit

__all__ = ["DPTHead", "FeatureFusionBlock", "ResidualConvUnit"]
