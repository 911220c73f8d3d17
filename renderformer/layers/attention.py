"""renderformer.layers.attention (reference: layers/attention.py:34-688) on the sm_100a kernels:
the classes live in renderformer_b200/modules.py."""
from renderformer_b200.modules import (EPS, AttentionLayer, FeedForwardSwiGLU, MultiHeadAttention, SwinSelfAttention,
                                       TransformerDecoder, TransformerEncoder, get_swin_attn_mask, window_partition,
                                       window_reverse)

__all__ = ["EPS", "AttentionLayer", "FeedForwardSwiGLU", "MultiHeadAttention", "SwinSelfAttention", "TransformerDecoder",
           "TransformerEncoder", "get_swin_attn_mask", "window_partition", "window_reverse"]
