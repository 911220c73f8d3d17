"""renderformer.encodings.rope (reference: encodings/rope.py:41-206)."""
from renderformer_b200.modules import (TriangleRotaryEmbedding, apply_rotary_emb_cossin, apply_rotary_emb_one_cossin,
                                       freqs_to_cos_sin, rotate_half_hf)

__all__ = ["TriangleRotaryEmbedding", "apply_rotary_emb_cossin", "apply_rotary_emb_one_cossin", "freqs_to_cos_sin",
           "rotate_half_hf"]
