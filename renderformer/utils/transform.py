"""renderformer.utils.transform (reference: utils/transform.py:7-27) on rfb_positions."""
from renderformer_b200.modules import trans_to_cam_coord

__all__ = ["trans_to_cam_coord"]
