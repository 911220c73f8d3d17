"""renderformer.utils.ray_generator (reference: utils/ray_generator.py:6-50) on rfb_ray_map."""
from renderformer_b200.modules import RayGenerator

__all__ = ["RayGenerator"]
