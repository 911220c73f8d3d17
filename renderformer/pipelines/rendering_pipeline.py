"""renderformer.pipelines.rendering_pipeline -> renderformer_b200.model.RenderFormerRenderingPipeline."""
from renderformer_b200.model import RenderFormerRenderingPipeline

__all__ = ["RenderFormerRenderingPipeline"]
