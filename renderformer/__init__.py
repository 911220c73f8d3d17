"""Drop-in `renderformer` package: the reference's import surface (renderformer/__init__.py:1-4)
backed by the B200 kernels in renderformer_b200.  infer.py / batch_infer.py import only this."""
from renderformer.models.renderformer import RenderFormer
from renderformer.pipelines.rendering_pipeline import RenderFormerRenderingPipeline

__all__ = ["RenderFormerRenderingPipeline", "RenderFormer"]
