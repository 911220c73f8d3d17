"""Import the UNMODIFIED reference implementation at run time.  TEST / BENCH INFRASTRUCTURE ONLY.

Where it comes from (first hit wins): $RFB_REFERENCE, baseline/_ref (the pip --target install written by
tools/install_reference.py: git-ignored, travels to the GPU box with gpurun), /root/reference (the
read-only mount of the build container).  Nothing under renderformer_b200/ or renderformer/ imports this
module (tests/test_abi.py enforces it); users: oracle/make_golden.py, bench.py's reference arms and
cpu_baseline leg, tests/.

The reference hard-imports `roma` (utils/transform.py:3), which this image lacks: a 20-line stand-in with
the five methods it uses (Rigid.from_homogeneous / inverse / __getitem__ / apply / linear_apply,
utils/transform.py:22-27) is put into sys.modules first.  ATTN_IMPL defaults to 'sdpa' (flash-attn cannot
run on CPU; layers/attention.py:18-27 reads it at import time)."""
import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference():
    for p in (os.environ.get("RFB_REFERENCE"), os.path.join(REPO, "baseline", "_ref"), "/root/reference"):
        if p and os.path.isdir(os.path.join(p, "renderformer", "models")):
            return os.path.abspath(p)
    return None


def install_roma_stub():
    if "roma" in sys.modules:
        return

    class Rigid:
        def __init__(self, linear, translation):
            self.linear, self.translation = linear, translation

        @staticmethod
        def from_homogeneous(M):
            return Rigid(M[..., :3, :3], M[..., :3, 3])

        def inverse(self):
            Rt = self.linear.transpose(-1, -2)
            return Rigid(Rt, -(Rt @ self.translation[..., None])[..., 0])

        def __getitem__(self, idx):
            return Rigid(self.linear[idx], self.translation[idx])

        def linear_apply(self, v):
            return (self.linear @ v[..., None])[..., 0]

        def apply(self, v):
            return self.linear_apply(v) + self.translation

    m = types.ModuleType("roma")
    m.Rigid = Rigid
    sys.modules["roma"] = m


def load_reference(attn_impl: str = "sdpa"):
    """(RefConfig, RefModel, RefPipeline, path) of the unmodified reference, or raises ImportError.  The
    repo's own drop-in shim of the same package name is kept off sys.path / sys.modules while importing."""
    ref = find_reference()
    if ref is None:
        raise ImportError("reference not found (baseline/_ref missing: run tools/install_reference.py in the build container)")
    os.environ["ATTN_IMPL"] = attn_impl
    install_roma_stub()
    saved_path = list(sys.path)
    sys.path = [ref] + [p for p in sys.path if os.path.abspath(p or ".") != REPO]
    for k in [k for k in sys.modules if k == "renderformer" or k.startswith("renderformer.")]:
        del sys.modules[k]
    try:
        import renderformer  # noqa: F401  (reference)
        assert os.path.abspath(renderformer.__file__).startswith(ref), renderformer.__file__
        from renderformer.models.config import RenderFormerConfig as RefConfig
        from renderformer.models.renderformer import RenderFormer as RefModel
        from renderformer.pipelines.rendering_pipeline import RenderFormerRenderingPipeline as RefPipe
    finally:
        sys.path = saved_path + ([REPO] if REPO not in saved_path else [])
    return RefConfig, RefModel, RefPipe, ref


def build_reference_pipeline(cfg, state_dict, device="cpu", attn_impl: str = "sdpa"):
    """Reference pipeline object holding OUR seeded weights (strict load: the state_dict keys are the
    contract)."""
    RefConfig, RefModel, RefPipe, ref = load_reference(attn_impl)
    model = RefModel(RefConfig(**cfg.to_dict()))
    model.load_state_dict(state_dict, strict=True)
    model.eval()
    pipe = RefPipe(model)
    if str(device) != "cpu":
        pipe.to(device)
    return pipe, ref
