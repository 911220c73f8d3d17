"""CPU: the C-ABI library loads without a GPU and exports every symbol include/rfb200.h declares;
the ctypes structs mirror the C layouts; calls without a device fail loudly (no fallback)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from renderformer_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def handle():
    if not os.path.exists(lib.LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "renderformer_b200", "csrc"), "-j8"], check=True)
    return lib.load()


def test_header_symbols_exported(handle):
    with open(os.path.join(ROOT, "include", "rfb200.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(rfb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)
    for name in declared:
        assert hasattr(handle, name), f"{name} not exported by librfb200.so"
    assert handle.rfb_version() >= 100


def test_struct_layouts_match_c(handle, tmp_path):
    """sizeof() of the two argument structs as seen by the C compiler vs ctypes."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "rfb200.h"\nint main(){printf("%zu %zu\\n", sizeof(rfb_gemm_args), sizeof(rfb_attn_args));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    a, b = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert a == ctypes.sizeof(lib.GemmArgs)
    assert b == ctypes.sizeof(lib.AttnArgs)


def test_bad_arguments_are_rejected_without_gpu(handle):
    assert handle.rfb_gemm(None, None) == -1
    assert handle.rfb_attention(None, None) == -1
    assert handle.rfb_rmsnorm(None, 0, None, None, 0, 0, 0, 0, 0.0, None, None) == -1


def test_product_path_has_no_cpu_fallback():
    from renderformer_b200.config import RenderFormerConfig
    from renderformer_b200.model import RenderFormer, RenderFormerRenderingPipeline
    from renderformer_b200.synth import make_scene
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    pipe = RenderFormerRenderingPipeline(RenderFormer(RenderFormerConfig.named("tiny_swin")))
    sc = make_scene(8, 1)
    with pytest.raises(lib.RfbError):
        pipe(sc["triangles"], sc["texture"], sc["mask"], sc["vn"], sc["c2w"], sc["fov"], resolution=64)
    with pytest.raises(lib.RfbError):
        list(pipe.render_stream(iter([sc]), resolution=64))


def test_product_does_not_import_oracle():
    """Only tests/ (incl. tests/diagnostics), __graft_entry__.smoke() and bench.py's CPU arm may touch oracle/;
    tools/ are user-facing and stay oracle-free as well (install_reference.py only names the loader in its docstring)."""
    for base in ("renderformer_b200", "renderformer", "renderformer_liger_kernel", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith(".py"):
                    with open(os.path.join(dirpath, fn)) as f:
                        src = f.read()
                    assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, fn)


def test_integration_doc_binding_matches_the_library_struct():
    """INTEGRATION.md shows the ctypes struct a maintainer would write on the reference side: it has to stay the
    field-for-field mirror of rfb_gemm_args (same names, same order, same widths) as the header evolves."""
    import ctypes
    import re
    from renderformer_b200 import lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    block = doc[doc.index("class GemmArgs"):doc.index("def linear_bf16")]
    shown = re.findall(r'\("(\w+)", ctypes\.(\w+)\)', block)
    real = [(n, t) for n, t in lib.GemmArgs._fields_]
    assert [n for n, _ in shown] == [n for n, _ in real]
    for (n, tname), (_, t) in zip(shown, real):
        assert ctypes.sizeof(getattr(ctypes, tname)) == ctypes.sizeof(t), n
