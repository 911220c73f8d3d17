"""CPU: the frame writers of the batch data path (reference batch_infer.py:146-174): EXR / PNG files
written with numpy + zlib only are read back by independent decoders (OpenCV's OpenEXR, Pillow) and the other
way round; the writer pool copies ring-owned frames, applies the CLIs' LDR conversion and keeps the order."""
import os

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")  # must be set before cv2 is imported

import numpy as np
import pytest
import torch

from renderformer_b200 import frame_io as fio


def _hdr(h, w, seed=0, c=3):
    rng = np.random.default_rng(seed)
    img = rng.gamma(0.7, 0.6, size=(h, w, c)).astype(np.float32)
    img[0, 0] = 0.0
    img[-1, -1] = 5000.0  # a light source, raw HDR
    return img


@pytest.mark.parametrize("compression", ["none", "zips", "zip"])
@pytest.mark.parametrize("shape", [(1, 1), (16, 16), (37, 53), (64, 64)])
def test_exr_round_trip_is_exact(tmp_path, compression, shape):
    img = _hdr(*shape, seed=shape[0])
    p = str(tmp_path / "a.exr")
    fio.write_exr(p, img, compression=compression)
    back, names = fio.read_exr(p, return_channels=True)
    assert names == ["R", "G", "B"] and back.dtype == np.float32 and np.array_equal(back, img)


def test_exr_half_and_other_channel_sets(tmp_path):
    img = _hdr(20, 24, seed=3)
    p = str(tmp_path / "h.exr")
    fio.write_exr(p, np.minimum(img, 60000.0), half=True)
    assert np.array_equal(fio.read_exr(p), np.minimum(img, 60000.0).astype(np.float16).astype(np.float32))
    rgba = _hdr(9, 11, seed=4, c=4)
    fio.write_exr(p, rgba)
    back, names = fio.read_exr(p, return_channels=True)
    assert names == ["R", "G", "B", "A"] and np.array_equal(back, rgba)
    grey = _hdr(9, 11, seed=5, c=1)
    fio.write_exr(p, grey[:, :, 0])
    back, names = fio.read_exr(p, return_channels=True)
    assert names == ["Y"] and np.array_equal(back, grey)
    with pytest.raises(ValueError):
        fio.write_exr(p, img, compression="piz")
    with pytest.raises(ValueError):
        fio.write_exr(p, np.zeros((4, 4, 2), np.float32))


def test_exr_incompressible_block_is_stored_raw(tmp_path):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 2 ** 32, size=(16, 16, 3), dtype=np.uint64).astype(np.uint32).view(np.float32)
    img = np.nan_to_num(img, nan=1.0, posinf=2.0, neginf=-2.0)  # random bit patterns do not deflate
    p = str(tmp_path / "r.exr")
    fio.write_exr(p, img, compression="zip")
    assert np.array_equal(fio.read_exr(p), img)


def test_exr_cross_checked_with_opencv(tmp_path):
    cv2 = pytest.importorskip("cv2")
    if "OpenEXR" not in cv2.getBuildInformation():
        pytest.skip("OpenCV built without OpenEXR")
    img = _hdr(48, 40, seed=7)
    p = str(tmp_path / "ours.exr")
    for comp in ("none", "zips", "zip"):
        fio.write_exr(p, img, compression=comp)
        theirs = cv2.imread(p, cv2.IMREAD_UNCHANGED)  # BGR
        assert theirs is not None and theirs.dtype == np.float32 and np.array_equal(theirs[:, :, ::-1], img), comp
    fio.write_exr(p, np.minimum(img, 60000.0), half=True)
    theirs = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    assert np.array_equal(theirs[:, :, ::-1].astype(np.float32), np.minimum(img, 60000.0).astype(np.float16).astype(np.float32))
    # the other direction: files OpenCV writes (float and half, ZIP / ZIPS / uncompressed)
    q = str(tmp_path / "theirs.exr")
    for flags in ([], [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF],
                  [cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_NO],
                  [cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_ZIPS]):
        small = np.minimum(img, 60000.0)
        assert cv2.imwrite(q, np.ascontiguousarray(small[:, :, ::-1]), flags)
        want = cv2.imread(q, cv2.IMREAD_UNCHANGED)[:, :, ::-1].astype(np.float32)
        assert np.array_equal(fio.read_exr(q), want), flags


@pytest.mark.parametrize("c", [1, 3, 4])
def test_png_round_trip_and_pillow(tmp_path, c):
    rng = np.random.default_rng(c)
    img = rng.integers(0, 256, size=(33, 47, c), dtype=np.uint8)
    p = str(tmp_path / "a.png")
    fio.write_png(p, img)
    assert np.array_equal(fio.read_png(p), img)
    Image = pytest.importorskip("PIL.Image")
    with Image.open(p) as im:
        theirs = np.asarray(im)
    assert np.array_equal(theirs.reshape(img.shape), img)
    # Pillow's encoder chooses scanline filters adaptively: the reader has to undo all of them
    smooth = (np.add.outer(np.arange(33), np.arange(47))[:, :, None] * np.arange(1, c + 1)).astype(np.uint8)
    Image.fromarray(smooth[:, :, 0] if c == 1 else smooth).save(p, optimize=True)
    assert np.array_equal(fio.read_png(p), smooth)
    with pytest.raises(ValueError):
        fio.write_png(p, img.astype(np.float32))


def test_ldr_conversion_matches_the_reference_clis():
    hdr = _hdr(32, 32, seed=9) - 0.1  # negatives clip to 0
    assert np.array_equal(fio.hdr_to_ldr_host(hdr), (np.clip(hdr, 0, 1) * 255).astype(np.uint8))  # batch_infer.py:153-157
    ramp = np.repeat(np.linspace(0, 8, 400, dtype=np.float32)[:, None], 3, axis=1)
    pbr = fio.hdr_to_ldr_host(ramp, "pbr_neutral").astype(int)
    assert (np.diff(pbr[:, 0]) >= 0).all() and pbr[0, 0] == 0 and 250 <= pbr[-1, 0] <= 255
    # a saturated colour above the knee is compressed and desaturated towards white, never above 1
    sat = fio.hdr_to_ldr_host(np.array([[4.0, 0.2, 0.1]], np.float32), "Khronos PBR Neutral")[0].astype(int)
    assert sat[0] >= sat[1] >= sat[2] and sat[1] > fio.hdr_to_ldr_host(np.array([[0.9, 0.2, 0.1]], np.float32), "pbr_neutral")[0, 1] - 80
    with pytest.raises(ValueError, match="AgX"):
        fio.hdr_to_ldr_host(ramp, "agx")


def test_frame_writer_copies_ring_buffers_and_keeps_order(tmp_path):
    ring = torch.zeros(8, 8, 3)  # a buffer the producer overwrites right after submit, like render_stream's ring
    frames = []
    with fio.FrameWriter(str(tmp_path), workers=3, max_pending=2, keep_ldr=True) as fw:
        for i in range(9):
            ring.fill_(i / 8.0)
            frames.append(ring.numpy().copy())
            fw.submit(f"f{i}", ring)
        ldr = fw.close()
    assert len(ldr) == 9
    for i, fr in enumerate(frames):
        assert np.array_equal(fio.read_exr(str(tmp_path / f"f{i}.exr")), fr)
        want = (np.clip(fr, 0, 1) * 255).astype(np.uint8)
        assert np.array_equal(fio.read_png(str(tmp_path / f"f{i}.png")), want) and np.array_equal(ldr[i], want)
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".part")]


def test_frame_writer_surfaces_worker_errors(tmp_path):
    fw = fio.FrameWriter(str(tmp_path), workers=1)
    os.chmod(tmp_path, 0o500)
    try:
        if os.access(str(tmp_path), os.W_OK):  # running as root: permissions do not bite, provoke it differently
            fw.out_dir = str(tmp_path / "missing" / "dir")
        fw.submit("x", np.zeros((4, 4, 3), np.float32))
        with pytest.raises(OSError):
            fw.close()
    finally:
        os.chmod(tmp_path, 0o700)
    with pytest.raises(RuntimeError):
        fw.submit("y", np.zeros((4, 4, 3), np.float32))
    with pytest.raises(ValueError):
        fio.FrameWriter(str(tmp_path), tone_mapper="filmic")


class _FakePipeline:
    """Stands in for RenderFormerRenderingPipeline.render_stream: yields a ring of reused host buffers."""

    def __init__(self):
        self.calls = []

    def render_stream(self, scenes, resolution=512, torch_dtype=None, ldr=None, pad_to=None):
        self.calls.append(dict(resolution=resolution, torch_dtype=torch_dtype, pad_to=pad_to))
        ring = [None, None, None]
        for i, sc in enumerate(scenes):
            B, V = sc["c2w"].shape[:2]
            if ring[i % 3] is None or ring[i % 3].shape[:2] != (B, V):
                ring[i % 3] = torch.empty(B, V, resolution, resolution, 3)
            buf = ring[i % 3]
            for b in range(B):
                for v in range(V):
                    buf[b, v] = float(sc["tag"]) + b + 0.25 * v
            yield buf


def test_render_to_files_names_frames_like_the_reference(tmp_path):
    scenes = [dict(tag=0.0, c2w=torch.zeros(1, 2, 4, 4)), dict(tag=0.5, c2w=torch.zeros(2, 1, 4, 4)),
              dict(tag=0.125, c2w=torch.zeros(1, 2, 4, 4)), dict(tag=0.75, c2w=torch.zeros(1, 2, 4, 4))]
    names = ["a", "b0", "b1", "c", "d"]
    pipe = _FakePipeline()
    try:
        import cv2  # noqa: F401
        video = True
    except ImportError:
        video = False
    paths = fio.render_to_files(pipe, iter(scenes), names, str(tmp_path), resolution=16, torch_dtype=torch.bfloat16,
                                pad_to=64, save_video=video)
    assert pipe.calls == [dict(resolution=16, torch_dtype=torch.bfloat16, pad_to=64)]
    want = {"a_view_0": 0.0, "a_view_1": 0.25, "b0_view_0": 0.5, "b1_view_0": 1.5, "c_view_0": 0.125, "c_view_1": 0.375,
            "d_view_0": 0.75, "d_view_1": 1.0}
    assert [os.path.basename(p) for p in paths] == list(want)
    for base, val in want.items():
        hdr = fio.read_exr(str(tmp_path / (base + ".exr")))
        assert hdr.shape == (16, 16, 3) and (hdr == np.float32(val)).all()
        assert (fio.read_png(str(tmp_path / (base + ".png"))) == int(min(val, 1.0) * 255)).all()
    if video:
        import cv2
        cap = cv2.VideoCapture(str(tmp_path / "video.mp4"))
        n = 0
        while cap.read()[0]:
            n += 1
        assert n == len(want)
    with pytest.raises(ValueError, match="names"):
        fio.render_to_files(_FakePipeline(), iter(scenes), names[:2], str(tmp_path / "o2"), resolution=16)


def test_render_folder_tool_orders_scenes_like_natsort(tmp_path):
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("render_folder", os.path.join(root, "tools", "render_folder.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    names = ["frame_10.npz", "frame_2.npz", "Frame_1.npz", "frame_02b.npz"]
    assert sorted(names, key=tool.natural_key) == ["Frame_1.npz", "frame_2.npz", "frame_02b.npz", "frame_10.npz"]
    assert tool.main(["--scene_folder", str(tmp_path), "--random_init", "tiny_swin"]) == 1  # empty folder: no GPU touched


def _sharded_worker(rank, world, port, out_dir, q):
    import torch.distributed as dist
    import renderformer_b200.dist as rdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def fake_stream(pipe, scenes, resolution=512, ldr=None, torch_dtype=None):
            for sc in scenes:  # what dist.render_stream_sharded yields: (this rank's view slice, its frames)
                V = sc["c2w"].shape[1]
                mine = rdist.view_slice(V, world, rank)
                img = torch.empty(1, mine.stop - mine.start, resolution, resolution, 3)
                for i, v in enumerate(range(mine.start, mine.stop)):
                    img[0, i] = float(sc["tag"]) + 0.125 * v
                yield mine, img
        rdist.render_stream_sharded = fake_stream
        scenes = [dict(tag=0.0, c2w=torch.zeros(1, 3, 4, 4)), dict(tag=0.5, c2w=torch.zeros(1, 3, 4, 4))]
        paths = fio.render_to_files(object(), iter(scenes), ["s0", "s1"], out_dir, resolution=16, save_video=True, sharded=True)
        q.put((rank, [os.path.basename(p) for p in paths]))
    finally:
        dist.destroy_process_group()


def test_render_to_files_sharded_world2_gloo(tmp_path):
    """Every rank writes its own views under the GLOBAL view index; rank 0 assembles the video from the PNG files
    after the barrier (host logic of the multi-GPU batch path, two gloo ranks on CPU)."""
    import socket
    import torch.multiprocessing as mp
    pytest.importorskip("cv2")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == ["s0_view_0", "s0_view_1", "s1_view_0", "s1_view_1"] and got[1] == ["s0_view_2", "s1_view_2"]
    for name, tag in (("s0", 0.0), ("s1", 0.5)):
        for v in range(3):
            assert (fio.read_exr(str(tmp_path / f"{name}_view_{v}.exr")) == np.float32(tag + 0.125 * v)).all()
    import cv2
    cap = cv2.VideoCapture(str(tmp_path / "video.mp4"))
    n = 0
    while cap.read()[0]:
        n += 1
    assert n == 6


class _EchoPipeline:
    """Stand-in with the pipeline attributes the tool touches; "renders" the mean diffuse colour of the valid triangles so
    that the files prove which scene / padding / texture form reached render_stream."""
    cuda_graphs = False

    def __init__(self):
        self.seen = []

    def render_stream(self, scenes, resolution=512, torch_dtype=None, ldr=None, pad_to=None):
        for sc in scenes:
            self.seen.append({k: tuple(v.shape) for k, v in sc.items()} | {"pad_to": pad_to, "dtype": torch_dtype})
            B, V = sc["c2w"].shape[:2]
            tex = sc["texture"]
            diffuse = tex[0, :, :3] if tex.dim() == 3 else tex[0, :, :3, 0, 0]
            colour = diffuse[sc["mask"][0]].mean(dim=0)
            yield colour.view(1, 1, 1, 1, 3).expand(B, V, resolution, resolution, 3).contiguous()


@pytest.mark.parametrize("flags", [[], ["--constant_texture"], ["--padding_length", "32"]],
                         ids=["texel-grid", "constant-texture", "padded"])
def test_render_folder_tool_host_logic(tmp_path, flags):
    """tools/render_folder.py end to end on the CPU with a stand-in pipeline: natural file order, scene loading,
    texture form, padding flag, precision, file names and contents."""
    import importlib.util
    from renderformer_b200 import scene_io as sio
    from renderformer_b200.synth import make_scene
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("render_folder", os.path.join(root, "tools", "render_folder.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    folder = tmp_path / "scenes"
    folder.mkdir()
    want = {}
    for name, n in (("frame_10", 9), ("frame_2", 14)):
        sc = make_scene(n, 2, seed=n)
        tex13 = sc["texture"][0, :, :, 0, 0].numpy()
        sio.save_npz({"triangles": sc["triangles"][0].numpy(), "vn": sc["vn"][0].numpy(), "tex13": tex13,
                      "c2w": sc["c2w"][0].numpy(), "fov": sc["fov"][0, :, 0].numpy()}, str(folder / f"{name}.npz"))
        want[name] = tex13[:, :3].mean(axis=0)
    pipe = _EchoPipeline()
    out = tmp_path / "out"
    rc = tool.main(["--scene_folder", str(folder), "--random_init", "tiny_swin", "--precision", "bf16", "--resolution", "16",
                    "--output_dir", str(out)] + flags, _pipeline=pipe)
    assert rc == 0
    assert [s["triangles"][1] for s in pipe.seen] == [14, 9]                       # frame_2 before frame_10 (natural order)
    assert all(s["dtype"] == torch.bfloat16 for s in pipe.seen)
    assert all(s["pad_to"] == (32 if "--padding_length" in flags else None) for s in pipe.seen)
    assert all(len(s["texture"]) == (3 if "--constant_texture" in flags else 5) for s in pipe.seen)
    assert pipe.cuda_graphs == ("--padding_length" in flags)
    for name, colour in want.items():
        for v in range(2):
            hdr = fio.read_exr(str(out / f"{name}_view_{v}.exr"))
            assert hdr.shape == (16, 16, 3) and np.allclose(hdr[3, 5], colour, atol=1e-6)
            assert np.array_equal(fio.read_png(str(out / f"{name}_view_{v}.png")), (np.clip(hdr, 0, 1) * 255).astype(np.uint8))


def test_prefetch_keeps_order_overlaps_and_propagates_errors():
    import threading
    import time
    made = []

    def slow_source(n, fail_at=None):
        for i in range(n):
            time.sleep(0.02)
            if i == fail_at:
                raise KeyError(f"scene {i}")
            made.append((i, threading.current_thread().name))
            yield i

    got, lead = [], []
    for x in fio.prefetch(slow_source(10), depth=2):
        time.sleep(0.03)  # the consumer's own work overlaps the producer's
        got.append(x)
        lead.append(len(made) - len(got))  # items the producer has finished beyond the ones consumed
    assert got == list(range(10)) and all(name == "rfb-prefetch" for _, name in made)
    assert max(lead) >= 1  # the producer ran ahead while the consumer was busy
    with pytest.raises(KeyError, match="scene 3"):
        out = []
        for x in fio.prefetch(slow_source(10, fail_at=3)):
            out.append(x)
    assert out == [0, 1, 2]
    # closing early stops the producer thread
    made.clear()
    g = fio.prefetch(slow_source(1000), depth=1)
    assert next(g) == 0
    g.close()
    n = len(made)
    time.sleep(0.1)
    assert len(made) <= n + 1 and not [t for t in threading.enumerate() if t.name == "rfb-prefetch" and t.is_alive()]
