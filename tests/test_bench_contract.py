"""CPU: the bench contract that can be exercised without a GPU -- `bench.py --impl reference` (the reference arm the
driver launches next to ours) prints ONE JSON line with the agreed keys, times the unmodified reference when
baseline/_ref or /root/reference exists (else the oracle port, and says which), and non-zero ranks of a torchrun
launch exit 0 without work.  Run on a tiny architecture so that it takes seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--impl", "reference", "--config", "tiny_swin", "--tris", "64", "--resolution", "64", "--total-views", "2",
        "--steps", "2", "--warmup", "1"]


def _run(extra_env=None, gpus=1):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", str(gpus)] + ARGS,
                          capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)


def test_reference_arm_prints_one_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in j, k
    assert j["impl"] == "reference" and j["unit"] == "frames/s" and j["higher_is_better"] is True
    assert j["vs_baseline"] is None and j["data"] == "synthetic" and j["dtype"] == "f32" and j["scaling"] == "strong"
    assert j["value"] > 0 and j["ms_per_step"] > 0 and j["gpu_launches"] == 0
    assert "workload" in j["config"] and "model" not in j["config"]
    cb = j["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(cb) and cb["value"] == j["value"]
    have_ref = any(os.path.isdir(os.path.join(p, "renderformer", "models"))
                   for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] == (os.cpu_count() or 1)
    e = j["e2e"]
    assert e["value"] == j["value"] and e["unit"] == j["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, gpus=2)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-500:])
