#!/usr/bin/env python
"""Install the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with gpurun).

    python tools/install_reference.py [/root/reference]

The reference ships no setup.py / pyproject.toml, so `pip install /root/reference` has nothing to build.
As the task statement allows for builds that need to write next to the sources, the tree is copied to a
scratch directory under /tmp, a five-line setup.py naming its package (`renderformer`) and its two CLI
modules (`infer`, `batch_infer`) is written THERE, and pip installs that copy with
`--no-index --no-build-isolation --no-deps --target baseline/_ref`.  No reference file is edited and none
enters the repository history (baseline/_ref is in .gitignore).  Used by: bench.py --impl reference and
the reference-CUDA context arm (oracle/reference_loader.py), tests/test_reference_clis_gpu.py.
Third-party imports the reference needs and this image lacks (roma; h5py / imageio / simple_ocio /
natsort for the CLIs) are stubbed at import time by oracle/reference_loader.py."""
import os
import shutil
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(REPO, "baseline", "_ref")

SETUP = '''from setuptools import setup, find_packages
setup(name="renderformer-reference", version="0.0.0", packages=find_packages(include=["renderformer", "renderformer.*"]),
      py_modules=["infer", "batch_infer"])
'''


def main(src: str = "/root/reference") -> int:
    if not os.path.isdir(os.path.join(src, "renderformer")):
        print(f"install_reference: {src} has no renderformer package; nothing installed")
        return 1
    tmp = tempfile.mkdtemp(prefix="rf_ref_")
    try:
        shutil.copytree(os.path.join(src, "renderformer"), os.path.join(tmp, "renderformer"))
        for f in ("infer.py", "batch_infer.py"):
            shutil.copy(os.path.join(src, f), os.path.join(tmp, f))
        with open(os.path.join(tmp, "setup.py"), "w") as f:
            f.write(SETUP)
        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        os.makedirs(DEST, exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", DEST, tmp]
        rc = subprocess.run(cmd).returncode
        if rc != 0:
            print("install_reference: pip failed with", rc)
            return rc
        # unmodified? compare every installed .py with its source byte for byte
        bad = []
        for root, _, files in os.walk(os.path.join(tmp, "renderformer")):
            for fn in files:
                if fn.endswith(".py"):
                    a = os.path.join(root, fn)
                    b = os.path.join(DEST, os.path.relpath(a, tmp))
                    if not os.path.exists(b) or open(a, "rb").read() != open(b, "rb").read():
                        bad.append(os.path.relpath(a, tmp))
        if bad:
            print("install_reference: installed files differ from the source:", bad)
            return 2
        print(f"install_reference: {src} -> {DEST} (unmodified, {sum(len(f) for _, _, f in os.walk(DEST))} files)")
        return 0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main(*sys.argv[1:2]))
