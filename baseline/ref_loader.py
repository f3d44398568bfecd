"""Stage and import the UNMODIFIED reference (bytedance/Video-As-Prompt's vendored diffusers) for the GPU-side comparisons.

    python baseline/ref_loader.py            # stage: pip-install /root/reference/diffusers into baseline/_ref (git-ignored)

`/root/reference` exists only in the authoring container; `baseline/_ref/` is git-ignored but NOT gpurun-ignored, so the
installed copy travels to the GPU box with the snapshot (like the built libvap_b200.so).  Nothing of the reference is
committed.  `load()` puts the staged tree (or, failing that, /root/reference/diffusers/src) on sys.path and returns the
`diffusers` module; `available()` says whether either exists.  Used ONLY by tests/ (drop-in + parity against the
reference's own classes on the GPU), bench.py's `reference_gpu` / `--impl reference` legs and tools/ — the product
package never imports it.
"""
from __future__ import annotations

import importlib
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("VAP_REFERENCE", "/root/reference")


def staged() -> bool:
    return os.path.isfile(os.path.join(STAGED, "diffusers", "__init__.py"))


def available() -> bool:
    return staged() or os.path.isdir(os.path.join(SOURCE, "diffusers", "src", "diffusers"))


def stage(force: bool = False) -> str:
    """pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference/diffusers>.
    (The source tree is read-only and setuptools writes an egg-info next to setup.py, hence the /tmp copy; dependency
    resolution is skipped: torch / transformers / safetensors / huggingface_hub of the image are used as they are.)"""
    if staged() and not force:
        return "reused"
    src = os.path.join(SOURCE, "diffusers")
    if not os.path.isdir(src):
        return "unavailable (no reference tree at %s)" % SOURCE
    tmp = tempfile.mkdtemp(prefix="vap_ref_")
    try:
        shutil.copytree(src, os.path.join(tmp, "diffusers"), ignore=shutil.ignore_patterns("docs", "tests", "examples", "benchmarks", "docker", ".git"))
        if os.path.isdir(STAGED):
            shutil.rmtree(STAGED)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
               "--target", STAGED, os.path.join(tmp, "diffusers")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not staged():
            return "failed: " + (r.stdout + r.stderr)[-400:]
        return "installed"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def load():
    """-> the reference's `diffusers` module (stock, unpatched), or raises ImportError."""
    if "diffusers" in sys.modules:
        return sys.modules["diffusers"]
    if staged():
        path = STAGED
    elif os.path.isdir(os.path.join(SOURCE, "diffusers", "src", "diffusers")):
        path = os.path.join(SOURCE, "diffusers", "src")
    else:
        raise ImportError("the reference is neither staged under baseline/_ref nor present at " + SOURCE)
    sys.dont_write_bytecode = True  # never write __pycache__ into a read-only reference tree
    if path not in sys.path:
        sys.path.insert(0, path)
    return importlib.import_module("diffusers")


if __name__ == "__main__":
    print("baseline/_ref:", stage(force="--force" in sys.argv))
