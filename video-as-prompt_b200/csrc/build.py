"""Build libvap_b200.so in-tree with nvcc for sm_100a (no torch / pybind involvement: the boundary is a plain C ABI).

    python video-as-prompt_b200/csrc/build.py [--force] [--verbose]
    python video-as-prompt_b200/csrc/build.py --debug      # libvap_b200_debug.so: mbarrier watchdog timeouts print before they trap

Objects are compiled in parallel (one nvcc per translation unit) and linked into
``video-as-prompt_b200/libvap_b200.so``.  A content hash of the sources + flags is stored next to the library so
rebuilds are skipped when nothing changed.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent
PKG = CSRC.parent
ROOT = PKG.parent
LIB = PKG / "libvap_b200.so"
BUILD = CSRC / "build"
SOURCES = ["capi.cu", "norm_kernels.cu", "gemm_sm100.cu", "attn_sm100.cu", "attn_bwd_sm100.cu", "probe_sm100.cu"]
HEADERS = ["vap_common.cuh", "vap_kernels.cuh", str(ROOT / "include" / "vap_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _digest() -> str:
    h = hashlib.sha256()
    h.update(" ".join(FLAGS).encode())
    for name in SOURCES + HEADERS:
        p = Path(name) if os.path.isabs(name) else CSRC / name
        h.update(p.read_bytes())
    return h.hexdigest()


LAST_BUILD = "not run"  # "compiled" or "reused (source hash match)": what the last build() call did


def build(force: bool = False, verbose: bool = False) -> Path:
    global LAST_BUILD
    stamp = PKG / ".libvap_b200.hash"
    digest = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text().strip() == digest:
        LAST_BUILD = "reused (source hash match)"
        return LIB
    LAST_BUILD = "compiled"
    BUILD.mkdir(exist_ok=True)

    def compile_one(src: str):
        obj = BUILD / (Path(src).stem + ".o")
        cmd = [NVCC, *FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    log = []
    for src, obj, r in results:
        log.append(f"==== {src}\n{r.stdout}{r.stderr}")
        if r.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    (BUILD / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    objs = [str(o) for _, o, _ in results]
    cmd = [NVCC, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    stamp.write_text(digest)
    return LIB


def build_debug() -> Path:
    """libvap_b200_debug.so: the same sources with -DVAP_MBAR_VERBOSE — a kernel whose mbarrier wait times out (a protocol bug; the product build
    traps silently because a printf in that path costs the attention kernel 5 % of its throughput) prints the barrier, block and thread first.
    Select it with VAP_B200_LIB=<path> when a launch dies with "unspecified launch failure"."""
    out = PKG / "libvap_b200_debug.so"
    dbg = BUILD / "debug"
    dbg.mkdir(parents=True, exist_ok=True)

    def compile_one(src: str):
        obj = dbg / (Path(src).stem + ".o")
        return src, obj, subprocess.run([NVCC, *FLAGS, "-DVAP_MBAR_VERBOSE", "-c", str(CSRC / src), "-o", str(obj)], capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    for src, _, r in results:
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
    r = subprocess.run([NVCC, "-shared", "-o", str(out), *[str(o) for _, o, _ in results], "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl",
                        "-lpthread", "-lrt"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return out


if __name__ == "__main__":
    if "--debug" in sys.argv:
        print(f"compiled: {build_debug()}")
        sys.exit(0)
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(f"{LAST_BUILD}: {lib}")
