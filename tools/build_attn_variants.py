"""Build A/B variants of the attention kernel (developer tool).  Each variant = attn_sm100.cu compiled with extra -D flags and
linked with the other objects of the in-tree build into build_variants/libvap_<name>.so (git-ignored, travels to the GPU box).
    python tools/build_attn_variants.py name1:-DFOO=1,-DBAR=2 name2: ...
"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "video-as-prompt_b200", "csrc")
OUT = os.path.join(ROOT, "build_variants")
os.makedirs(OUT, exist_ok=True)
sys.path.insert(0, CSRC)
import build as vb  # noqa: E402
vb.build()
others = [os.path.join(CSRC, "build", s.replace(".cu", ".o")) for s in vb.SOURCES if s != "attn_sm100.cu"]

def one(spec):
    name, _, flags = spec.partition(":")
    flags = [f for f in flags.split(",") if f]
    obj = os.path.join(OUT, f"attn_{name}.o")
    r = subprocess.run([vb.NVCC, *vb.FLAGS, *flags, "-c", os.path.join(CSRC, "attn_sm100.cu"), "-o", obj], capture_output=True, text=True)
    if r.returncode:
        return name, r.stderr[-2000:]
    info = [l for l in (r.stdout + r.stderr).splitlines() if "registers" in l or "spill" in l]
    lib = os.path.join(OUT, f"libvap_{name}.so")
    r2 = subprocess.run([vb.NVCC, "-shared", "-o", lib, obj, *others, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lpthread", "-lrt"], capture_output=True, text=True)
    return name, (r2.stderr[-2000:] if r2.returncode else " | ".join(i.strip() for i in info[-4:]))

with ThreadPoolExecutor(8) as ex:
    for name, msg in ex.map(one, sys.argv[1:]):
        print(name, "->", msg)
