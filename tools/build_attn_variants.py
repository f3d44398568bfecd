"""Build A/B variants of one kernel translation unit (developer tool).  Each variant = the unit (default attn_sm100.cu; VAP_VARIANT_UNIT
picks another, e.g. gemm_sm100.cu) compiled with extra -D flags and linked with the other objects of the in-tree build into
build_variants/libvap_<name>.so (git-ignored, travels to the GPU box; select it at run time with VAP_B200_LIB).
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
UNIT = os.environ.get("VAP_VARIANT_UNIT", "attn_sm100.cu")
others = [os.path.join(CSRC, "build", s.replace(".cu", ".o")) for s in vb.SOURCES if s != UNIT]

def one(spec):
    name, _, flags = spec.partition(":")
    flags = [f for f in flags.split(",") if f]
    obj = os.path.join(OUT, f"{UNIT.split('_')[0]}_{name}.o")
    r = subprocess.run([vb.NVCC, *vb.FLAGS, *flags, "-c", os.path.join(CSRC, UNIT), "-o", obj], capture_output=True, text=True)
    if r.returncode:
        return name, r.stderr[-2000:]
    info = [l for l in (r.stdout + r.stderr).splitlines() if "registers" in l or "spill" in l]
    lib = os.path.join(OUT, f"libvap_{name}.so")
    r2 = subprocess.run([vb.NVCC, "-shared", "-o", lib, obj, *others, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lpthread", "-lrt"], capture_output=True, text=True)
    return name, (r2.stderr[-2000:] if r2.returncode else " | ".join(i.strip() for i in info[-4:]))

with ThreadPoolExecutor(8) as ex:
    for name, msg in ex.map(one, sys.argv[1:]):
        print(name, "->", msg)
