"""A/B of GEMM builds per EPILOGUE (bias / GELU / gated residual / residual add) on the MoT projection shapes, one process, round-robin.

    python tools/gemm_epi_ab.py [--rounds 3] > gpurun_out/gemm_epi_ab.json

Libraries: the in-tree one + build_variants/libvap_*.so.  Shapes: the Wan-14B O-projection / FFN-down at the full 480p token count (one GPU)
and at the 2 535 rows one rank owns under 8-way Ulysses.  Every variant's output is compared with the in-tree library's (must be bit-identical:
an epilogue change may reorder loads, never arithmetic) and with torch's bf16 matmul for the bias-only case.
"""
import argparse
import glob
import importlib
import json
import os
import sys
from pathlib import Path

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vap = importlib.import_module("video-as-prompt_b200")
ops = vap.ops
EPI = {"bias": ops.EPI_BIAS, "gelu": ops.EPI_BIAS_GELU, "gate_res_f32": ops.EPI_GATE_RES_F32, "res_add": ops.EPI_RES_ADD, "gate_res_bf16": ops.EPI_GATE_RES_BF16}
SHAPES = [(20280, 5120, 5120), (20280, 5120, 13824), (2535, 5120, 5120), (2535, 5120, 13824), (17776, 3072, 3072)]


def use_lib(path):
    vap._lib._lib = None
    vap._lib.LIB_PATH = Path(path)
    vap._lib.load()


def timed(fn, iters=6):
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=3)
    a = ap.parse_args()
    libs = {"intree": os.path.join(ROOT, "video-as-prompt_b200", "libvap_b200.so")}
    for f in sorted(glob.glob(os.path.join(ROOT, "build_variants", "libvap_*.so"))):
        libs[os.path.basename(f)[7:-3]] = f
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: torch.randn(s, generator=g, device="cuda").to(torch.bfloat16)  # noqa: E731
    res = {}
    for M, N, K in SHAPES:
        x, w, b, r = rn(M, K), rn(N, K) / K ** 0.5, rn(N), rn(M, N)
        gate = torch.randn((1, N), generator=g, device="cuda")
        flop = 2.0 * M * N * K
        want = {}
        for rnd in range(a.rounds):
            for name, path in libs.items():
                use_lib(path)
                for ename, e in EPI.items():
                    kw = dict(epilogue=e)
                    if e >= ops.EPI_GATE_RES_F32:
                        kw["residual"] = r
                    if e in (ops.EPI_GATE_RES_F32, ops.EPI_GATE_RES_BF16):
                        kw["gate"] = gate
                    key = f"{M}x{N}x{K}/{ename}/{name}"
                    out = ops.linear(x, w, b, **kw)
                    if rnd == 0:
                        if name == "intree":
                            want[ename] = out.clone()
                            if e == ops.EPI_BIAS:
                                ref = torch.nn.functional.linear(x, w, b).float()
                                res[key + "/err_vs_torch"] = round(((out.float() - ref).abs().max() / ref.abs().max()).item(), 5)
                        else:
                            res[key + "/bit_identical_to_intree"] = bool(torch.equal(out, want[ename]))
                    ms = timed(lambda: ops.linear(x, w, b, **kw))
                    ent = res.setdefault(key, dict(ms=1e9))
                    ent["ms"] = round(min(ent["ms"], ms), 4)
                    ent["tflops"] = round(flop / ent["ms"] / 1e9, 1)
        print(json.dumps({k: v for k, v in res.items() if k.startswith(f"{M}x{N}x{K}/")}), flush=True)
        del x, w, b, r, want


if __name__ == "__main__":
    main()
