"""A/B of attention-kernel builds and run-time modes in ONE process (a fresh GPU box pays a minute for the first `import torch`).

    python tools/build_attn_variants.py rowp0:-DVAP_ATTN_ROW_POLY_D128=0 ...      # here (CPU): build_variants/libvap_<name>.so
    python tools/attn_ab.py [--rounds 2] [--shapes wan,cog,sp8] > gpurun_out/attn_ab.json   # on the GPU

Every library (the in-tree one + build_variants/libvap_*.so) is loaded with ctypes and timed in each run-time mode
(VAP_ATTN_SOFTMAX = lane16 | row, VAP_ATTN_PAIR = 0 | 1, VAP_ATTN_CLUSTER = 0 | 2; all read per call) on the joint-attention shapes of
BASELINE.json configs #2/#3, round-robin over `--rounds` rounds (thermal / power-cap drift hits every variant alike), beside
torch SDPA's cuDNN backend on the same box.  Each variant is checked against torch SDPA on two heads before it is timed.
"""
import argparse
import glob
import importlib
import json
import os
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vap = importlib.import_module("video-as-prompt_b200")
ops = vap.ops
SHAPES = {"wan": (40, 40560, 128), "cog": (48, 35552, 64), "sp8": (5, 40560, 128), "j16k": (40, 16384, 128)}


def use_lib(path):
    vap._lib._lib = None
    vap._lib.LIB_PATH = Path(path)
    vap._lib.load()


def timed(fn, iters=4):
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--shapes", default="wan,cog")
    ap.add_argument("--only", default="", help="comma list of library names (default: all)")
    ap.add_argument("--modes", default="", help="comma list restricting the run-time modes (lane16, pair, row)")
    ap.add_argument("--cl", default="0", help="comma list of VAP_ATTN_CLUSTER values to run (0 = no cluster, 2 = K/V multicast pairs)")
    a = ap.parse_args()
    libs = {"intree": os.path.join(ROOT, "video-as-prompt_b200", "libvap_b200.so")}
    for f in sorted(glob.glob(os.path.join(ROOT, "build_variants", "libvap_*.so"))):
        libs[os.path.basename(f)[7:-3]] = f
    if a.only:
        libs = {k: v for k, v in libs.items() if k in a.only.split(",")}
    variants = []
    for name in libs:
        modes = ["row"] if name.startswith("row") else ["lane16"] if name.startswith("l16") else ["pair"] if name.startswith("pair") else ["lane16", "short", "pair", "row"]
        for m in modes:
            if a.modes and m not in a.modes.split(","):
                continue
            for cl in a.cl.split(","):
                variants.append((name, m, cl))
    res = {}
    data = {}
    for sh in a.shapes.split(","):
        H, J, D = SHAPES[sh]
        g = torch.Generator(device="cuda").manual_seed(0)
        qkv = torch.randn((1, J, 3 * H * D), generator=g, device="cuda").to(torch.bfloat16)
        q, k, v = (qkv[..., i * H * D:(i + 1) * H * D].unflatten(2, (H, D)).transpose(1, 2) for i in range(3))
        ref = F.scaled_dot_product_attention(q[:, :2, :4096], k[:, :2], v[:, :2]).float()
        tail = F.scaled_dot_product_attention(q[:, :1, -300:], k[:, :1], v[:, :1]).float()
        data[sh] = (q, k, v, ref, tail, 4.0 * H * J * J * D)
        from torch.nn.attention import SDPBackend, sdpa_kernel
        qc, kc, vc = q.contiguous(), k.contiguous(), v.contiguous()
        try:
            with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
                F.scaled_dot_product_attention(qc, kc, vc)
                best, avg = timed(lambda: F.scaled_dot_product_attention(qc, kc, vc))
            print(json.dumps(dict(shape=sh, variant="cudnn_sdpa", ms=round(best, 3), tflops=round(data[sh][5] / best / 1e9, 1))), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps(dict(shape=sh, variant="cudnn_sdpa", error=str(e)[:200])), flush=True)
        del qc, kc, vc
    for rnd in range(a.rounds):
        for name, mode, cl in variants:
            use_lib(libs[name])
            os.environ["VAP_ATTN_SOFTMAX"] = "row" if mode == "row" else "lane16"
            os.environ["VAP_ATTN_PAIR"] = "1" if mode == "pair" else "0"  # CTA-pair kernel (cta_group::2), D = 128 only
            os.environ["VAP_ATTN_SHORT"] = "1" if mode == "short" else "0"  # one Q tile per CTA, two CTAs per SM (the short-KV kernel) forced on
            os.environ["VAP_ATTN_CLUSTER"] = cl
            for sh, (q, k, v, ref, tail, flop) in data.items():
                key = f"{sh}/{name}/{mode}/cl{cl}"
                try:
                    if rnd == 0:
                        o, lse = ops.attention(q, k, v, return_lse=True)  # return_lse keeps the unsplit kernel
                        torch.cuda.synchronize()
                        err = ((o[:, :2, :4096].float() - ref).abs().max() / ref.abs().max()).item()
                        err_t = ((o[:, :1, -300:].float() - tail).abs().max() / tail.abs().max()).item()
                        res[key] = dict(err=round(err, 5), err_tail=round(err_t, 5), ms=1e9)
                        if not (err < 1.5e-2 and err_t < 1.5e-2):
                            res[key]["WRONG"] = True
                    fn = lambda: ops.attention(q, k, v, return_lse=True)  # noqa: E731
                    fn()
                    best, avg = timed(fn)
                    r = res[key]
                    r["ms"] = round(min(r["ms"], best), 3)
                    r["tflops"] = round(flop / r["ms"] / 1e9, 1)
                except Exception as e:  # noqa: BLE001 — a trapping variant poisons the context: report and stop
                    print(json.dumps(dict(key=key, error=f"{type(e).__name__}: {str(e)[:300]}")), flush=True)
                    print(json.dumps(res), flush=True)
                    sys.exit(3)
        print(json.dumps({"round": rnd, **{k: v for k, v in res.items()}}), flush=True)
    os.environ.pop("VAP_ATTN_SOFTMAX", None)
    os.environ.pop("VAP_ATTN_PAIR", None)
    os.environ.pop("VAP_ATTN_SHORT", None)
    os.environ.pop("VAP_ATTN_CLUSTER", None)
    best = {}
    for key, r in res.items():
        sh = key.split("/")[0]
        if not r.get("WRONG") and (sh not in best or r["ms"] < best[sh][1]["ms"]):
            best[sh] = (key, r)
    print(json.dumps({"best": {sh: dict(variant=k, **r) for sh, (k, r) in best.items()}}), flush=True)


if __name__ == "__main__":
    main()
