"""Short-KV attention kernel (one Q tile per CTA, two CTAs per SM) against the two-tile ping-pong kernel on the Wan cross-attention shapes
(text 512 / image 257 tokens, with the accumulate epilogue for the second) and on longer KV sequences — where does the switch-over belong?
    python tools/attn_short_ab.py > gpurun_out/attn_short_ab.json
VAP_ATTN_SHORT = 0 / 1 is read per call.  Outputs of the two kernels are compared (same arithmetic per row up to the accumulation order of P V).
"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vap = importlib.import_module("video-as-prompt_b200")
ops = vap.ops


def timed(fn, iters=10):
    fn()
    torch.cuda.synchronize()
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
    H, D = 40, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: torch.randn(s, generator=g, device="cuda").to(torch.bfloat16)  # noqa: E731
    for Lq in (20280, 2535):
        q = rn(1, Lq, H * D).unflatten(2, (H, D)).transpose(1, 2)
        for Lkv in (257, 512, 1024, 2048, 4096):
            kv = rn(1, Lkv, 2 * H * D)
            k, v = (kv[..., i * H * D:(i + 1) * H * D].unflatten(2, (H, D)).transpose(1, 2) for i in range(2))
            flop = 4.0 * H * Lq * Lkv * D
            res = dict(Lq=Lq, Lkv=Lkv)
            outs = {}
            for mode in ("0", "1"):
                os.environ["VAP_ATTN_SHORT"] = mode
                outs[mode] = ops.attention(q, k, v)
                ms = timed(lambda: ops.attention(q, k, v))
                acc = outs[mode].clone()
                ms_acc = timed(lambda: ops.attention(q, k, v, out=acc, accumulate=True))
                tag = "short" if mode == "1" else "long"
                res[tag + "_us"] = round(ms * 1e3, 1)
                res[tag + "_tflops"] = round(flop / ms / 1e9, 1)
                res[tag + "_accumulate_us"] = round(ms_acc * 1e3, 1)
            os.environ.pop("VAP_ATTN_SHORT")
            ref = outs["0"].float()
            res["short_vs_long_err"] = round(((outs["1"].float() - ref).abs().max() / ref.abs().max()).item(), 5)
            print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
