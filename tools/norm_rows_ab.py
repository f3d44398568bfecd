"""adaLN-LayerNorm and q/k-norm + RoPE at the row counts one rank owns under 2 / 4 / 8-way Ulysses (and the full stream): staged kernels with
1 / 2 / 4 (default) / 8 minimum rows per warp, and the register kernels (VAP_NORM_STAGED=0), beside a device copy of the same bytes.
    python tools/norm_rows_ab.py > gpurun_out/norm_rows_ab.json
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


def timed(fn, iters=20):
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
    d, H, D = 5120, 40, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for rows in (2535, 5070, 10140, 20280):
        x = torch.randn((rows, d), generator=g, device="cuda").to(torch.bfloat16)
        qkv = torch.randn((rows, 3 * d), generator=g, device="cuda").to(torch.bfloat16)
        s1p, sh = torch.randn((1, d), device="cuda"), torch.randn((1, d), device="cuda")
        w = torch.ones(d, device="cuda")
        cos, sin = torch.rand((rows, D // 2), device="cuda"), torch.rand((rows, D // 2), device="cuda")
        y = torch.empty_like(x)
        res = {"rows": rows, "copy_us": round(timed(lambda: y.copy_(x)) * 1e3, 1)}
        ln = lambda: ops.adaln_layernorm(x, eps=1e-6, rounding=0, scale1p=s1p, shift=sh)  # noqa: E731
        qk = lambda: ops.qk_norm_rope_(qkv[:, :d], qkv[:, d:2 * d], heads=H, head_dim=D, wq=w, wk=w, cos=cos, sin=sin, rows_per_batch=rows, eps=1e-6, mode=0)  # noqa: E731
        for tag, env in (("staged_rpw1", {"VAP_NORM_STAGED_ROWS_PER_WARP": "1"}), ("staged_rpw2", {"VAP_NORM_STAGED_ROWS_PER_WARP": "2"}),
                         ("staged_rpw4", {"VAP_NORM_STAGED_ROWS_PER_WARP": "4"}), ("staged_rpw8", {"VAP_NORM_STAGED_ROWS_PER_WARP": "8"}),
                         ("register", {"VAP_NORM_STAGED": "0"})):
            os.environ.update(env)
            res[f"ln_{tag}_us"] = round(timed(ln) * 1e3, 1)
            res[f"qk_{tag}_us"] = round(timed(qk) * 1e3, 1)
            for k in env:
                os.environ.pop(k)
        print(json.dumps(res), flush=True)
    del flush


if __name__ == "__main__":
    main()
