"""Small single-kernel targets for `ncu --set full` (each launch is replayed ~40x, so sizes are kept modest).
    python tools/prof_target.py attn|attn64|gemm|ln|qk
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vap = importlib.import_module("video-as-prompt_b200")
ops = vap.ops
what = sys.argv[1] if len(sys.argv) > 1 else "attn"
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(s, generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)  # noqa: E731
if what in ("attn", "attn64"):
    H, J, D = (16, 16384, 128) if what == "attn" else (16, 16384, 64)
    qkv = rn(1, J, 3 * H * D)
    q, k, v = (qkv[..., i * H * D:(i + 1) * H * D].unflatten(2, (H, D)).transpose(1, 2) for i in range(3))
    for _ in range(3):
        o = ops.attention(q, k, v)
elif what == "gemm":
    x, w, b = rn(20280, 5120), rn(5120, 5120), rn(5120)
    for _ in range(3):
        o = ops.linear(x, w, b)
elif what == "ln":
    x = rn(40560, 5120)
    s1p, sh = torch.randn((1, 5120), device="cuda"), torch.randn((1, 5120), device="cuda")
    for _ in range(3):
        o = ops.adaln_layernorm(x, eps=1e-6, rounding=0, scale1p=s1p, shift=sh)
elif what == "qk":
    qkv = rn(40560, 3 * 5120)
    w = torch.ones(5120, device="cuda")
    cos, sin = torch.rand((40560, 64), device="cuda"), torch.rand((40560, 64), device="cuda")
    for _ in range(3):
        ops.qk_norm_rope_(qkv[:, :5120], qkv[:, 5120:10240], heads=40, head_dim=128, wq=w, wk=w, cos=cos, sin=sin, rows_per_batch=40560, eps=1e-6, mode=0)
torch.cuda.synchronize()
print("done", what)
