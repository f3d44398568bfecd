"""Micro-benchmarks of the two tensor-core kernels against the library kernels PyTorch would dispatch to on the same GPU
(BASELINE.json configs[4]: joint-attention sweep 8k-128k, head_dim 128; plus the real joint lengths and the GEMM shapes
of the Wan-14B / CogVideoX-5B blocks).  CUDA-event timing, 3 warm-ups, L2 flushed between iterations.

    python tools/kernel_bench.py [--attn] [--gemm] [--mem] [--bwd] [--quick] > gpurun_out/kernel_bench.json
"""
import argparse
import importlib
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vap = importlib.import_module("video-as-prompt_b200")
ops = vap.ops
DEV = "cuda"
_flush = None


def timeit(fn, iters=5, warmup=3):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        _flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def attn_case(H, J, D, results):
    g = torch.Generator(device=DEV).manual_seed(0)
    qkv = torch.randn((1, J, 3 * H * D), generator=g, device=DEV, dtype=torch.float32).to(torch.bfloat16)
    q, k, v = (qkv[..., i * H * D:(i + 1) * H * D].unflatten(2, (H, D)).transpose(1, 2) for i in range(3))
    flops = 4.0 * H * J * J * D
    rec = dict(kind="attention", H=H, J=J, D=D, flop=flops)
    best, avg = timeit(lambda: ops.attention(q, k, v))
    rec["vap_ms"], rec["vap_tflops"] = best, flops / best / 1e9
    qc, kc, vc = q.contiguous(), k.contiguous(), v.contiguous()
    try:
        best, _ = timeit(lambda: F.scaled_dot_product_attention(qc, kc, vc))
        rec["torch_sdpa_default_ms"], rec["torch_sdpa_default_tflops"] = best, flops / best / 1e9
    except Exception as e:  # noqa: BLE001
        rec["torch_sdpa_default_err"] = str(e)[:200]
    from torch.nn.attention import SDPBackend, sdpa_kernel
    for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION)):
        try:
            with sdpa_kernel(be):
                best, _ = timeit(lambda: F.scaled_dot_product_attention(qc, kc, vc))
            rec[f"torch_sdpa_{name}_ms"], rec[f"torch_sdpa_{name}_tflops"] = best, flops / best / 1e9
        except Exception as e:  # noqa: BLE001
            rec[f"torch_sdpa_{name}_err"] = str(e)[:120]
    results.append(rec)
    print(json.dumps(rec), flush=True)


def attn_bwd_case(H, J, D, results):
    """Backward of the joint attention: vap_attention_bwd (delta + dQ + dK/dV kernels) against torch autograd through SDPA (cuDNN / flash).
    FLOP convention: 2.5 x the forward's 4 H J^2 D (five J x J x D products); the kernels execute seven (S and dP are recomputed)."""
    g = torch.Generator(device=DEV).manual_seed(0)
    qkv = torch.randn((1, J, 3 * H * D), generator=g, device=DEV, dtype=torch.float32).to(torch.bfloat16)
    q, k, v = (qkv[..., i * H * D:(i + 1) * H * D].unflatten(2, (H, D)).transpose(1, 2) for i in range(3))
    go = torch.randn((1, J, H, D), generator=g, device=DEV, dtype=torch.float32).to(torch.bfloat16).transpose(1, 2)
    flops = 10.0 * H * J * J * D
    rec = dict(kind="attention_bwd", H=H, J=J, D=D, flop=flops)
    o, lse = ops.attention(q, k, v, return_lse=True)
    best, _ = timeit(lambda: ops.attention_bwd(q, k, v, o, lse, go))
    rec["vap_ms"], rec["vap_tflops"] = best, flops / best / 1e9
    from torch.nn.attention import SDPBackend, sdpa_kernel
    for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION)):
        try:
            leaves = [t.contiguous().requires_grad_(True) for t in (q, k, v)]
            with sdpa_kernel(be):
                out = F.scaled_dot_product_attention(*leaves)
                best, _ = timeit(lambda: torch.autograd.grad(out, leaves, go, retain_graph=True))
            rec[f"torch_{name}_bwd_ms"], rec[f"torch_{name}_bwd_tflops"] = best, flops / best / 1e9
        except Exception as e:  # noqa: BLE001
            rec[f"torch_{name}_bwd_err"] = str(e)[:120]
    results.append(rec)
    print(json.dumps(rec), flush=True)


def gemm_case(M, N, K, results, epilogue=0):
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn((M, K), generator=g, device=DEV, dtype=torch.float32).to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g, device=DEV, dtype=torch.float32) / K ** 0.5).to(torch.bfloat16)
    b = torch.zeros((N,), device=DEV, dtype=torch.bfloat16)
    out = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
    flops = 2.0 * M * N * K
    rec = dict(kind="gemm", M=M, N=N, K=K, flop=flops, epilogue=epilogue)
    best, _ = timeit(lambda: ops.linear(x, w, b, epilogue=epilogue, out=out))
    rec["vap_ms"], rec["vap_tflops"] = best, flops / best / 1e9
    best, _ = timeit(lambda: F.linear(x, w, b))
    rec["cublas_ms"], rec["cublas_tflops"] = best, flops / best / 1e9
    results.append(rec)
    print(json.dumps(rec), flush=True)


def mem_case(rows, d, results):
    """HBM-bound kernels: the staged (bulk-copy ring) kernels against the register kernels (VAP_NORM_STAGED=0, read per call) and a plain
    device copy of the same bytes on the same box (the measured-peak yardstick)."""
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn((rows, d), generator=g, device=DEV, dtype=torch.float32).to(torch.bfloat16)
    s1p = torch.randn((1, d), device=DEV)
    sh = torch.randn((1, d), device=DEV)
    out = torch.empty_like(x)
    best, _ = timeit(lambda: out.copy_(x))
    rec = dict(kind="device_copy", rows=rows, d=d, bytes=4.0 * rows * d, ms=best, gbs=4.0 * rows * d / best / 1e6)
    results.append(rec)
    print(json.dumps(rec), flush=True)

    def ab(fn):
        os.environ["VAP_NORM_STAGED"] = "1"
        a, _ = timeit(fn)
        os.environ["VAP_NORM_STAGED"] = "0"
        b, _ = timeit(fn)
        os.environ.pop("VAP_NORM_STAGED")
        return a, b

    a, b = ab(lambda: ops.adaln_layernorm(x, eps=1e-6, rounding=0, scale1p=s1p, shift=sh, out=out))
    rec = dict(kind="adaln_layernorm", rows=rows, d=d, bytes=4.0 * rows * d, vap_ms=a, vap_gbs=4.0 * rows * d / a / 1e6, register_kernel_ms=b,
               register_kernel_gbs=4.0 * rows * d / b / 1e6)
    results.append(rec)
    print(json.dumps(rec), flush=True)
    if d % 128 == 0 and d >= 4096:
        H = d // 128
        qkv = torch.randn((rows, 3 * d), generator=g, device=DEV, dtype=torch.float32).to(torch.bfloat16)
        wq = torch.ones(d, device=DEV)
        cos, sin = torch.rand((rows, 64), device=DEV), torch.rand((rows, 64), device=DEV)
        a, b = ab(lambda: ops.qk_norm_rope_(qkv[:, :d], qkv[:, d:2 * d], heads=H, head_dim=128, wq=wq, wk=wq, cos=cos, sin=sin, rows_per_batch=rows,
                                            eps=1e-6, mode=0))
        rec = dict(kind="qk_norm_rope", rows=rows, d=d, bytes=8.0 * rows * d, vap_ms=a, vap_gbs=8.0 * rows * d / a / 1e6, register_kernel_ms=b,
                   register_kernel_gbs=8.0 * rows * d / b / 1e6)
        results.append(rec)
        print(json.dumps(rec), flush=True)
    else:  # CogVideoX LayerNormZero: affine + modulation, bf16 rounding points
        w, bb = torch.randn(d, device=DEV), torch.randn(d, device=DEV)
        a, b = ab(lambda: ops.adaln_layernorm(x, eps=1e-5, rounding=1, ln_w=w, ln_b=bb, scale1p=s1p, shift=sh, out=out))
        rec = dict(kind="adaln_layernorm_cog", rows=rows, d=d, bytes=4.0 * rows * d, vap_ms=a, vap_gbs=4.0 * rows * d / a / 1e6, register_kernel_ms=b,
                   register_kernel_gbs=4.0 * rows * d / b / 1e6)
        results.append(rec)
        print(json.dumps(rec), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--attn", action="store_true")
    ap.add_argument("--gemm", action="store_true")
    ap.add_argument("--mem", action="store_true")
    ap.add_argument("--bwd", action="store_true", help="attention backward (not part of the default set)")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--sp", action="store_true", help="joint-attention shapes one rank sees under P-way Ulysses (H/P heads over the full joint sequence), "
                    "P = 2 / 4 / 8, incl. the automatic split-KV path at 5 heads (BASELINE.json configs[4] at 1/2/4/8 GPUs)")
    a = ap.parse_args()
    if not (a.attn or a.gemm or a.mem or a.bwd or a.sp):
        a.attn = a.gemm = a.mem = True
    results = []
    if a.attn:
        cases = [(40, 40560, 128), (48, 35552, 64)] if a.quick else [(40, 8192, 128), (40, 16384, 128), (40, 32768, 128), (40, 40560, 128),
                                                                       (40, 65536, 128), (5, 151200, 128), (48, 35552, 64)]
        for H, J, D in cases:
            attn_case(H, J, D, results)
    if a.sp:
        for J in (8192, 16384, 32768, 40560, 65536, 131072, 151200):
            for P in (2, 4, 8):
                if J * (40 // P) <= 40560 * 40:  # keep every case under ~30 ms
                    attn_case(40 // P, J, 128, results)
                    results[-1]["ulysses_ranks"] = P
    if a.bwd:
        for H, J, D in ([(40, 16384, 128)] if a.quick else [(40, 16384, 128), (40, 40560, 128), (48, 35552, 64)]):
            attn_bwd_case(H, J, D, results)
    if a.gemm:
        cases = [(20280, 15360, 5120), (20280, 5120, 5120)] if a.quick else [(20280, 15360, 5120), (20280, 5120, 5120), (20280, 13824, 5120),
                                                                              (20280, 5120, 13824), (17776, 9216, 3072), (17776, 12288, 3072), (769, 10240, 5120)]
        for M, N, K in cases:
            gemm_case(M, N, K, results)
    if a.mem:
        mem_case(40560, 5120, results)
        mem_case(20280, 5120, results)  # one stream = one launch of the step
        mem_case(35552, 3072, results)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "kernel_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
