"""CPU: the differentiable B2 seam.  `joint_sdpa` under autograd routes through `_JointAttention` (forward keeps O + LSE, backward
calls ops.attention_bwd); here both ops are the torch stand-ins (tests/cpu_standin_ops.py), whose backward restates the CUDA
kernel's formulas — so this pins (a) the autograd wiring (saved tensors, strides, scale, non-contiguous grad_output) and (b) that the
LSE-recompute formulas of attn_bwd_sm100.cu are the gradients of softmax attention, against torch's own autograd of SDPA.
The kernel itself is checked on the GPU (tests/gpu_checks.py:check_attention_bwd)."""
import importlib
import os
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(q_out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    vap = importlib.import_module("video-as-prompt_b200")
    import cpu_standin_ops
    cpu_standin_ops.install(vap)
    torch.set_num_threads(2)
    res = {}
    for name, (B, H, Lq, Lkv, D, scale) in {"joint_d128": (1, 2, 200, 200, 128, None), "cross_d64": (2, 3, 150, 77, 64, 0.2)}.items():
        g = torch.Generator().manual_seed(11)
        # q, k, v as strided head views of one [B, L, 3, H, D]-like buffer, the way the MoT block hands them over
        qkv = [(torch.randn((B, L, H, D), generator=g)).bfloat16().transpose(1, 2).requires_grad_(True) for L in (Lq, Lkv, Lkv)]
        go = torch.randn((B, H, Lq, D), generator=g).bfloat16()
        o = vap.joint_sdpa(*qkv, scale=scale)
        assert o.grad_fn is not None and o.shape == (B, H, Lq, D)
        grads = torch.autograd.grad(o, qkv, go.transpose(1, 2).contiguous().transpose(1, 2))  # a strided grad_output
        ref_in = [t.detach().float().requires_grad_(True) for t in qkv]
        ref_o = torch.nn.functional.scaled_dot_product_attention(*ref_in, scale=scale)
        ref_grads = torch.autograd.grad(ref_o, ref_in, go.float())
        res[name] = [((a.float() - b).abs().max() / b.abs().max()).item() for a, b in zip(grads, ref_grads)]
        res[name + "_fwd"] = ((o.float() - ref_o).abs().max() / ref_o.abs().max()).item()
    # no grad needed -> plain forward, no graph
    with torch.no_grad():
        res["nograd_has_fn"] = vap.joint_sdpa(*[t.detach() for t in qkv]).grad_fn is not None
    q_out.put(res)


def test_joint_sdpa_is_differentiable_and_matches_torch_autograd():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_worker, args=(q,))
    p.start()
    res = q.get(timeout=600)
    p.join(60)
    for name in ("joint_d128", "cross_d64"):
        assert res[name + "_fwd"] < 1e-2, res
        assert all(e < 2e-2 for e in res[name]), res  # bf16 operands (P, dS) and bf16 outputs against fp32 autograd
    assert res["nograd_has_fn"] is False


def test_c_abi_rejects_bad_backward_arguments_without_a_gpu():
    import ctypes
    vap = importlib.import_module("video-as-prompt_b200")
    lib = vap._lib.load()
    strides = (ctypes.c_int64 * 24)(*([8] * 24))
    assert lib.vap_attention_bwd(16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 1, 1, 8, 8, 96, strides, 1.0, 0) == -1
    assert b"head_dim" in lib.vap_last_error()
    assert lib.vap_attention_bwd(16, 16, 16, 16, 16, 0, 16, 16, 16, 16, 1, 1, 8, 8, 128, strides, 1.0, 0) == -1
    assert b"null" in lib.vap_last_error()
