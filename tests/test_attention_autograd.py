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
    # an expanded gradient (o.sum().backward()) must be materialised before it reaches the kernel wrapper
    leaves = [t.detach().clone().requires_grad_(True) for t in qkv]
    seen = {}
    real_bwd = vap.ops.attention_bwd

    def spy(q, k, v, o, lse, dout, **kw):
        seen["strides"] = dout.stride()
        return real_bwd(q, k, v, o, lse, dout, **kw)

    vap.ops.attention_bwd = spy
    vap.joint_sdpa(*leaves, scale=0.2).sum().backward()
    vap.ops.attention_bwd = real_bwd
    res["expanded_grad_strides"] = list(seen["strides"])
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
    assert res["expanded_grad_strides"][-1] == 1 and all(st > 0 for st in res["expanded_grad_strides"]), res


def test_c_abi_rejects_bad_backward_arguments_without_a_gpu():
    import ctypes
    vap = importlib.import_module("video-as-prompt_b200")
    lib = vap._lib.load()
    strides = (ctypes.c_int64 * 24)(*([8] * 24))
    assert lib.vap_attention_bwd(16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 1, 1, 8, 8, 96, strides, 1.0, 0) == -1
    assert b"head_dim" in lib.vap_last_error()
    assert lib.vap_attention_bwd(16, 16, 16, 16, 16, 0, 16, 16, 16, 16, 1, 1, 8, 8, 128, strides, 1.0, 0) == -1
    assert b"null" in lib.vap_last_error()


def _emulate_bwd_kernel(q, k, v, o, lse, dout, scale, dkv: bool):
    """Tile-level restatement of attn_bwd_kernel<D, kDKV> (attn_bwd_sm100.cu) for ONE (batch, head): 128-row resident tile, 128-row
    streamed tiles zero-filled past the end (TMA), the `col >= valid || !own_ok` masking, lse2 / delta = 0 for streamed q rows past
    Lq, P and dS rounded to bf16 before the TMEM-operand products, fp32 accumulation, bf16 outputs of the in-range rows only."""
    T = 128
    BF = torch.bfloat16
    Lq, D = q.shape
    Lkv = k.shape[0]
    delta = (dout.float() * o.float()).sum(-1)
    log2e = 1.4426950408889634
    c = scale * log2e

    def tile(x, r0):
        t = torch.zeros((T, x.shape[1]), dtype=torch.float32)
        n = max(0, min(T, x.shape[0] - r0))
        t[:n] = x[r0:r0 + n].float()
        return t

    own_len, str_len = (Lkv, Lq) if dkv else (Lq, Lkv)
    outs = [torch.zeros((own_len, D), dtype=BF) for _ in range(2 if dkv else 1)]
    for own0 in range(0, own_len, T):
        own_a, own_b = (tile(k, own0), tile(v, own0)) if dkv else (tile(q, own0), tile(dout, own0))
        own_ok = (own0 + torch.arange(T)) < own_len
        acc = [torch.zeros((T, D)) for _ in outs]
        for it in range((str_len + T - 1) // T):
            s0 = it * T
            str_a, str_b = (tile(q, s0), tile(dout, s0)) if dkv else (tile(k, s0), tile(v, s0))
            s_t, dp_t = own_a @ str_a.T, own_b @ str_b.T            # [own rows, streamed rows]
            valid = str_len - s0
            col_ok = torch.arange(T) < valid
            if dkv:   # statistics per column (streamed q rows), 0 past Lq
                rows = s0 + torch.arange(T)
                ok = rows < Lq
                lse2 = torch.where(ok, lse[rows.clamp(max=Lq - 1)] * log2e, torch.zeros(()))[None, :]
                dl = torch.where(ok, delta[rows.clamp(max=Lq - 1)], torch.zeros(()))[None, :]
            else:     # per row (resident q rows), 0 past Lq
                rows = own0 + torch.arange(T)
                ok = rows < Lq
                lse2 = torch.where(ok, lse[rows.clamp(max=Lq - 1)] * log2e, torch.zeros(()))[:, None]
                dl = torch.where(ok, delta[rows.clamp(max=Lq - 1)], torch.zeros(()))[:, None]
            pe = torch.exp2(s_t * c - lse2)
            pe = torch.where(col_ok[None, :] & own_ok[:, None], pe, torch.zeros(()))
            ds = (pe * (dp_t - dl) * scale).to(BF).float()
            pb = pe.to(BF).float()
            if dkv:
                acc[0] += pb @ str_b      # dV += P^T dO   (pb is already [kv rows, q rows])
                acc[1] += ds @ str_a      # dK += dS^T Q
            else:
                acc[0] += ds @ str_a      # dQ += dS K
        n = min(T, own_len - own0)
        for a, out in zip(acc, outs):
            out[own0:own0 + n] = a[:n].to(BF)
    return outs


def test_backward_tile_algorithm_matches_autograd_on_ragged_shapes():
    """The tiling / masking / statistics logic of the backward kernels, emulated tile by tile on the CPU, against torch autograd:
    ragged q and kv lengths (tails in both the resident and the streamed direction), D = 64 and 128, a non-default scale."""
    for (Lq, Lkv, D, scale) in ((130, 5 * 128 + 7, 128, 128 ** -0.5), (300, 77, 64, 0.2), (128, 128, 64, 0.125)):
        g = torch.Generator().manual_seed(Lq + Lkv)
        q, k, v = (torch.randn((L, D), generator=g).bfloat16() for L in (Lq, Lkv, Lkv))
        go = torch.randn((Lq, D), generator=g).bfloat16()
        leaves = [t.float().requires_grad_(True) for t in (q, k, v)]
        s = (leaves[0] @ leaves[1].T) * scale
        o_ref = torch.softmax(s, dim=-1) @ leaves[2]
        ref = torch.autograd.grad(o_ref, leaves, go.float())
        lse = torch.logsumexp(s.detach(), dim=-1)
        o = o_ref.detach().bfloat16()
        (dq,) = _emulate_bwd_kernel(q, k, v, o, lse, go, scale, dkv=False)
        dv, dk = _emulate_bwd_kernel(q, k, v, o, lse, go, scale, dkv=True)
        for name, got, want in (("dq", dq, ref[0]), ("dk", dk, ref[1]), ("dv", dv, ref[2])):
            err = ((got.float() - want).abs().max() / want.abs().max()).item()
            assert err < 2e-2, (Lq, Lkv, D, name, err)
