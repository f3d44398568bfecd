"""GPU parity checks shared by the pytest `-m gpu` suite and tools/gpu_diag.py.

Every check runs the CUDA path through the C ABI (ops.* -> libvap_b200.so) and compares it with the oracle
(oracle/*, evaluated on CPU) on the same seeded inputs, or — at sizes where the CPU oracle would take minutes — with a
size-independent property / torch's own GPU library kernel.  Each returns a dict of measured errors and raises
AssertionError when a tolerance is exceeded.  Tolerances: north_star says per-block max-abs <= 2e-2 relative for bf16
activations; single kernels are held to tighter bounds noted inline.
"""
from __future__ import annotations

import importlib
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cog_oracle, common as ocommon, denoise as odenoise, wan_oracle  # noqa: E402

vap = importlib.import_module("video-as-prompt_b200")
ops, synth = vap.ops, vap.synth
GOLDEN = os.path.join(ROOT, "tests", "golden")
DEV = "cuda"


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def cosine(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm())).item()


def _with_env(env, fn, **kw):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn(**kw)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _randn(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype)


# ---------------------------------------------------------------------------------------------------------------
# tcgen05 descriptor probes
# ---------------------------------------------------------------------------------------------------------------
def check_probe(a_in_tmem: bool, b_mn_major: bool, N: int = 128, K: int = 128, **desc):
    a = _randn((128, K), 1)
    b = _randn((K, N) if b_mn_major else (N, K), 2)
    d = ops.probe_umma(a.to(DEV), b.to(DEV), a_in_tmem=a_in_tmem, b_mn_major=b_mn_major, **desc)
    torch.cuda.synchronize()
    ref = a.float() @ (b.float() if b_mn_major else b.float().t())
    err = rel_err(d, ref)
    assert err < 1e-3, f"probe a_in_tmem={a_in_tmem} b_mn_major={b_mn_major} N={N} K={K}: rel err {err}"
    return dict(err=err)


# ---------------------------------------------------------------------------------------------------------------
# LayerNorm / q-k norm + RoPE
# ---------------------------------------------------------------------------------------------------------------
def check_layernorm_wan(rows=333, d=5120, batch=1, affine=False, modulate=True):
    x = _randn((batch, rows, d), 3, 2.0)
    scale = _randn((batch, 1, d), 4, 0.3, torch.float32)
    shift = _randn((batch, 1, d), 5, 0.3, torch.float32)
    w = (1 + _randn((d,), 6, 0.1, torch.float32)).to(torch.bfloat16)
    b = _randn((d,), 7, 0.1)
    # oracle: transformer_wan_mot.py:620-623 / :668-669
    y = ocommon.fp32_layer_norm(x.float(), w if affine else None, b if affine else None, 1e-6)
    if modulate:
        y = y * (1 + scale) + shift
    ref = y.type_as(x)
    out = ops.adaln_layernorm(x.to(DEV), eps=1e-6, rounding=ops.ROUND_WAN, ln_w=w.float().to(DEV) if affine else None,
                              ln_b=b.float().to(DEV) if affine else None, scale1p=(1 + scale).to(DEV) if modulate else None,
                              shift=shift.to(DEV) if modulate else None)
    err = rel_err(out, ref)
    mism = (out.cpu() != ref).float().mean().item()
    assert err < 8e-3, f"layernorm wan: rel err {err}"
    return dict(err=err, mismatch_frac=mism)


def check_layernorm_cog(rows=226, d=3072, nmod=2):
    x = _randn((nmod, rows, d), 8, 1.5)
    w = (1 + _randn((d,), 9, 0.1, torch.float32)).to(torch.bfloat16)
    b = _randn((d,), 10, 0.1)
    scale = _randn((nmod, d), 11, 0.3)
    shift = _randn((nmod, d), 12, 0.3)
    ref = F.layer_norm(x, (d,), w, b, 1e-5) * (1 + scale)[:, None, :] + shift[:, None, :]  # normalization.py:468
    out = ops.adaln_layernorm(x.reshape(nmod * rows, d).to(DEV), eps=1e-5, rounding=ops.ROUND_COG, ln_w=w.float().to(DEV), ln_b=b.float().to(DEV),
                              scale1p=(1 + scale).float().to(DEV), shift=shift.float().to(DEV), rows_per_batch=rows)
    err = rel_err(out.view(nmod, rows, d), ref)
    mism = (out.view(nmod, rows, d).cpu() != ref).float().mean().item()
    assert err < 8e-3, f"layernorm cog: rel err {err}"
    return dict(err=err, mismatch_frac=mism)


def check_wan_modulation():
    """vap_wan_modulation == (scale_shift_table + temb.float()) with the "+ 1" of the scale chunks, bit for bit (fp32 adds), for the bf16 and
    the fp32 (from_pretrained keeps scale_shift_table in fp32) parameter."""
    d = 5120
    res = {}
    for tdt in (torch.bfloat16, torch.float32):
        table = _randn((1, 6, d), 90, d ** -0.5, tdt).to(DEV)
        temb = _randn((2, 6, d), 91, 0.5).to(DEV)
        ref = table.float() + temb.float()
        ref[:, 1] = 1 + ref[:, 1]
        ref[:, 4] = 1 + ref[:, 4]
        got = ops.wan_modulation(table, temb)
        res[str(tdt)] = bool(torch.equal(got, ref))
    assert all(res.values()), f"wan modulation is not bit-exact: {res}"
    return res


def check_qk_wan(S=300, H=40, D=128, grid=(3, 10, 10)):
    d = H * D
    qkv = _randn((1, S, 3 * d), 13, 1.0)
    wq = (1 + _randn((d,), 14, 0.1, torch.float32)).to(torch.bfloat16)
    wk = (1 + _randn((d,), 15, 0.1, torch.float32)).to(torch.bfloat16)
    frames, gh, gw = grid
    freqs = wan_oracle.wan_rope(D, (1, 2, 2), 1024, (frames, 2 * gh, 2 * gw), ref=True)  # [1,1,300,64] complex128, negative t
    assert freqs.shape[2] == S
    q, k = qkv[..., :d], qkv[..., d:2 * d]
    ref_q = wan_oracle.apply_rope_complex(ocommon.rms_norm_across(q, wq, 1e-6).unflatten(2, (H, -1)).transpose(1, 2), freqs)
    ref_k = wan_oracle.apply_rope_complex(ocommon.rms_norm_across(k, wk, 1e-6).unflatten(2, (H, -1)).transpose(1, 2), freqs)
    g = qkv.to(DEV)
    cos, sin = vap.rope.as_tables(freqs, D, DEV)
    ops.qk_norm_rope_(g[0, :, :d], g[0, :, d:2 * d], heads=H, head_dim=D, wq=wq.float().to(DEV), wk=wk.float().to(DEV), cos=cos, sin=sin,
                      rows_per_batch=S, eps=1e-6, mode=ops.QK_WAN)
    out_q = g[..., :d].unflatten(2, (H, -1)).transpose(1, 2)
    out_k = g[..., d:2 * d].unflatten(2, (H, -1)).transpose(1, 2)
    eq, ek = rel_err(out_q, ref_q), rel_err(out_k, ref_k)
    assert torch.equal(g[..., 2 * d:].cpu(), qkv[..., 2 * d:]), "v must be untouched"
    assert max(eq, ek) < 8e-3, f"qk wan: rel err q {eq} k {ek}"
    # device-built tables == converted reference tables
    t2 = vap.rope.wan_rope_tables(D, (1, 2, 2), (frames, 2 * gh, 2 * gw), ref=True, device=DEV)
    assert torch.allclose(t2[0], cos, atol=1e-6) and torch.allclose(t2[1], sin, atol=1e-6)
    return dict(err_q=eq, err_k=ek)


def check_qk_cog(T=226, S=150, H=48, D=64):
    d = H * D
    L = T + S
    qkv = _randn((1, L, 3 * d), 16, 1.0)
    nw = lambda s: ((1 + _randn((D,), s, 0.1, torch.float32)).to(torch.bfloat16), _randn((D,), s + 100, 0.1))  # noqa: E731
    (wq, bq), (wk, bk) = nw(17), nw(18)
    rope = cog_oracle.cog_rope_3d(D, ((0, 0), (5, 10)), (5, 10), 3, mot_num=1)  # [150, 64] negative temporal positions
    heads = lambda t: t.view(1, L, H, D).transpose(1, 2)  # noqa: E731
    q = F.layer_norm(heads(qkv[..., :d]), (D,), wq, bq, 1e-6)
    k = F.layer_norm(heads(qkv[..., d:2 * d]), (D,), wk, bk, 1e-6)
    q[:, :, T:] = cog_oracle.apply_rope_real(q[:, :, T:], *rope)
    k[:, :, T:] = cog_oracle.apply_rope_real(k[:, :, T:], *rope)
    g = qkv.to(DEV)
    cos, sin = vap.rope.as_tables(tuple(t.to(DEV) for t in rope), D, DEV)
    ops.qk_norm_rope_(g[0, :, :d], g[0, :, d:2 * d], heads=H, head_dim=D, wq=wq.float().to(DEV), bq=bq.float().to(DEV), wk=wk.float().to(DEV),
                      bk=bk.float().to(DEV), cos=cos, sin=sin, rows_per_batch=L, rope_row0=T, eps=1e-6, mode=ops.QK_COG)
    eq, ek = rel_err(heads(g[..., :d]), q), rel_err(heads(g[..., d:2 * d]), k)
    assert max(eq, ek) < 8e-3, f"qk cog: rel err q {eq} k {ek}"
    return dict(err_q=eq, err_k=ek)


# ---------------------------------------------------------------------------------------------------------------
# GEMM (+ epilogues)
# ---------------------------------------------------------------------------------------------------------------
def _gemm_ref(x, w, bias, epilogue, res=None, gate=None, rows_per_batch=None):
    """Oracle arithmetic of the Linear call sites with the reference's rounding points (fp32 accumulate on CPU)."""
    y = (x.float() @ w.float().t() + (bias.float() if bias is not None else 0)).to(torch.bfloat16)  # nn.Linear output is bf16
    if epilogue == ops.EPI_BIAS:
        return y
    if epilogue == ops.EPI_BIAS_GELU:
        return F.gelu(y, approximate="tanh")
    M = x.shape[0]
    g = None
    if gate is not None:
        g = gate.repeat_interleave(rows_per_batch, dim=0)[:M] if gate.dim() == 2 and gate.shape[0] > 1 else gate.reshape(1, -1)
    if epilogue == ops.EPI_GATE_RES_F32:
        return (res.float() + y * g).to(torch.bfloat16)  # transformer_wan_mot.py:658
    if epilogue == ops.EPI_RES_ADD:
        return res + y  # :675
    if epilogue == ops.EPI_GATE_RES_BF16:
        return res + g.to(torch.bfloat16) * y  # cogvideox_transformer_3d_mot.py:445
    raise ValueError(epilogue)


def check_gemm(M=300, N=512, K=256, epilogue=0, bias=True, nbatch=1, seed=20, tol=6e-3):
    x = _randn((M, K), seed, 1.0)
    w = _randn((N, K), seed + 1, 1.0 / math.sqrt(K))
    b = _randn((N,), seed + 2, 0.1) if bias else None
    res = _randn((M, N), seed + 3, 1.0)
    rpb = (M + nbatch - 1) // nbatch
    gate = _randn((nbatch, N), seed + 4, 0.5, torch.float32)
    if epilogue == ops.EPI_GATE_RES_BF16:
        gate = gate.to(torch.bfloat16).float()
    needs_res = epilogue >= ops.EPI_GATE_RES_F32
    needs_gate = epilogue in (ops.EPI_GATE_RES_F32, ops.EPI_GATE_RES_BF16)
    ref = _gemm_ref(x, w, b, epilogue, res, gate, rpb)
    out = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV) if bias else None, epilogue=epilogue, residual=res.to(DEV) if needs_res else None,
                     gate=gate.to(DEV) if needs_gate else None, rows_per_batch=rpb)
    torch.cuda.synchronize()
    err = rel_err(out, ref)
    assert err < tol, f"gemm M={M} N={N} K={K} epi={epilogue}: rel err {err}"
    return dict(err=err, mismatch_frac=(out.cpu() != ref).float().mean().item())


def check_gemm_large(M=20280, N=15360, K=5120):
    """Full Wan-14B fused-QKV size (BASELINE config #3): compare against torch's own bf16 GEMM on the GPU (library
    cross-check; the CPU oracle would need minutes) + linearity property f(2x) == 2 f(x) exactly (power-of-two scaling)."""
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn((M, K), generator=g, device=DEV, dtype=torch.float32).to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g, device=DEV, dtype=torch.float32) / math.sqrt(K)).to(torch.bfloat16)
    b = (torch.randn((N,), generator=g, device=DEV, dtype=torch.float32) * 0.1).to(torch.bfloat16)
    out = ops.linear(x, w, b)
    ref = F.linear(x, w, b)
    err = rel_err(out[::97], ref[::97])
    out2 = ops.linear(x * 2, w, None)
    out1 = ops.linear(x, w, None)
    lin = torch.equal(out2, out1 * 2)
    assert err < 6e-3 and lin, f"gemm large: rel err vs cuBLAS {err}, linearity {lin}"
    return dict(err_vs_cublas=err, linear=lin)


# ---------------------------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------------------------
def check_attention(B=1, H=2, Lq=300, Lkv=300, D=128, joint_layout=True, seed=30, tol=1e-2):
    """vs the definition-level fp32 oracle.  joint_layout=True reads q/k/v as column slices of a [B, L, 3*H*D] buffer,
    exactly how the block forward feeds the kernel."""
    if joint_layout and Lq == Lkv:
        buf = _randn((B, Lq, 3 * H * D), seed, 1.0)
        views = lambda t: [t[..., i * H * D:(i + 1) * H * D].unflatten(2, (H, D)).transpose(1, 2) for i in range(3)]  # noqa: E731
        q, k, v = views(buf)
        gq, gk, gv = views(buf.to(DEV))
    else:
        q, k, v = _randn((B, H, Lq, D), seed), _randn((B, H, Lkv, D), seed + 1), _randn((B, H, Lkv, D), seed + 2)
        gq, gk, gv = q.to(DEV), k.to(DEV), v.to(DEV)
    ref = ocommon.sdpa_explicit_fp32(q, k, v)
    out, lse = ops.attention(gq, gk, gv, return_lse=True)
    torch.cuda.synchronize()
    err = rel_err(out, ref)
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * D ** -0.5
    lse_err = (lse.cpu() - torch.logsumexp(s, dim=-1)).abs().max().item()
    assert out.shape == (B, H, Lq, D) and out.transpose(1, 2).is_contiguous()
    assert err < tol and lse_err < 2e-2, f"attention B={B} H={H} Lq={Lq} Lkv={Lkv} D={D}: rel err {err}, lse err {lse_err}"
    return dict(err=err, lse_err=lse_err)


def check_attention_accumulate(D=128, H=3, Lq=700, kv=(257, 512)):
    """vap_attention_fwd_accumulate (the cross-attention's second softmax added to the first in the epilogue) must equal two plain launches
    followed by torch's bf16 tensor add — bit for bit — in both softmax organisations' epilogues."""
    q = _randn((1, H, Lq, D), 45).to(DEV)
    ks = [_randn((1, H, n, D), 46 + t).to(DEV) for t, n in enumerate(kv)]
    vs = [_randn((1, H, n, D), 48 + t).to(DEV) for t, n in enumerate(kv)]
    want = ops.attention(q, ks[0], vs[0]) + ops.attention(q, ks[1], vs[1])
    got = ops.attention(q, ks[0], vs[0])
    ops.attention(q, ks[1], vs[1], out=got, accumulate=True)
    ref = ocommon.sdpa_explicit_fp32(q.cpu(), ks[0].cpu(), vs[0].cpu()).to(torch.bfloat16) + ocommon.sdpa_explicit_fp32(q.cpu(), ks[1].cpu(), vs[1].cpu()).to(torch.bfloat16)
    err = rel_err(got, ref)
    assert torch.equal(got, want), f"accumulate epilogue differs from two launches + bf16 add: {rel_err(got, want)}"
    assert err < 1e-2, f"accumulated cross-attention vs the oracle: {err}"
    return dict(bit_exact=True, err=err)


def check_attention_variants():
    """The two opt-in organisations of the forward kernel — one thread per query row (VAP_ATTN_SOFTMAX=row) and the CTA-pair kernel on
    tcgen05.mma cta_group::2 (VAP_ATTN_PAIR=1, D = 128) — against the oracle on the shapes the default kernel is checked on (the switches are
    read per call).  Both are measured slower or equal and stay opt-in; this keeps them correct."""
    res = {}
    for name, env in (("row", {"VAP_ATTN_SOFTMAX": "row"}), ("pair", {"VAP_ATTN_PAIR": "1"})):
        res[name + "_d128"] = _with_env(env, check_attention, B=1, H=3, Lq=1000, Lkv=1000, D=128)["err"]
        res[name + "_cross"] = _with_env(env, check_attention, B=1, H=2, Lq=600, Lkv=257, D=128, joint_layout=False)["err"]
        res[name + "_peaky"] = _with_env(env, check_attention_peaky)["err"]
        res[name + "_splitkv"] = _with_env(env, check_attention_splitkv, B=1, H=2, Lq=300, Lkv=1000, D=128, splits=2)["err"]
    res["row_d64"] = _with_env({"VAP_ATTN_SOFTMAX": "row"}, check_attention, B=2, H=4, Lq=452, Lkv=452, D=64)["err"]
    return res


def check_attention_short_and_long():
    """The short-KV kernel (one Q tile per CTA, two CTAs per SM; chosen by itself up to 8 KV tiles) and the two-tile ping-pong kernel, each FORCED
    on the shapes the other one gets by default (VAP_ATTN_SHORT is read per call): cross-attention shapes through the long kernel, multi-tile /
    peaky / accumulate / split-KV / peer-store / D = 64 shapes through the short one."""
    res = {}
    on, off = {"VAP_ATTN_SHORT": "1"}, {"VAP_ATTN_SHORT": "0"}
    res["long_cross_257"] = _with_env(off, check_attention, B=1, H=2, Lq=600, Lkv=257, D=128, joint_layout=False)["err"]
    res["long_cross_512"] = _with_env(off, check_attention, B=1, H=2, Lq=300, Lkv=512, D=128, joint_layout=False)["err"]
    res["long_one_tile"] = _with_env(off, check_attention, B=1, H=1, Lq=64, Lkv=100, D=128, joint_layout=False)["err"]
    res["long_accumulate"] = _with_env(off, check_attention_accumulate)["err"]
    res["short_multi_tile"] = _with_env(on, check_attention, B=1, H=3, Lq=1000, Lkv=1000, D=128)["err"]
    res["short_17_tiles"] = _with_env(on, check_attention, B=1, H=2, Lq=300, Lkv=2100, D=128, joint_layout=False)["err"]
    res["short_d64"] = _with_env(on, check_attention, B=2, H=4, Lq=452, Lkv=900, D=64, joint_layout=False)["err"]
    res["short_peaky"] = _with_env(on, check_attention_peaky)["err"]
    res["short_accumulate_d64"] = _with_env(on, check_attention_accumulate, D=64, H=2, Lq=300, kv=(100, 226))["err"]
    res["short_splitkv"] = _with_env(on, check_attention_splitkv, B=1, H=2, Lq=300, Lkv=1000, D=128, splits=2)["err"]
    res["short_p2p_emulated"] = _with_env(on, check_ulysses_p2p_emulated, P=4, L=96, H=8, D=128, mode=0)["err"]
    return res


def check_attention_peaky(D=128):
    """Scores with a large dynamic range (exercises the lazy O-rescale path: the running max keeps growing by > 2^8)."""
    H, L = 2, 1024
    q = _randn((1, H, L, D), 40, 3.0)
    k = _randn((1, H, L, D), 41, 3.0)
    k[:, :, 1::128] *= 4  # a few dominant keys appearing late in every tile sequence
    v = _randn((1, H, L, D), 42)
    ref = ocommon.sdpa_explicit_fp32(q, k, v)
    out = ops.attention(q.to(DEV), k.to(DEV), v.to(DEV))
    err = rel_err(out, ref)
    assert err < 1.5e-2, f"attention peaky: rel err {err}"
    return dict(err=err)


def check_attention_full_size(J=40560, H=40, D=128, heads_checked=2):
    """BASELINE config #3 joint attention size.  Size-independent properties + library cross-check on the GPU:
    (1) V = 1  =>  O = 1 exactly-ish (softmax rows sum to one);  (2) permuting the K/V rows leaves O unchanged (the joint
    order [target|ref] is irrelevant for unmasked attention, SURVEY §8 note 1);  (3) first heads vs torch SDPA."""
    g = torch.Generator(device=DEV).manual_seed(1)
    qkv = torch.randn((1, J, 3 * H * D), generator=g, device=DEV, dtype=torch.float32).to(torch.bfloat16)
    q, k, v = (qkv[..., i * H * D:(i + 1) * H * D].unflatten(2, (H, D)).transpose(1, 2) for i in range(3))
    out = ops.attention(q, k, v)
    ones = torch.ones_like(v[:, :1].contiguous())
    o1 = ops.attention(q[:, :1], k[:, :1], ones)
    p1 = (o1.float() - 1).abs().max().item()
    perm = torch.randperm(J, device=DEV, generator=g)
    o_perm = ops.attention(q[:, :1], k[:, :1, perm].contiguous(), v[:, :1, perm].contiguous())
    p2 = rel_err(o_perm, out[:, :1])
    hs = slice(0, heads_checked)
    ref = F.scaled_dot_product_attention(q[:, hs].contiguous(), k[:, hs].contiguous(), v[:, hs].contiguous())
    p3 = rel_err(out[:, hs], ref)
    assert p1 < 1e-2 and p2 < 1e-2 and p3 < 1.5e-2, f"attention full size: ones {p1}, permutation {p2}, vs torch SDPA {p3}"
    return dict(ones_err=p1, perm_err=p2, err_vs_torch_sdpa=p3)


def check_attention_splitkv(B=1, H=2, Lq=300, Lkv=1000, D=128, splits=2, seed=33, tol=1e-2):
    """Split-KV attention + merge (vap_attention_fwd_splitkv / vap_attention_combine) vs the definition-level fp32 oracle and vs
    the unsplit kernel; the KV tail (Lkv not a multiple of 128) lands in the last range."""
    q, k, v = _randn((B, H, Lq, D), seed), _randn((B, H, Lkv, D), seed + 1), _randn((B, H, Lkv, D), seed + 2)
    gq, gk, gv = q.to(DEV), k.to(DEV), v.to(DEV)
    ref = ocommon.sdpa_explicit_fp32(q, k, v)
    out = ops.attention_splitkv(gq, gk, gv, splits)
    base = ops.attention(gq, gk, gv)
    torch.cuda.synchronize()
    err, err_base = rel_err(out, ref), rel_err(out, base)
    assert out.shape == (B, H, Lq, D) and out.transpose(1, 2).is_contiguous()
    assert err < tol and err_base < tol, f"split-KV attention H={H} Lq={Lq} Lkv={Lkv} D={D} splits={splits}: rel err {err}, vs unsplit {err_base}"
    return dict(err=err, err_vs_unsplit=err_base)


def check_attention_splitkv_peers(P=4, L=64, H=2, D=128, splits=2):
    """The merge kernel in peer mode (Ulysses exchange #2): rows scattered to P emulated ranks' buffers must equal — bit for bit —
    the merge into one plain tensor."""
    J = P * L
    q, k, v = (_randn((1, H, J, D), 70 + i).to(DEV) for i in range(3))
    plain = ops.attention_splitkv(q, k, v, splits).transpose(1, 2).flatten(2, 3)[0]  # [J, H*D]
    out = [torch.zeros((L, H * D), dtype=torch.bfloat16, device=DEV) for _ in range(P)]
    ops.attention_splitkv(q, k, v, splits, o_ptrs=[t.data_ptr() for t in out], rows_per_peer=L, o_strides=(0, D, H * D))
    torch.cuda.synchronize()
    assert torch.equal(torch.cat(out, 0), plain), "split-KV merge: peer scatter differs from the plain output"
    return dict(bit_exact=True)


def check_attention_splitkv_sp8_shape(J=40560, H=5, D=128):
    """The shape one rank sees under 8-way Ulysses at BASELINE config #3 (5 heads, 795 work items = 5.37 waves on 148 SMs), where
    ops.attention cuts the KV sequence by itself: the automatic path must agree with the unsplit kernel, and both are timed."""
    g = torch.Generator(device=DEV).manual_seed(2)
    qkv = torch.randn((1, J, 3 * H * D), generator=g, device=DEV, dtype=torch.float32).to(torch.bfloat16)
    q, k, v = (qkv[..., i * H * D:(i + 1) * H * D].unflatten(2, (H, D)).transpose(1, 2) for i in range(3))
    auto = ops.attention_kv_splits(1, H, J, J)
    base, _ = ops.attention(q, k, v, return_lse=True)  # return_lse keeps the unsplit kernel
    out = ops.attention(q, k, v)

    def ms(fn, n=5):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    t_base, t_auto = ms(lambda: ops.attention(q, k, v, return_lse=True)), ms(lambda: ops.attention(q, k, v))
    err = rel_err(out, base)
    assert err < 1e-2, f"automatic split-KV ({auto} ranges) vs unsplit: rel err {err}"
    flop = 4.0 * H * J * J * D
    return dict(auto_splits=auto, err_vs_unsplit=err, unsplit_ms=t_base, auto_ms=t_auto, unsplit_tflops=flop / t_base / 1e9, auto_tflops=flop / t_auto / 1e9)


# ---------------------------------------------------------------------------------------------------------------
# block / model level against the golden fixtures recorded from the reference
# ---------------------------------------------------------------------------------------------------------------
def _golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def _to_dev(x):
    if torch.is_tensor(x):
        return x.to(DEV)
    if isinstance(x, (list, tuple)):
        return type(x)(_to_dev(t) for t in x)
    if isinstance(x, dict):
        return {k: _to_dev(v) for k, v in x.items()}
    return x


def build_wan(cfg, seed):
    m = vap.WanTransformer3DMOTModel(**cfg).to(torch.bfloat16)
    synth.fill_module_(m, seed=seed, num_layers=cfg["num_layers"])
    return m.to(DEV).eval()


def build_cog(cfg, seed):
    m = vap.CogVideoXTransformer3DMOTModel(**cfg).to(torch.bfloat16)
    synth.fill_module_(m, seed=seed, num_layers=cfg["num_layers"])
    return m.to(DEV).eval()


def check_wan_blocks(tol=2e-2):
    """Teacher-forced: block i of the CUDA path on the reference's recorded inputs of block i vs its recorded outputs."""
    g = _golden("wan_tiny.pt")
    cfg = g["cfg"]
    model = build_wan(cfg, g["weight_seed"])
    f, h, w = g["latent"]
    rope = vap.rope.wan_rope_tables(128, cfg["patch_size"], (f, h, w), ref=False, device=DEV)
    rope_r = vap.rope.wan_rope_tables(128, cfg["patch_size"], (f, h, w), ref=True, device=DEV)
    sh = {k: v.to(DEV) for k, v in g["shared"].items()}
    errs = {}
    with torch.no_grad():
        for i, blk in g["blocks"].items():
            x, xr = model.blocks[i](hidden_states=blk["hidden_states"].to(DEV), encoder_hidden_states=sh["encoder_hidden_states"], temb=sh["temb"],
                                    rotary_emb=rope, hidden_states_mot_ref=blk["hidden_states_mot_ref"].to(DEV),
                                    encoder_hidden_states_mot_ref=sh["encoder_hidden_states_mot_ref"], temb_mot_ref=sh["temb_mot_ref"],
                                    rotary_emb_mot_ref=rope_r, num_mot_ref=1)
            errs[f"block{i}"] = rel_err(x, blk["out"])
            errs[f"block{i}_ref"] = rel_err(xr, blk["out_ref"])
    assert max(errs.values()) <= tol, f"wan blocks: {errs}"
    return errs


def check_wan_model(tol=2e-2):
    g = _golden("wan_tiny.pt")
    cfg = g["cfg"]
    model = build_wan(cfg, g["weight_seed"])
    inp = _to_dev(synth.wan_inputs(cfg, *g["latent"], seed=g["input_seed"]))
    with torch.no_grad():
        out = model(**inp, return_dict=False)[0]
    err, cos = rel_err(out, g["final"]), cosine(out, g["final"])
    assert err <= tol and cos >= 0.999, f"wan model: rel err {err}, cosine {cos}"
    return dict(err=err, cosine=cos)


def check_wan_denoise():
    """north_star: final denoised latents cosine >= 0.999 after 4 steps (with CFG, FlowMatchEuler shift 3)."""
    g = _golden("wan_tiny.pt")
    cfg, dn = g["cfg"], g["denoise"]
    model = build_wan(cfg, g["weight_seed"])
    f, h, w = g["latent"]
    inp = synth.wan_inputs(cfg, f, h, w, seed=g["input_seed"])
    neg = synth.wan_inputs(cfg, f, h, w, seed=dn["neg_seed"])
    gen = torch.Generator().manual_seed(dn["seed"])
    lat0 = torch.randn((1, 16, f, h, w), generator=gen)
    lat_ref = torch.randn((1, 16, f, h, w), generator=gen)
    kw = {k: inp[k] for k in ("encoder_hidden_states", "encoder_hidden_states_image", "encoder_hidden_states_mot_ref",
                              "encoder_hidden_states_image_mot_ref", "num_mot_ref")}
    kw_u = dict(kw, encoder_hidden_states=neg["encoder_hidden_states"], encoder_hidden_states_mot_ref=neg["encoder_hidden_states_mot_ref"])
    with torch.no_grad():
        lat = vap.denoise.wan_denoise(model, lat0.to(DEV), inp["hidden_states"][:, 16:].float().to(DEV), lat_ref.to(DEV),
                                      inp["hidden_states_mot_ref"][:, 16:].float().to(DEV), _to_dev(kw), _to_dev(kw_u), dn["steps"], dn["shift"],
                                      dn["guidance"], cache_context=False)  # the cached loop: CHECKS_PENDING["wan_denoise_cached"]
    cos, err = cosine(lat, dn["final_latents"]), rel_err(lat, dn["final_latents"])
    assert cos >= 0.999, f"wan 4-step denoise: cosine {cos} (rel err {err})"
    return dict(cosine=cos, err=err)


def check_cog_blocks(case="small", tol=2e-2):
    g = _golden("cog_tiny.pt")
    cfg, c = g["cfg"], g["cases"][case]
    model = build_cog(cfg, g["weight_seed"])
    inp = synth.cog_inputs(cfg, *c["latent"], seed=c["input_seed"], num_mot_ref=c["num_mot_ref"])
    rope, rope_r = _to_dev(inp["image_rotary_emb"]), _to_dev(inp["image_rotary_emb_mot_ref"])
    sh = c["shared"]
    errs = {}
    with torch.no_grad():
        for i, blk in c["blocks"].items():
            outs = model.transformer_blocks[i](
                hidden_states=blk["hidden_states"].to(DEV), encoder_hidden_states=blk["encoder_hidden_states"].to(DEV), temb=sh["temb"].to(DEV),
                image_rotary_emb=rope, hidden_states_mot_ref=blk["hidden_states_mot_ref"].to(DEV),
                encoder_hidden_states_mot_ref=blk["encoder_hidden_states_mot_ref"].to(DEV),
                temb_mot_ref=None if sh["temb_mot_ref"] is None else sh["temb_mot_ref"].to(DEV),
                temb_list_mot_ref=None if sh["temb_list_mot_ref"] is None else [t.to(DEV) for t in sh["temb_list_mot_ref"]],
                image_rotary_emb_mot_ref=rope_r)
            for n, o in zip(("out_v", "out_e", "out_v_ref", "out_e_ref"), outs):
                if blk[n] is not None:
                    errs[f"block{i}_{n}"] = rel_err(o, blk[n])
    assert max(errs.values()) <= tol, f"cog blocks[{case}]: {errs}"
    return errs


def check_cog_model(case="config1", tol=2e-2):
    g = _golden("cog_tiny.pt")
    cfg, c = g["cfg"], g["cases"][case]
    model = build_cog(cfg, g["weight_seed"])
    inp = synth.cog_inputs(cfg, *c["latent"], seed=c["input_seed"], num_mot_ref=c["num_mot_ref"])
    if c["multi"]:
        inp["timestep_list_mot_ref"] = [torch.full((1,), t) for t in c["timestep_list"]]
    with torch.no_grad():
        out = model(**_to_dev(inp), return_dict=False)[0]
    err, cos = rel_err(out, c["final"]), cosine(out, c["final"])
    assert err <= tol and cos >= 0.999, f"cog model[{case}]: rel err {err}, cosine {cos}"
    return dict(err=err, cosine=cos)


def check_cog_denoise():
    """north_star: final denoised latents cosine >= 0.999 after 4 steps — CogVideoX loop: CFG as one B=2 forward per step, dynamic
    guidance, CogVideoXDPMScheduler (fixture recorded from the reference scheduler + reference transformer on CPU)."""
    g = _golden("cog_tiny.pt")
    cfg, dn = g["cfg"], g["cases"]["denoise"]
    model = build_cog(cfg, g["weight_seed"])
    f, h, w = dn["latent"]
    inp = synth.cog_inputs(cfg, f, h, w, seed=dn["input_seed"], batch=2)
    kw2 = _to_dev({k: inp[k] for k in ("encoder_hidden_states", "encoder_hidden_states_mot_ref", "image_rotary_emb", "image_rotary_emb_mot_ref", "num_mot_ref")})
    gen = torch.Generator().manual_seed(dn["latent_seed"])
    lat0, img, lat_ref, img_ref = (torch.randn((1, f, 16, h, w), generator=gen).to(DEV) for _ in range(4))
    lat = vap.denoise.cog_denoise(model, lat0, img, lat_ref, img_ref, kw2, dn["steps"], dn["guidance"], dn["dynamic_cfg"], dn["noise_seed"])
    cos, err = cosine(lat, dn["final_latents"]), rel_err(lat, dn["final_latents"])
    assert cos >= 0.999, f"cog 4-step DPM denoise: cosine {cos} (rel err {err})"
    return dict(cosine=cos, err=err)


def check_processor_level():
    """Boundary B1/B2: the drop-in processors + joint_sdpa reproduce the fused block (same kernels, reference-shaped calls)."""
    g = _golden("wan_tiny.pt")
    cfg = g["cfg"]
    model = build_wan(cfg, g["weight_seed"])
    blk = model.blocks[0]
    f, h, w = g["latent"]
    rope = vap.rope.wan_rope_tables(128, cfg["patch_size"], (f, h, w), ref=False, device=DEV)
    x = g["blocks"][0]["hidden_states"].to(DEV).contiguous()
    with torch.no_grad():
        q, k, v, _ = blk.attn1(hidden_states=x, rotary_emb=rope, is_before_attn=True)
        o = vap.joint_sdpa(q, k, v)
        y = blk.attn1(hidden_states=o, is_before_attn=False)
    ref_q, ref_k, ref_v = wan_oracle.self_attn_pre({k_: v_.cpu() for k_, v_ in blk.state_dict().items() if k_.startswith("attn1.")}, "attn1",
                                                   x.cpu(), cfg["num_attention_heads"], cfg["eps"],
                                                   wan_oracle.wan_rope(128, cfg["patch_size"], 1024, (f, h, w), ref=False))
    e = max(rel_err(q, ref_q), rel_err(k, ref_k), rel_err(v, ref_v))
    assert e < 1e-2 and y.shape == x.shape, f"processor level: qkv rel err {e}"
    try:
        vap.joint_sdpa(q, k, v, is_causal=True)
        raise AssertionError("joint_sdpa must reject is_causal=True")
    except ValueError:
        pass
    return dict(err=e)


def check_ulysses_relayout():
    L, P, chunk = 37, 4, 64
    src = _randn((L, P * chunk), 50).to(DEV)
    out = torch.empty((P, L, 3, chunk), dtype=torch.bfloat16, device=DEV)
    for wi in range(3):
        ops.ulysses_pack(src, P, out[:, :, wi, :])
    ref = src.view(L, P, chunk).permute(1, 0, 2)
    ok = all(torch.equal(out[:, :, wi, :], ref) for wi in range(3))
    back = ops.ulysses_unpack(out[:, :, 1, :])
    assert ok and torch.equal(back, src), "ulysses pack/unpack round trip"
    return dict(ok=True)


def check_ulysses_p2p_emulated(P=4, L=96, H=8, D=128, mode=0):
    """The fused Ulysses exchange (vap_qkv_scatter + vap_attention_fwd_scatter) with P ranks EMULATED on one device: rank r's
    kernels get the other ranks' buffers as "peer" pointers.  Must reproduce — bit for bit — q/k-norm + RoPE in place followed
    by one attention over the whole sequence, because every rank computes the same per-head arithmetic on the same rows."""
    d = H * D
    J = P * L
    qkv = _randn((J, 3 * d), 60).to(DEV)
    g = torch.Generator().manual_seed(61)
    wq, wk = (torch.rand(d if mode == 0 else D, generator=g) + 0.5).to(DEV), (torch.rand(d if mode == 0 else D, generator=g) + 0.5).to(DEV)
    bq = bk = None
    if mode == 1:
        bq, bk = (torch.randn(D, generator=g) * 0.1).to(DEV), (torch.randn(D, generator=g) * 0.1).to(DEV)
    ang = torch.rand((J, D // 2), generator=g) * 6.28
    cos, sin = torch.cos(ang).to(DEV).contiguous(), torch.sin(ang).to(DEV).contiguous()
    # reference: one device, no exchange.  The scatter mode runs on the register kernel; the in-place reference is pinned to the same kernel
    # (VAP_NORM_STAGED is read per call) because the staged kernel sums a row's squares in another order — equal within fp32 rounding
    # (qk_wan_staged checks it against the oracle), but this check is about the exchange and asks for bit-exactness.
    ref_qkv = qkv.clone()
    _with_env({"VAP_NORM_STAGED": "0"}, ops.qk_norm_rope_, q=ref_qkv[:, :d], k=ref_qkv[:, d:2 * d], heads=H, head_dim=D, wq=wq, wk=wk, bq=bq, bk=bk, cos=cos,
              sin=sin, rows_per_batch=J, eps=1e-6, mode=mode)
    q, k, v = (ref_qkv[None, :, i * d:(i + 1) * d].unflatten(2, (H, D)).transpose(1, 2) for i in range(3))
    o_ref = ops.attention(q, k, v).transpose(1, 2).flatten(2, 3)[0]  # [J, d]
    # emulated ranks
    hp = d // P
    recv = [torch.zeros((P, L, 3, hp), dtype=torch.bfloat16, device=DEV) for _ in range(P)]
    out = [torch.zeros((L, d), dtype=torch.bfloat16, device=DEV) for _ in range(P)]
    half = L // 2
    for r in range(P):
        loc = qkv[r * L:(r + 1) * L]
        for row0, n in ((0, half), (half, L - half)):  # two "streams" per rank, as the MoT block dispatches them
            sl = slice(row0, row0 + n)
            ops.qkv_scatter(loc[sl, :d], loc[sl, d:2 * d], loc[sl, 2 * d:], heads=H, head_dim=D, wq=wq, wk=wk, bq=bq, bk=bk,
                            cos=cos[r * L + row0:r * L + row0 + n].contiguous(), sin=sin[r * L + row0:r * L + row0 + n].contiguous(), rows_per_batch=n,
                            eps=1e-6, mode=mode, dst_ptrs=[t.data_ptr() for t in recv], dst_slot=r, slot_rows=L, dst_row0=row0)
    assert torch.equal(qkv, _randn((J, 3 * d), 60).to(DEV)), "scatter mode must not modify its inputs"
    for s_ in range(P):
        joint = recv[s_].view(1, J, 3, H // P, D)
        qs, ks, vs = (joint[:, :, w].transpose(1, 2) for w in range(3))
        ops.attention_scatter(qs, ks, vs, o_ptrs=[t.data_ptr() + s_ * hp * 2 for t in out], rows_per_peer=L, o_strides=(0, D, d))
    o = torch.cat(out, dim=0)
    exact = torch.equal(o, o_ref)
    err = rel_err(o, o_ref)
    assert err < 1e-6, f"emulated p2p Ulysses differs from the single-device path: {err}"
    return dict(err=err, bit_exact=exact)


def check_ulysses_p2p_emulated_splitkv():
    """The emulated fused exchange with the split-KV path forced on (what 8-way Ulysses uses at 480p: 5 heads per rank): the merge
    kernel does the peer stores.  Still bit-exact against the single-device path, which splits the same way."""
    old = os.environ.get("VAP_ATTN_SPLITKV")
    os.environ["VAP_ATTN_SPLITKV"] = "2"
    ops._SPLIT_CACHE.clear()
    try:
        return check_ulysses_p2p_emulated(P=4, L=96, H=8, D=128, mode=0)
    finally:
        if old is None:
            del os.environ["VAP_ATTN_SPLITKV"]
        else:
            os.environ["VAP_ATTN_SPLITKV"] = old
        ops._SPLIT_CACHE.clear()


def check_attention_bwd(B=1, H=2, Lq=300, Lkv=300, D=128, joint_layout=True, seed=80, tol=2e-2):
    """vap_attention_bwd (dQ, dK, dV from dO + the forward's LSE) against torch autograd of the definition-level fp32 attention on the
    GPU, called both directly and through the differentiable B2 seam (`joint_sdpa` under autograd).  Gate: max-abs <= 2e-2 relative
    per gradient (P and dS are bf16 tensor-core operands, the outputs are bf16)."""
    if joint_layout:
        qkv = _randn((B, Lq, 3, H, D), seed).to(DEV)
        q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
    else:
        q = _randn((B, Lq, H, D), seed).to(DEV).transpose(1, 2)
        k = _randn((B, Lkv, H, D), seed + 1).to(DEV).transpose(1, 2)
        v = _randn((B, Lkv, H, D), seed + 2).to(DEV).transpose(1, 2)
    go = _randn((B, Lq, H, D), seed + 3).to(DEV).transpose(1, 2)
    ref_in = [t.detach().float().requires_grad_(True) for t in (q, k, v)]
    s_ref = torch.matmul(ref_in[0], ref_in[1].transpose(-1, -2)) * D ** -0.5
    o_ref = torch.matmul(torch.softmax(s_ref, dim=-1), ref_in[2])
    ref = torch.autograd.grad(o_ref, ref_in, go.float())
    o, lse = ops.attention(q, k, v, return_lse=True)
    got = ops.attention_bwd(q, k, v, o, lse, go)
    errs = [rel_err(a, b) for a, b in zip(got, ref)]
    # through autograd (what a trainer does): leaves that require grad, joint_sdpa, backward
    leaves = [t.detach().clone().requires_grad_(True) for t in (q, k, v)]
    vap.joint_sdpa(*leaves).backward(go)
    errs_seam = [rel_err(t.grad, b) for t, b in zip(leaves, ref)]
    assert max(errs) <= tol and max(errs_seam) <= tol, f"attention bwd B={B} H={H} Lq={Lq} Lkv={Lkv} D={D}: dq/dk/dv rel err {errs}, via autograd {errs_seam}"
    return dict(dq=errs[0], dk=errs[1], dv=errs[2], seam=max(errs_seam))


def check_attention_bwd_split1(**kw):
    """The backward's simplest configuration — ONE set of elementwise warps, no S prefetch (the defaults are two sets + prefetch; the
    switches are read per call)."""
    return _with_env({"VAP_ATTN_BWD_SPLIT": "1", "VAP_ATTN_BWD_PREFETCH": "0"}, check_attention_bwd, **kw)


def check_attention_bwd_noprefetch(**kw):
    """Two elementwise sets, dQ kernel without the double-buffered S."""
    return _with_env({"VAP_ATTN_BWD_SPLIT": "2", "VAP_ATTN_BWD_PREFETCH": "0"}, check_attention_bwd, **kw)


def check_cfg_flow_match_step(B=2, inner=16 * 3 * 16 * 16):
    """vap_cfg_flow_match_step against the reference's own tensor expression evaluated by torch ON THE GPU (integer-exact comparison
    of the bf16 results): pipeline_wan_i2v_mot.py:874 + scheduling_flow_match_euler_discrete.py:433-467.  The scheduler's dt is a 0-dim
    fp32 tensor; torch casts such an operand to the tensor operand's dtype (bf16) before the multiply (verified on the CPU; expected on
    CUDA too, where only CPU-scalar operands keep fp32), so the kernel is handed the bf16-rounded dt.  The check also records whether the
    unrounded dt would have matched, in case the CUDA kernel of this torch build behaves differently."""
    c, u = _randn((B, inner), 71).to(DEV), _randn((B, inner), 72).to(DEV)
    sig = vap.denoise.flow_match_schedule(4, 3.0, device=DEV)[1]
    sig_h = vap.denoise.flow_match_schedule(4, 3.0, device="cpu")[1]
    res = {}
    for tag, sample in (("f32", _randn((B, inner), 73, dtype=torch.float32).to(DEV)), ("bf16", _randn((B, inner), 74).to(DEV))):
        for i in (0, 2):
            noise = u + 5.0 * (c - u)
            ref = (sample.to(torch.float32) + (sig[i + 1] - sig[i]) * noise).to(noise.dtype)
            ref1 = (sample.to(torch.float32) + (sig[i + 1] - sig[i]) * c).to(c.dtype)
            dt = float(sig_h[i + 1] - sig_h[i])
            got = ops.cfg_flow_match_step(c, u, sample, guidance_scale=5.0, dt=dt)
            got_b = ops.cfg_flow_match_step(c, u, sample, guidance_scale=5.0, dt=float(torch.tensor(dt).bfloat16()))
            got1 = ops.cfg_flow_match_step(c, None, sample, guidance_scale=1.0, dt=float(torch.tensor(dt).bfloat16()))
            res[f"{tag}_{i}"] = dict(exact_fp32_dt=bool(torch.equal(got, ref)), exact_bf16_dt=bool(torch.equal(got_b, ref)), no_cfg=bool(torch.equal(got1, ref1)),
                                     err=rel_err(got, ref))
    # strided output: the latent channels of the next step's transformer input
    x_in = torch.zeros((B, 36, 3, 16, 16), dtype=torch.bfloat16, device=DEV)
    c5, u5, s5 = c.view(B, 16, 3, 16, 16), u.view(B, 16, 3, 16, 16), _randn((B, 16, 3, 16, 16), 75).to(DEV)
    ops.cfg_flow_match_step(c5, u5, s5, guidance_scale=5.0, dt=-0.25, out=x_in[:, :16])
    plain = ops.cfg_flow_match_step(c5, u5, s5, guidance_scale=5.0, dt=-0.25)
    assert torch.equal(x_in[:, :16], plain) and x_in[:, 16:].abs().max().item() == 0, "strided output"
    assert all(r["exact_bf16_dt"] and r["no_cfg"] for r in res.values()), f"cfg_flow_match_step is not bit-exact with torch on the GPU: {res}"
    return res


def check_wan_denoise_fused():
    """wan_denoise(fused_step=True) must reproduce the unfused loop on the same device bit for bit (same kernels, same rounding points)."""
    g = _golden("wan_tiny.pt")
    cfg, dn = g["cfg"], g["denoise"]
    model = build_wan(cfg, g["weight_seed"])
    f, h, w = g["latent"]
    inp = synth.wan_inputs(cfg, f, h, w, seed=g["input_seed"])
    neg = synth.wan_inputs(cfg, f, h, w, seed=dn["neg_seed"])
    gen = torch.Generator().manual_seed(dn["seed"])
    lat0 = torch.randn((1, 16, f, h, w), generator=gen)
    lat_ref = torch.randn((1, 16, f, h, w), generator=gen)
    kw = {k: inp[k] for k in ("encoder_hidden_states", "encoder_hidden_states_image", "encoder_hidden_states_mot_ref",
                              "encoder_hidden_states_image_mot_ref", "num_mot_ref")}
    kw_u = dict(kw, encoder_hidden_states=neg["encoder_hidden_states"], encoder_hidden_states_mot_ref=neg["encoder_hidden_states_mot_ref"])
    outs = []
    with torch.no_grad():
        for fused in (False, True):
            outs.append(vap.denoise.wan_denoise(model, lat0.to(DEV), inp["hidden_states"][:, 16:].float().to(DEV), lat_ref.to(DEV),
                                                inp["hidden_states_mot_ref"][:, 16:].float().to(DEV), _to_dev(kw), _to_dev(kw_u), dn["steps"], dn["shift"],
                                                dn["guidance"], fused_step=fused, cache_context=False))
    cos = cosine(outs[1], dn["final_latents"])
    assert torch.equal(outs[0], outs[1]), f"fused step differs from the torch step: rel err {rel_err(outs[1], outs[0])}"
    assert cos >= 0.999
    return dict(cosine=cos, bit_exact=True)


def check_wan_denoise_cached():
    """wan_denoise with the context cache (embeddings + cross-attention K / V computed once per loop) == the uncached loop, bit for bit."""
    g = _golden("wan_tiny.pt")
    cfg, dn = g["cfg"], g["denoise"]
    model = build_wan(cfg, g["weight_seed"])
    f, h, w = g["latent"]
    inp = synth.wan_inputs(cfg, f, h, w, seed=g["input_seed"])
    neg = synth.wan_inputs(cfg, f, h, w, seed=dn["neg_seed"])
    gen = torch.Generator().manual_seed(dn["seed"])
    lat0 = torch.randn((1, 16, f, h, w), generator=gen)
    lat_ref = torch.randn((1, 16, f, h, w), generator=gen)
    kw = _to_dev({k: inp[k] for k in ("encoder_hidden_states", "encoder_hidden_states_image", "encoder_hidden_states_mot_ref",
                                      "encoder_hidden_states_image_mot_ref", "num_mot_ref")})
    kw_u = dict(kw, encoder_hidden_states=neg["encoder_hidden_states"].to(DEV), encoder_hidden_states_mot_ref=neg["encoder_hidden_states_mot_ref"].to(DEV))
    outs = []
    with torch.no_grad():
        for cache in (False, True):
            outs.append(vap.denoise.wan_denoise(model, lat0.to(DEV), inp["hidden_states"][:, 16:].float().to(DEV), lat_ref.to(DEV),
                                                inp["hidden_states_mot_ref"][:, 16:].float().to(DEV), kw, kw_u, dn["steps"], dn["shift"], dn["guidance"],
                                                cache_context=cache))
        lat_b2 = vap.denoise.wan_denoise(model, lat0.to(DEV), inp["hidden_states"][:, 16:].float().to(DEV), lat_ref.to(DEV),
                                         inp["hidden_states_mot_ref"][:, 16:].float().to(DEV), kw, kw_u, dn["steps"], dn["shift"], dn["guidance"],
                                         cache_context=True, batch_cfg=True)
    assert torch.equal(outs[0], outs[1]), f"cached loop differs: rel err {rel_err(outs[1], outs[0])}"
    cos = cosine(outs[1], dn["final_latents"])
    assert cos >= 0.999
    # the B = 2 forward runs the same kernels per sample (per-batch launches or batch-strided launches): expected bit-exact as well
    err_b2, cos_b2 = rel_err(lat_b2, outs[0]), cosine(lat_b2, dn["final_latents"])
    assert cos_b2 >= 0.999 and err_b2 < 2e-2, f"batched-CFG loop: cosine {cos_b2}, rel err vs the sequential loop {err_b2}"
    return dict(cosine=cos, bit_exact=True, batch_cfg_bit_exact=bool(torch.equal(lat_b2, outs[0])), batch_cfg_err=err_b2, batch_cfg_cosine=cos_b2)


def check_wan_dead_ref_skip(tol=5e-3):
    """WanTransformer3DMOTModel.skip_dead_reference_work: the last MoT block drops the expert stream's query rows, O-projection,
    cross-attention and FFN.  Same model output up to the attention kernel's warp-level rescale vote in the one 256-row block that used to
    hold both target and reference rows (the target rows' arithmetic is otherwise identical), and still within the gate of the reference."""
    g = _golden("wan_tiny.pt")
    cfg = g["cfg"]
    model = build_wan(cfg, g["weight_seed"])
    inp = _to_dev(synth.wan_inputs(cfg, *g["latent"], seed=g["input_seed"]))
    with torch.no_grad():
        full = model(**inp, return_dict=False)[0]
        model.skip_dead_reference_work = True
        lean = model(**inp, return_dict=False)[0]
        model.skip_dead_reference_work = False
    err, err_ref = rel_err(lean, full), rel_err(lean, g["final"])
    assert err <= tol and err_ref <= 2e-2, f"dead-work skip: rel err vs the full path {err}, vs the reference {err_ref}"
    return dict(err=err, err_ref=err_ref, bit_exact=bool(torch.equal(lean, full)))


CHECKS = {
    "probe_ss": lambda: check_probe(False, False, 128, 128),
    "probe_ss_n256": lambda: check_probe(False, False, 256, 64),
    "probe_mn": lambda: check_probe(False, True, 128, 128),
    "probe_mn_n64": lambda: check_probe(False, True, 64, 128),
    "probe_ts": lambda: check_probe(True, False, 128, 128),
    "probe_ts_mn": lambda: check_probe(True, True, 128, 128),
    "probe_lane16_ld": lambda: check_probe(False, False, 128, 128, lane16_shapes=True),
    "probe_lane16_st_ld": lambda: check_probe(True, True, 128, 128, lane16_shapes=True),
    "ln_wan": lambda: check_layernorm_wan(),
    "ln_wan_affine": lambda: check_layernorm_wan(rows=100, d=256, affine=True, modulate=False),
    "ln_wan_batch": lambda: check_layernorm_wan(rows=64, d=3072, batch=2),
    "ln_cog": lambda: check_layernorm_cog(),
    "wan_modulation": check_wan_modulation,
    "qk_wan": lambda: check_qk_wan(),
    # >= 1024 rows at d = 5120 / 3072: the staged kernels (bulk-copy ring per warp, per-channel vectors in shared memory); ragged row counts
    "ln_wan_staged": lambda: check_layernorm_wan(rows=1531, d=5120),
    "ln_wan_staged_affine": lambda: check_layernorm_wan(rows=1100, d=5120, affine=True, modulate=False),
    "ln_wan_staged_batch2": lambda: check_layernorm_wan(rows=1033, d=5120, batch=2),
    "ln_cog_staged": lambda: check_layernorm_cog(rows=1203, d=3072, nmod=2),
    "qk_wan_staged": lambda: check_qk_wan(S=1200, H=40, D=128, grid=(3, 20, 20)),
    "qk_wan_tiny": lambda: check_qk_wan(S=300, H=2, D=128),
    "qk_cog": lambda: check_qk_cog(),
    "gemm_small": lambda: check_gemm(300, 512, 256, 0),
    "gemm_n128": lambda: check_gemm(129, 128, 64, 0),
    "gemm_tails": lambda: check_gemm(1000, 776, 328, 0),
    "gemm_gelu": lambda: check_gemm(257, 1024, 512, 1),
    "gemm_gate_f32": lambda: check_gemm(512, 256, 512, 2, nbatch=2),
    "gemm_res_add": lambda: check_gemm(300, 256, 256, 3),
    "gemm_gate_bf16": lambda: check_gemm(452, 256, 256, 4, nbatch=2),
    "gemm_nobias_k5120": lambda: check_gemm(640, 768, 5120, 0, bias=False),
    "attn_d128": lambda: check_attention(1, 2, 300, 300, 128),
    "attn_d128_multi_tile": lambda: check_attention(1, 3, 1000, 1000, 128),
    "attn_d64": lambda: check_attention(2, 4, 452, 452, 64),
    "attn_cross": lambda: check_attention(1, 2, 600, 257, 128, joint_layout=False),
    "attn_cross_512": lambda: check_attention(1, 2, 300, 512, 128, joint_layout=False),
    "attn_one_tile": lambda: check_attention(1, 1, 64, 100, 128, joint_layout=False),
    "attn_peaky": lambda: check_attention_peaky(),
    "attn_variants_row_pair": check_attention_variants,
    "attn_short_and_long_forced": check_attention_short_and_long,
    "attn_accumulate": lambda: check_attention_accumulate(),
    "attn_accumulate_d64": lambda: check_attention_accumulate(D=64, H=2, Lq=300, kv=(100, 226)),
    "attn_splitkv_2": lambda: check_attention_splitkv(1, 2, 300, 1000, 128, 2),
    "attn_splitkv_3_d64": lambda: check_attention_splitkv(2, 3, 452, 900, 64, 3),
    "attn_splitkv_uneven": lambda: check_attention_splitkv(1, 1, 130, 128 * 5 + 7, 128, 4),
    "attn_splitkv_peers": lambda: check_attention_splitkv_peers(),
    "attn_splitkv_peers_p8": lambda: check_attention_splitkv_peers(P=8, L=40, H=5, D=128, splits=2),
    "ulysses_relayout": check_ulysses_relayout,
    "ulysses_p2p_emulated_wan": lambda: check_ulysses_p2p_emulated(P=4, L=96, H=8, D=128, mode=0),
    "ulysses_p2p_emulated_p8": lambda: check_ulysses_p2p_emulated(P=8, L=300, H=40, D=128, mode=0),
    "ulysses_p2p_emulated_splitkv": check_ulysses_p2p_emulated_splitkv,
    "ulysses_p2p_emulated_cog": lambda: check_ulysses_p2p_emulated(P=2, L=130, H=6, D=64, mode=1),
    "wan_blocks": check_wan_blocks,
    "wan_model": check_wan_model,
    "wan_denoise": check_wan_denoise,
    "cog_blocks_small": lambda: check_cog_blocks("small"),
    "cog_blocks_multi": lambda: check_cog_blocks("multi"),
    "cog_model_config1": lambda: check_cog_model("config1"),
    "cog_denoise": check_cog_denoise,
    "processor_level": check_processor_level,
    "gemm_large": check_gemm_large,
    "attn_full_size": check_attention_full_size,
    "attn_splitkv_sp8_shape": check_attention_splitkv_sp8_shape,
    # first run on a B200 in round 2 (profiles/r02_pending_and_reference_checks.json): the attention backward (dQ / dK / dV on tcgen05 from the
    # forward's LSE, also through the differentiable joint_sdpa seam), the fused CFG + FlowMatchEuler step, the cached / batched-CFG / fused denoise loops
    "attn_bwd_one_tile": lambda: check_attention_bwd(1, 1, 100, 100, 128, joint_layout=False),
    "attn_bwd_d128": lambda: check_attention_bwd(1, 2, 300, 300, 128),
    "attn_bwd_d64": lambda: check_attention_bwd(2, 3, 452, 260, 64, joint_layout=False),
    "attn_bwd_tails": lambda: check_attention_bwd(1, 1, 130, 128 * 5 + 7, 128, joint_layout=False),
    "attn_bwd_multi_tile": lambda: check_attention_bwd(1, 2, 1000, 1000, 128),
    "attn_bwd_split1_d128": lambda: check_attention_bwd_split1(B=1, H=2, Lq=300, Lkv=647, D=128, joint_layout=False),
    "attn_bwd_split1_d64": lambda: check_attention_bwd_split1(B=2, H=3, Lq=452, Lkv=260, D=64, joint_layout=False),
    "attn_bwd_noprefetch_d128": lambda: check_attention_bwd_noprefetch(B=1, H=2, Lq=300, Lkv=647, D=128, joint_layout=False),
    "attn_bwd_noprefetch_d64": lambda: check_attention_bwd_noprefetch(B=1, H=2, Lq=452, Lkv=260, D=64, joint_layout=False),
    "cfg_flow_match_step": check_cfg_flow_match_step,
    "wan_denoise_fused": check_wan_denoise_fused,
    "wan_denoise_cached": check_wan_denoise_cached,
    "wan_dead_ref_skip": check_wan_dead_ref_skip,
}

# Checks of code that has not been on a GPU yet: run by `tools/gpu_diag.py --pending` in a development call, promoted into CHECKS (and so into
# `pytest -m gpu`) once they have passed on a B200.  Empty: everything written in round 1 without GPU access passed its first hardware run.
CHECKS_PENDING = {}
