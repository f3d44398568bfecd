"""TEST INFRASTRUCTURE ONLY: torch-on-CPU stand-ins with the call signatures of video-as-prompt_b200/ops.py, so that the HOST logic
around the kernels — block forwards, transformer shells, Ulysses sharding / exchange / gather, zero-row guards — can run in the
CPU suite (there is no GPU in the authoring container, and the product path has no CPU fallback by design).

`install(vap)` monkeypatches the ops module of an imported package inside ONE test process; nothing in the product imports this
file.  The arithmetic follows the kernels' contracts in include/vap_b200.h (rounding points included) built from the oracle's
primitives, but it is the sm_100a kernels — not this file — that the `-m gpu` parity tests check against the oracle."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from oracle import common as oc

BF16 = torch.bfloat16


def _rows(t: torch.Tensor) -> torch.Tensor:
    return t.reshape(-1, t.shape[-1])


def _per_row(vec: Optional[torch.Tensor], rows: int, rows_per_batch: Optional[int], lead_rows: int) -> Optional[torch.Tensor]:
    """Broadcast per-batch vectors [nb, d] / [nb, 1, d] to one vector per row (row r takes batch r // rows_per_batch)."""
    if vec is None:
        return None
    v = vec.reshape(-1, vec.shape[-1]).float()
    if v.shape[0] == 1:
        return v.expand(rows, -1)
    rpb = rows_per_batch if rows_per_batch is not None else lead_rows
    idx = torch.arange(rows) // max(int(rpb), 1)
    return v[idx]


def adaln_layernorm(x, *, eps, rounding, ln_w=None, ln_b=None, scale1p=None, shift=None, rows_per_batch=None, out=None):
    xr = _rows(x)
    rows, d = xr.shape
    lead = x.shape[-2] if x.dim() >= 3 else rows
    y = F.layer_norm(xr.float(), (d,), ln_w, ln_b, eps)
    s1p, sh = _per_row(scale1p, rows, rows_per_batch, lead), _per_row(shift, rows, rows_per_batch, lead)
    if rounding == 1:  # CogVideoX: bf16 tensor arithmetic
        y = y.to(BF16).float()
        if s1p is not None:
            y = (y * s1p).to(BF16).float()
        if sh is not None:
            y = (y + sh).to(BF16).float()
    else:
        if s1p is not None:
            y = y * s1p
        if sh is not None:
            y = y + sh
    y = y.to(BF16).reshape(x.shape)
    if out is None:
        return y
    out.copy_(y.reshape(out.shape))
    return out


def qk_norm_rope_(q, k, *, heads, head_dim, wq, wk=None, bq=None, bk=None, cos=None, sin=None, rows_per_batch, rope_row0=0, eps, mode):
    for t, w, b in ((q, wq, bq), (k, wk, bk)):
        if t is None:
            continue
        x = _rows(t)
        rows = x.shape[0]
        if mode == 0:  # Wan: RMSNorm across all heads, rounded to bf16 before the weight multiply
            y = (oc.rms_norm_across(x, w.to(BF16), eps)).float()
        else:          # CogVideoX: per-head LayerNorm(head_dim), affine
            y = F.layer_norm(x.float().view(rows, heads, head_dim), (head_dim,), w, b, eps).to(BF16).float().view(rows, heads * head_dim)
        if cos is not None and cos.shape[0] > 0:
            tpos = torch.arange(rows) % int(rows_per_batch) - int(rope_row0)
            rot = tpos >= 0
            c = cos[tpos.clamp(min=0)][:, None, :]  # [rows, 1, D/2]
            s = sin[tpos.clamp(min=0)][:, None, :]
            yh = y.view(rows, heads, head_dim // 2, 2)
            a, b2 = yh[..., 0], yh[..., 1]
            ra, rb = a * c - b2 * s, a * s + b2 * c
            yr = torch.stack([ra, rb], dim=-1).view(rows, heads * head_dim)
            y = torch.where(rot[:, None], yr, y)
        t.copy_(y.to(BF16).reshape(t.shape))


def attention(q, k, v, *, scale=None, out=None, return_lse=False, accumulate=False):
    B, H, Lq, D = q.shape
    if scale is None or scale == D ** -0.5:
        o = oc.sdpa_explicit_fp32(q, k, v).to(BF16)  # [B, H, Lq, D]
    else:  # a caller-chosen softmax scale (the oracle helper fixes it at D^-1/2)
        o = torch.matmul(torch.softmax(torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale, dim=-1), v.float()).to(BF16)
    res = torch.empty((B, Lq, H, D), dtype=BF16).transpose(1, 2)  # token-major memory like the kernel's output
    res.copy_(o)
    if out is not None:
        out.copy_(out + res if accumulate else res)  # accumulate: the bf16 tensor add of vap_attention_fwd_accumulate
        res = out
    if return_lse:
        s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * (scale or D ** -0.5)
        return res, torch.logsumexp(s, dim=-1)
    return res


def attention_bwd(q, k, v, o, lse, dout, *, scale=None):
    """vap_attention_bwd's arithmetic (attn_bwd_sm100.cu): P recomputed from the stored log-sum-exp, delta = rowsum(dO o O),
    P and dS rounded to bf16 before the second round of products (they are tensor-core operands), fp32 accumulation."""
    B, H, Lq, D = q.shape
    sc = scale or D ** -0.5
    qf, kf, vf, of, gf = (t.float() for t in (q, k, v, o, dout))
    p = torch.exp(torch.matmul(qf, kf.transpose(-1, -2)) * sc - lse.unsqueeze(-1))
    delta = (gf * of).sum(-1, keepdim=True)
    ds = (p * (torch.matmul(gf, vf.transpose(-1, -2)) - delta) * sc).to(BF16).float()
    pb = p.to(BF16).float()
    outs = (torch.matmul(ds, kf), torch.matmul(ds.transpose(-1, -2), qf), torch.matmul(pb.transpose(-1, -2), gf))
    res = []
    for g in outs:
        r = torch.empty((B, g.shape[2], H, D), dtype=BF16).transpose(1, 2)
        r.copy_(g.to(BF16))
        res.append(r)
    return tuple(res)


def linear(x, weight, bias=None, *, epilogue=0, residual=None, gate=None, rows_per_batch=None, out=None):
    xr = _rows(x)
    M = xr.shape[0]
    y = F.linear(xr.float(), weight.float(), bias.float() if bias is not None else None).to(BF16)
    if epilogue == 1:
        y = oc.gelu_tanh(y.float()).to(BF16)
    elif epilogue in (2, 3, 4):
        r = _rows(residual).float()
        if epilogue == 3:
            y = (r + y.float()).to(BF16)
        else:
            lead = x.shape[-2] if x.dim() >= 3 else M
            g = _per_row(gate, M, rows_per_batch, lead)
            y = (r + y.float() * g).to(BF16) if epilogue == 2 else (r + (g * y.float()).to(BF16).float()).to(BF16)
    y = y.reshape(x.shape[:-1] + (weight.shape[0],))
    if out is None:
        return y
    out.copy_(y.reshape(out.shape))
    return out


def ulysses_pack(src, nsplit, out):
    L, width = src.shape
    out.copy_(src.reshape(L, nsplit, width // nsplit).permute(1, 0, 2))
    return out


def ulysses_unpack(src, out=None):
    n, L, c = src.shape
    res = src.permute(1, 0, 2).reshape(L, n * c)
    if out is None:
        return res.contiguous()
    out.copy_(res)
    return out


def cfg_flow_match_step(noise_cond, noise_uncond, sample, *, guidance_scale, dt, out=None):
    """vap_cfg_flow_match_step's contract (include/vap_b200.h (8)): every reference tensor op rounds once."""
    n = noise_cond.float()
    if noise_uncond is not None:
        u = noise_uncond.float()
        d = (n - u).to(BF16).float()
        m = (torch.tensor(guidance_scale, dtype=torch.float32) * d).to(BF16).float()
        n = (u + m).to(BF16).float()
    y = (sample.float() + (torch.tensor(dt, dtype=torch.float32) * n).to(BF16).float()).to(BF16)
    if out is None:
        return y
    out.copy_(y)
    return out


def wan_modulation(table, temb, plus_one_mask=0b010010):
    """vap_wan_modulation: (table + temb.float()) with 1 added to the masked chunks, fp32 (transformer_wan_mot.py:606-608, 620-622)."""
    mod = table.float() + temb.float()
    for c in range(mod.shape[1]):
        if (plus_one_mask >> c) & 1:
            mod[:, c] += 1
    return mod


def install(vap) -> None:
    """Replace the kernel wrappers of `vap.ops` by the stand-ins (one test process only)."""
    for name in ("adaln_layernorm", "qk_norm_rope_", "attention", "attention_bwd", "linear", "ulysses_pack", "ulysses_unpack", "cfg_flow_match_step", "wan_modulation"):
        setattr(vap.ops, name, globals()[name])
    vap.ops.sm_count = lambda: 148
