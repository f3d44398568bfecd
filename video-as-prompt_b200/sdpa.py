"""Boundary B2: a drop-in for the `F.scaled_dot_product_attention` slot the MoT blocks call
(transformer_wan_mot.py:637-644, cogvideox_transformer_3d_mot.py:424-431), with the signature of finetrainers'
`attention_dispatch` (finetrainers/models/attention_dispatch.py:416-458) so it can be installed the same way
finetrainers installs its own dispatcher (finetrainers/patches/__init__.py:66-69)."""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
import torch.nn.functional as F

from . import ops

_ORIGINAL_SDPA = None
_STRICT = True  # False: calls outside the kernel's envelope go to the ORIGINAL torch SDPA (see patch_scaled_dot_product_attention)


class _JointAttention(torch.autograd.Function):
    """Differentiable joint attention for the trainer's B2 seam: forward = vap_attention_fwd (keeps O and the log-sum-exp),
    backward = vap_attention_bwd (dQ / dK / dV within 2-5e-3 of fp32 autograd on a B200, tests/gpu_checks.py attn_bwd_*)."""

    @staticmethod
    def forward(ctx, q, k, v, scale):
        o, lse = ops.attention(q, k, v, scale=scale, return_lse=True)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.scale = scale
        return o

    @staticmethod
    def backward(ctx, dout):
        q, k, v, o, lse = ctx.saved_tensors
        # autograd may hand over an expanded (stride 0) or oddly strided gradient, e.g. after o.sum(): the kernel wants D contiguous and
        # (batch, head, token) strides that are positive multiples of 8 elements
        if dout.stride(-1) != 1 or any(n > 1 and (st <= 0 or st % 8) for n, st in zip(dout.shape[:3], dout.stride()[:3])) or dout.data_ptr() % 16:
            dout = dout.contiguous()
        dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, dout, scale=ctx.scale)
        return dq, dk, dv, None


def _unsupported(query, key, value, attn_mask, dropout_p, is_causal, enable_gqa) -> Optional[str]:
    """Why this call is outside the joint-attention kernel's envelope (None if it is inside)."""
    if attn_mask is not None:
        return "attn_mask is not supported (the MoT joint attention is unmasked)"
    if dropout_p != 0.0:
        return "dropout_p must be 0.0"
    if is_causal:
        return "is_causal=True is not supported"
    if query.dim() != 4 or key.dim() != 4 or value.dim() != 4:
        return f"q/k/v must be [B,H,L,D], got {query.dim()}-d"
    if enable_gqa or key.shape[1] != query.shape[1]:
        return "grouped-query attention is not supported"
    if query.dtype != torch.bfloat16 or key.dtype != torch.bfloat16 or value.dtype != torch.bfloat16:
        return f"q/k/v must be bfloat16, got {query.dtype}/{key.dtype}/{value.dtype}"
    if query.shape[-1] not in (64, 128):
        return f"head_dim {query.shape[-1]} not in (64, 128)"
    return None


def joint_sdpa(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, attn_mask: Optional[torch.Tensor] = None,
               dropout_p: float = 0.0, is_causal: bool = False, scale: Optional[float] = None, enable_gqa: bool = False,
               attention_kwargs: Optional[Dict[str, Any]] = None) -> torch.Tensor:
    """[B,H,L,D] in, [B,H,L,D] out, same dtype.  Constraint violations raise ValueError like attention_dispatch's
    provider checks (attention_dispatch.py:471-530).  Only under patch_scaled_dot_product_attention(strict=False) are calls OUTSIDE the kernel's
    envelope passed to the original torch function (they are not MoT attention); nothing inside the envelope ever falls back."""
    why = _unsupported(query, key, value, attn_mask, dropout_p, is_causal, enable_gqa)
    if why is not None:
        if not _STRICT and _ORIGINAL_SDPA is not None:  # e.g. the VAE's or the text encoder's attention under a global patch
            kw = {} if not enable_gqa else {"enable_gqa": True}
            return _ORIGINAL_SDPA(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale, **kw)
        raise ValueError("joint_sdpa: " + why)
    q, k, v = (t if t.stride(-1) == 1 else t.contiguous() for t in (query, key, value))
    if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad):
        return _JointAttention.apply(q, k, v, scale)  # trainer path: O keeps a grad_fn
    return ops.attention(q, k, v, scale=scale)


def patch_scaled_dot_product_attention(strict: bool = True) -> None:
    """Install `joint_sdpa` as torch.nn.functional.scaled_dot_product_attention (mirrors finetrainers/patches/__init__.py:66-69).
    strict=True (default): a call the kernel does not cover raises ValueError, like attention_dispatch's constraint checks.
    strict=False: such calls (masks, dropout, causal, other head dims / dtypes — the VAE's and the text encoders' attention) are handed to the
    ORIGINAL torch function unchanged.  A call INSIDE the envelope always runs the sm_100a kernel — on CPU tensors it raises VapError, strict or not."""
    global _ORIGINAL_SDPA, _STRICT
    if _ORIGINAL_SDPA is None:
        _ORIGINAL_SDPA = F.scaled_dot_product_attention
    _STRICT = bool(strict)
    F.scaled_dot_product_attention = joint_sdpa


def unpatch_scaled_dot_product_attention() -> None:
    global _ORIGINAL_SDPA, _STRICT
    if _ORIGINAL_SDPA is not None:
        F.scaled_dot_product_attention = _ORIGINAL_SDPA
        _ORIGINAL_SDPA = None
    _STRICT = True
