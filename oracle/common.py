"""Shared arithmetic primitives of the oracle (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Every function restates one reference building block with its dtype flow kept.
Paths are relative to /root/reference/diffusers/src/diffusers.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def linear(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """nn.Linear with the reference's parameter names (``<name>.weight`` / ``<name>.bias``)."""
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def fp32_layer_norm(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    """FP32LayerNorm.forward, models/normalization.py:85-94: normalise in fp32, cast back to the input dtype."""
    d = x.shape[-1]
    return F.layer_norm(
        x.float(), (d,), weight.float() if weight is not None else None, bias.float() if bias is not None else None, eps
    ).to(x.dtype)


def rms_norm_across(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """RMSNorm.forward (non-NPU branch), models/normalization.py:554-568.

    variance in fp32 over the whole last dim (all heads: qk_norm="rms_norm_across_heads",
    models/attention_processor.py:207-210); x(bf16)*rsqrt(fp32) -> fp32; rounded to the weight
    dtype BEFORE the multiply by the (bf16) weight.
    """
    variance = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    y = x * torch.rsqrt(variance + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        y = y.to(weight.dtype)
    return y * weight


def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    """GELU(approximate="tanh"), models/activations.py:65-91."""
    return F.gelu(x, approximate="tanh")


def feed_forward(sd: SD, prefix: str, x: torch.Tensor, approximate: str = "tanh") -> torch.Tensor:
    """FeedForward(activation_fn="gelu-approximate"), models/attention.py:1191-1251:
    net.0.proj Linear -> gelu -> (Dropout 0) -> net.2 Linear."""
    h = linear(sd, prefix + ".net.0.proj", x)
    h = F.gelu(h, approximate=approximate)
    return linear(sd, prefix + ".net.2", h)


def timestep_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos: bool = True, downscale_freq_shift: float = 0.0,
                       max_period: int = 10000) -> torch.Tensor:
    """get_timestep_embedding, models/embeddings.py:25-77 (scale=1)."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


def sdpa(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """F.scaled_dot_product_attention(attn_mask=None, dropout_p=0, is_causal=False), default scale D^-1/2
    (transformer_wan_mot.py:637-644, cogvideox_transformer_3d_mot.py:424-431)."""
    return F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)


def sdpa_explicit_fp32(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Definition-level attention in fp32 (softmax(QK^T/sqrt(D))V), used to cross-check `sdpa` and as
    the dense reference for the joint-attention kernel tests."""
    scale = q.shape[-1] ** -0.5
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, v.float())
