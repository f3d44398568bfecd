"""Parameter containers that mirror the reference's module tree (same attribute names => same state_dict keys).

They exist so the B200 path can be built and run stand-alone (the GPU box has no copy of the reference) while every
``state_dict`` stays interchangeable with the reference's ``WanTransformer3DMOTModel`` /
``CogVideoXTransformer3DMOTModel``.  Only the weights live here; the arithmetic of the hot path is in the fused block
forwards (``wan.py`` / ``cogvideox.py``) which call the sm_100a kernels through ``ops``.

Reference (paths relative to /root/reference/diffusers/src/diffusers):
  Attention            models/attention_processor.py:49-306 (init), :532-550 (set_processor), :566-610 (forward)
  RMSNorm              models/normalization.py:511-568
  FP32LayerNorm        models/normalization.py:85-94
  FeedForward / GELU   models/attention.py:1191-1251, models/activations.py:65-91
  CogVideoXLayerNormZero models/normalization.py:449-471
"""
from __future__ import annotations

import inspect
import logging
from typing import Optional

import torch
import torch.nn as nn

logger = logging.getLogger("vap_b200")


class RMSNorm(nn.Module):
    """Weight container of diffusers' RMSNorm(dim, eps, elementwise_affine=True)."""

    def __init__(self, dim: int, eps: float):
        super().__init__()
        self.eps = eps
        self.dim = torch.Size((dim,))
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = None


class FP32LayerNorm(nn.LayerNorm):
    """nn.LayerNorm whose reference forward runs in fp32 (normalization.py:85-94); weights only here."""


class GELU(nn.Module):
    """`proj` Linear + GELU(tanh) of FeedForward.net[0] (activations.py:65-91); weights only."""

    def __init__(self, dim_in: int, dim_out: int, approximate: str = "tanh", bias: bool = True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate


class FeedForward(nn.Module):
    """net = [GELU(proj), Dropout, Linear(, Dropout)] exactly like attention.py:1191-1251 so keys are net.0.proj.* / net.2.*."""

    def __init__(self, dim: int, dim_out: Optional[int] = None, mult: int = 4, inner_dim: Optional[int] = None,
                 activation_fn: str = "gelu-approximate", final_dropout: bool = False, bias: bool = True):
        super().__init__()
        inner_dim = int(dim * mult) if inner_dim is None else inner_dim
        dim_out = dim if dim_out is None else dim_out
        if activation_fn == "gelu-approximate":
            act = GELU(dim, inner_dim, approximate="tanh", bias=bias)
        elif activation_fn == "gelu":
            act = GELU(dim, inner_dim, approximate="none", bias=bias)
        else:
            raise ValueError(f"unsupported activation_fn {activation_fn!r} on the VAP path")
        self.net = nn.ModuleList([act, nn.Dropout(0.0), nn.Linear(inner_dim, dim_out, bias=bias)])
        if final_dropout:
            self.net.append(nn.Dropout(0.0))


class CogVideoXLayerNormZero(nn.Module):
    def __init__(self, conditioning_dim: int, embedding_dim: int, elementwise_affine: bool = True, eps: float = 1e-5, bias: bool = True):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(conditioning_dim, 6 * embedding_dim, bias=bias)
        self.norm = nn.LayerNorm(embedding_dim, eps=eps, elementwise_affine=elementwise_affine)


class AdaLayerNorm(nn.Module):
    """AdaLayerNorm(chunk_dim=1) used as CogVideoX norm_out (normalization.py:28-82); weights only."""

    def __init__(self, embedding_dim: int, output_dim: int, norm_elementwise_affine: bool, norm_eps: float):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(embedding_dim, output_dim)
        self.norm = nn.LayerNorm(output_dim // 2, norm_eps, norm_elementwise_affine)


class Attention(nn.Module):
    """Weights container + processor plug point with the reference's attribute names and call contract.

    ``forward`` keeps the reference's behaviour of dropping (with a warning) every kwarg that the processor's
    ``__call__`` does not declare (attention_processor.py:593-602)."""

    def __init__(self, query_dim: int, heads: int, dim_head: int, qk_norm: Optional[str], eps: float = 1e-5, bias: bool = False,
                 out_bias: bool = True, added_kv_proj_dim: Optional[int] = None, added_proj_bias: bool = True, processor=None):
        super().__init__()
        self.inner_dim = heads * dim_head
        self.query_dim = query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.is_cross_attention = False
        self.added_kv_proj_dim = added_kv_proj_dim
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(0.0)])
        if qk_norm is None:
            self.norm_q = self.norm_k = None
        elif qk_norm == "layer_norm":
            self.norm_q = nn.LayerNorm(dim_head, eps=eps, elementwise_affine=True)
            self.norm_k = nn.LayerNorm(dim_head, eps=eps, elementwise_affine=True)
        elif qk_norm == "rms_norm_across_heads":
            self.norm_q = RMSNorm(dim_head * heads, eps=eps)
            self.norm_k = RMSNorm(dim_head * heads, eps=eps)
        else:
            raise ValueError(f"unsupported qk_norm {qk_norm!r} on the VAP path")
        self.add_k_proj = self.add_v_proj = self.norm_added_k = None
        if added_kv_proj_dim is not None:
            self.add_k_proj = nn.Linear(added_kv_proj_dim, self.inner_dim, bias=added_proj_bias)
            self.add_v_proj = nn.Linear(added_kv_proj_dim, self.inner_dim, bias=added_proj_bias)
            if qk_norm == "rms_norm_across_heads":
                self.norm_added_k = RMSNorm(dim_head * heads, eps=eps)
        self.processor = processor

    def set_processor(self, processor) -> None:
        self.processor = processor

    def get_processor(self):
        return self.processor

    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                attention_mask: Optional[torch.Tensor] = None, **cross_attention_kwargs):
        params = set(inspect.signature(self.processor.__call__).parameters.keys())
        unused = [k for k in cross_attention_kwargs if k not in params and k not in {"ip_adapter_masks", "ip_hidden_states"}]
        if unused:
            logger.warning(f"cross_attention_kwargs {unused} are not expected by {self.processor.__class__.__name__} and will be ignored.")
        kwargs = {k: w for k, w in cross_attention_kwargs.items() if k in params}
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states, attention_mask=attention_mask, **kwargs)


class TimestepEmbedding(nn.Module):
    """linear_1 -> SiLU -> linear_2 (embeddings.py:1307-1352); transformer-shell glue, runs in torch."""

    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample: torch.Tensor) -> torch.Tensor:
        return self.linear_2(self.act(self.linear_1(sample)))


class PixArtAlphaTextProjection(nn.Module):
    """linear_1 -> GELU(tanh) -> linear_2 (embeddings.py:2237-2264); shell glue."""

    def __init__(self, in_features: int, hidden_size: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, hidden_size)
        self.act_1 = nn.GELU(approximate="tanh")
        self.linear_2 = nn.Linear(hidden_size, hidden_size)

    def forward(self, caption: torch.Tensor) -> torch.Tensor:
        return self.linear_2(self.act_1(self.linear_1(caption)))
