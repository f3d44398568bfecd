"""Oracle: CogVideoX-VAP MoT transformer (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Functional restatement of models/transformers/cogvideox_transformer_3d_mot.py (MoT branch :375-513,
plain branch :171-203, shell :886-1106) on a flat state_dict, keeping the reference's all-bf16
arithmetic (so it is bit-identical to the reference module on CPU with the same torch build).
Single-reference path (`temb_mot_ref`, `timestep_list_mot_ref is None`) and the multi-reference
per-ref-timestep path (`temb_list_mot_ref`) are both restated.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .common import SD, feed_forward, linear, sdpa, timestep_embedding


# ----------------------------------------------------------------------------------------------
# RoPE tables
# ----------------------------------------------------------------------------------------------
def _rope_1d_real(dim: int, pos: torch.Tensor, theta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """get_1d_rotary_pos_embed(use_real=True, repeat_interleave_real=True), models/embeddings.py:1181-1196."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    freqs = torch.outer(pos, freqs)
    cos = freqs.cos().repeat_interleave(2, dim=1).float()
    sin = freqs.sin().repeat_interleave(2, dim=1).float()
    return cos, sin


def cog_rope_3d(embed_dim: int, crops_coords, grid_size, temporal_size: int, theta: float = 10000.0, mot_num: int = 0,
                ref_type: str = "continous_negative", start_point: int = 50, gap: int = 30):
    """get_3d_rotary_pos_embed(grid_type="linspace"), models/embeddings.py:816-949.
    mot_num > 0 selects the reference-video temporal positions: "continous_negative" ->
    linspace(-mot_num*t_range, -1, mot_num*T) (:871-881); "discrete_long_reference" -> 50+30*i+arange(T) (:886-890)."""
    start, stop = crops_coords
    gh, gw = grid_size
    grid_h = torch.linspace(start[0], stop[0] * (gh - 1) / gh, gh, dtype=torch.float32)
    grid_w = torch.linspace(start[1], stop[1] * (gw - 1) / gw, gw, dtype=torch.float32)
    grid_t = torch.linspace(0, temporal_size * (temporal_size - 1) / temporal_size, temporal_size, dtype=torch.float32)
    if mot_num > 0:
        if ref_type == "continous_negative":
            t_range = temporal_size * (temporal_size - 1) / temporal_size - 0 + 1
            temporal_size = temporal_size * mot_num
            grid_t = torch.linspace(-mot_num * t_range, -1, temporal_size, dtype=torch.float32)
        elif ref_type == "discrete_long_reference":
            offs = start_point + torch.arange(mot_num, dtype=torch.float32) * gap
            grid_t = (offs.unsqueeze(1) + torch.arange(temporal_size, dtype=torch.float32)).flatten()
            # NOTE: the reference does not rescale temporal_size here (:886-890), so the table only
            # broadcasts when mot_num == 1; kept as is.
        else:
            raise ValueError(f"Invalid {ref_type} passed for `ref_type`.")
    dim_t, dim_h, dim_w = embed_dim // 4, embed_dim // 8 * 3, embed_dim // 8 * 3
    t_cos, t_sin = _rope_1d_real(dim_t, grid_t, theta)
    h_cos, h_sin = _rope_1d_real(dim_h, grid_h, theta)
    w_cos, w_sin = _rope_1d_real(dim_w, grid_w, theta)

    def combine(ft, fh, fw):
        ft = ft[:, None, None, :].expand(-1, gh, gw, -1)
        fh = fh[None, :, None, :].expand(temporal_size, -1, gw, -1)
        fw = fw[None, None, :, :].expand(temporal_size, gh, -1, -1)
        return torch.cat([ft, fh, fw], dim=-1).reshape(temporal_size * gh * gw, -1)

    return combine(t_cos, h_cos, w_cos), combine(t_sin, h_sin, w_sin)


def apply_rope_real(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """apply_rotary_emb(use_real=True, use_real_unbind_dim=-1), models/embeddings.py:1229-1248:
    interleaved pairs (x[2i], x[2i+1]) -> (-x[2i+1], x[2i]); fp32 math; back to x.dtype."""
    cos, sin = cos[None, None], sin[None, None]
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return (x.float() * cos + rot.float() * sin).to(x.dtype)


# ----------------------------------------------------------------------------------------------
# Block pieces
# ----------------------------------------------------------------------------------------------
def layer_norm_zero(sd: SD, p: str, v: torch.Tensor, e: torch.Tensor, temb: torch.Tensor, eps: float):
    """CogVideoXLayerNormZero.forward, models/normalization.py:464-471 (all in the model dtype)."""
    shift, scale, gate, e_shift, e_scale, e_gate = linear(sd, p + ".linear", F.silu(temb)).chunk(6, dim=1)
    d = v.shape[-1]
    w, b = sd[p + ".norm.weight"], sd[p + ".norm.bias"]
    vn = F.layer_norm(v, (d,), w, b, eps) * (1 + scale)[:, None, :] + shift[:, None, :]
    en = F.layer_norm(e, (d,), w, b, eps) * (1 + e_scale)[:, None, :] + e_shift[:, None, :]
    return vn, en, gate[:, None, :], e_gate[:, None, :]


def attn_pre(sd: SD, p: str, vn: torch.Tensor, en: torch.Tensor, heads: int, rope):
    """CogVideoXAttnMOTProcessor2_0.__call__(is_before_attn=True), models/attention_processor.py:2912-2946."""
    T = en.shape[1]
    h = torch.cat([en, vn], dim=1)
    B = h.shape[0]
    q, k, v = linear(sd, p + ".to_q", h), linear(sd, p + ".to_k", h), linear(sd, p + ".to_v", h)
    hd = k.shape[-1] // heads
    q = q.view(B, -1, heads, hd).transpose(1, 2)
    k = k.view(B, -1, heads, hd).transpose(1, 2)
    v = v.view(B, -1, heads, hd).transpose(1, 2)
    q = F.layer_norm(q, (hd,), sd[p + ".norm_q.weight"], sd[p + ".norm_q.bias"], 1e-6)  # per-head LN, eps 1e-6
    k = F.layer_norm(k, (hd,), sd[p + ".norm_k.weight"], sd[p + ".norm_k.bias"], 1e-6)
    if rope is not None:
        q[:, :, T:] = apply_rope_real(q[:, :, T:], *rope)
        k[:, :, T:] = apply_rope_real(k[:, :, T:], *rope)
    return q, k, v


def attn_post(sd: SD, p: str, o: torch.Tensor, heads: int, T: int):
    """CogVideoXAttnMOTProcessor2_0.__call__(is_before_attn=False), attention_processor.py:2947-2959."""
    B, _, L, hd = o.shape
    h = linear(sd, p + ".to_out.0", o.transpose(1, 2).reshape(B, L, heads * hd))
    e, v = h.split([T, L - T], dim=1)
    return v, e


def cog_block(sd: SD, p: str, cfg: dict, with_mot_ref: bool, v: torch.Tensor, e: torch.Tensor, temb: torch.Tensor, rope,
              v_ref: Optional[torch.Tensor] = None, e_ref: Optional[torch.Tensor] = None,
              temb_ref: Optional[torch.Tensor] = None, temb_list_ref: Optional[List[torch.Tensor]] = None, rope_ref=None,
              trace: Optional[Callable[[str, torch.Tensor], None]] = None):
    """CogVideoXBlock.forward, cogvideox_transformer_3d_mot.py:156-515 (plain :171-203, MoT :375-513).
    `p` = "transformer_blocks.<i>".  v = video tokens, e = text tokens."""
    heads, eps = cfg["num_attention_heads"], cfg["norm_eps"]
    t = trace or (lambda name, tensor: None)
    T = e.shape[1]
    if not with_mot_ref:
        vn, en, gate, e_gate = layer_norm_zero(sd, p + ".norm1", v, e, temb, eps)
        q, k, vv = attn_pre(sd, p + ".attn1", vn, en, heads, rope)
        av, ae = attn_post(sd, p + ".attn1", sdpa(q, k, vv), heads, T)
        v = v + gate * av
        e = e + e_gate * ae
        vn, en, gate_ff, e_gate_ff = layer_norm_zero(sd, p + ".norm2", v, e, temb, eps)
        ff = feed_forward(sd, p + ".ff", torch.cat([en, vn], dim=1))
        v = v + gate_ff * ff[:, T:]
        e = e + e_gate_ff * ff[:, :T]
        return v, e, v_ref, e_ref

    B, S, d = v.shape
    S_ref, T_ref = v_ref.shape[-2], e_ref.shape[-2]
    n = S_ref // S
    multi = temb_list_ref is not None
    if multi == (temb_ref is not None):
        raise NotImplementedError("exactly one of temb_mot_ref / temb_list_mot_ref must be given")  # :402-403

    vn, en, gate, e_gate = layer_norm_zero(sd, p + ".norm1", v, e, temb, eps)
    if not multi:
        vn_r, en_r, gate_r, e_gate_r = layer_norm_zero(sd, p + ".norm1_mot_ref", v_ref, e_ref, temb_ref, eps)
    else:  # :393-401
        vn_r, en_r, gate_r, e_gate_r = layer_norm_zero(
            sd, p + ".norm1_mot_ref", v_ref.reshape(B * n, S, d), e_ref.reshape(B * n, T, d), torch.cat(temb_list_ref, 0), eps)
        vn_r, en_r = vn_r.reshape(B, n * S, d), en_r.reshape(B, n * T, d)
    t("norm1_video", vn), t("norm1_text", en), t("norm1_video_ref", vn_r), t("norm1_text_ref", en_r)

    q, k, vv = attn_pre(sd, p + ".attn1", vn, en, heads, rope)
    q_r, k_r, vv_r = attn_pre(sd, p + ".attn1_mot_ref", vn_r, en_r, heads, rope_ref)
    t("q", q), t("k", k), t("v", vv), t("q_ref", q_r), t("k_ref", k_r), t("v_ref", vv_r)
    o = sdpa(torch.cat([q, q_r], dim=-2), torch.cat([k, k_r], dim=-2), torch.cat([vv, vv_r], dim=-2))  # :424-431
    t("attn_joint", o)
    av, ae = attn_post(sd, p + ".attn1", o[..., : S + T, :], heads, T)
    av_r, ae_r = attn_post(sd, p + ".attn1_mot_ref", o[..., S + T:, :], heads, T_ref)

    v = v + gate * av
    e = e + e_gate * ae
    vn, en, gate_ff, e_gate_ff = layer_norm_zero(sd, p + ".norm2", v, e, temb, eps)
    ff = feed_forward(sd, p + ".ff", torch.cat([en, vn], dim=1))
    v = v + gate_ff * ff[:, T:]
    e = e + e_gate_ff * ff[:, :T]

    if not multi:  # :464-470, :498-500
        v_ref = v_ref + gate_r * av_r
        e_ref = e_ref + e_gate_r * ae_r
        vn_r, en_r, gate_ff_r, e_gate_ff_r = layer_norm_zero(sd, p + ".norm2_mot_ref", v_ref, e_ref, temb_ref, eps)
        ff_r = feed_forward(sd, p + ".ff_mot_ref", torch.cat([en_r, vn_r], dim=1))
        v_ref = v_ref + gate_ff_r * ff_r[:, T_ref:]
        e_ref = e_ref + e_gate_ff_r * ff_r[:, :T_ref]
    else:  # :471-488, :501-511
        v_ref = (v_ref.reshape(B, n, S, d) + gate_r.reshape(B, n, 1, d) * av_r.reshape(B, n, S, d)).reshape(B, -1, d)
        e_ref = (e_ref.reshape(B, n, T, d) + e_gate_r.reshape(B, n, 1, d) * ae_r.reshape(B, n, T, d)).reshape(B, -1, d)
        vn_r, en_r, gate_ff_r, e_gate_ff_r = layer_norm_zero(
            sd, p + ".norm2_mot_ref", v_ref.reshape(B * n, S, d), e_ref.reshape(B * n, T, d), torch.cat(temb_list_ref, 0), eps)
        vn_r, en_r = vn_r.reshape(B, n * S, d), en_r.reshape(B, n * T, d)
        ff_r = feed_forward(sd, p + ".ff_mot_ref", torch.cat([en_r, vn_r], dim=1))
        v_ref = (v_ref.reshape(B, n, S, d) + gate_ff_r.reshape(B, n, 1, d) * ff_r[:, T_ref:].reshape(B, n, S, d)).reshape(B, -1, d)
        e_ref = (e_ref.reshape(B, n, T, d) + e_gate_ff_r.reshape(B, n, 1, d) * ff_r[:, :T_ref].reshape(B, n, T, d)).reshape(B, -1, d)
    return v, e, v_ref, e_ref


# ----------------------------------------------------------------------------------------------
# Transformer shell
# ----------------------------------------------------------------------------------------------
def _time_embedding(sd: SD, p: str, timestep: torch.Tensor, inner_dim: int, dtype: torch.dtype) -> torch.Tensor:
    """time_proj + time_embedding, cogvideox_transformer_3d_mot.py:923-931 / :944-948."""
    t_emb = timestep_embedding(timestep, inner_dim, flip_sin_to_cos=True, downscale_freq_shift=0).to(dtype)
    return linear(sd, p + ".linear_2", F.silu(linear(sd, p + ".linear_1", t_emb)))


def _patch_embed(sd: SD, p: str, cfg: dict, text: torch.Tensor, video: torch.Tensor) -> torch.Tensor:
    """CogVideoXPatchEmbed.forward (patch_size_t None), models/embeddings.py:701-757, with rotary
    positional embeddings (no sincos table) and the optional learned `pos_embedding` buffer."""
    text = linear(sd, p + ".text_proj", text)
    B, Fr, C, H, W = video.shape
    ps = cfg["patch_size"]
    x = F.conv2d(video.reshape(-1, C, H, W), sd[p + ".proj.weight"], sd.get(p + ".proj.bias"), stride=ps)
    x = x.view(B, Fr, *x.shape[1:]).flatten(3).transpose(2, 3).flatten(1, 2)
    emb = torch.cat([text, x], dim=1).contiguous()
    if cfg.get("use_learned_positional_embeddings", False):
        emb = emb + sd[p + ".pos_embedding"].to(dtype=emb.dtype)
    return emb


def cog_forward(sd: SD, cfg: dict, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor, timestep: torch.Tensor,
                image_rotary_emb, hidden_states_mot_ref: torch.Tensor, encoder_hidden_states_mot_ref: torch.Tensor,
                image_rotary_emb_mot_ref, num_mot_ref: int = 1, timestep_list_mot_ref=None,
                block_io: Optional[Dict[int, dict]] = None) -> torch.Tensor:
    """CogVideoXTransformer3DMOTModel.forward, cogvideox_transformer_3d_mot.py:886-1106
    (no ofs / ref / effect embeddings, reference_train_mode None)."""
    B, Fr, C, H, W = hidden_states.shape
    inner = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    Ttok = encoder_hidden_states.shape[-2]
    dt = hidden_states.dtype
    emb = _time_embedding(sd, "time_embedding", timestep, inner, dt)
    if timestep_list_mot_ref is not None:
        emb_list_r = [_time_embedding(sd, "time_embedding_mot_ref", ts, inner, dt) for ts in timestep_list_mot_ref]
        emb_r = None
    else:
        emb_r = _time_embedding(sd, "time_embedding_mot_ref", timestep, inner, dt)
        emb_list_r = None
    assert hidden_states_mot_ref.shape[1] // Fr == num_mot_ref

    h = _patch_embed(sd, "patch_embed", cfg, encoder_hidden_states, hidden_states)
    e, v = h[:, :Ttok], h[:, Ttok:]
    vs, es = [], []
    for i in range(num_mot_ref):
        hi = _patch_embed(sd, "patch_embed_mot_ref", cfg, encoder_hidden_states_mot_ref[:, i * Ttok:(i + 1) * Ttok],
                          hidden_states_mot_ref[:, i * Fr:(i + 1) * Fr])
        es.append(hi[:, :Ttok]), vs.append(hi[:, Ttok:])
    v_r, e_r = torch.cat(vs, dim=1), torch.cat(es, dim=1)

    for i in range(cfg["num_layers"]):
        mot = i in cfg["block_idx_with_mot_ref"]
        if block_io is not None:
            block_io[i] = dict(v=v, e=e, v_ref=v_r, e_ref=e_r, temb=emb, temb_ref=emb_r, temb_list_ref=emb_list_r)
        v, e, v_r, e_r = cog_block(sd, f"transformer_blocks.{i}", cfg, mot, v, e, emb, image_rotary_emb, v_r, e_r, emb_r,
                                   emb_list_r, image_rotary_emb_mot_ref)
        if block_io is not None:
            block_io[i].update(out_v=v, out_e=e, out_v_ref=v_r, out_e_ref=e_r)

    d = v.shape[-1]
    v = F.layer_norm(v, (d,), sd["norm_final.weight"], sd["norm_final.bias"], cfg["norm_eps"])
    # AdaLayerNorm(chunk_dim=1), models/normalization.py:65-81: shift first, then scale
    shift, scale = linear(sd, "norm_out.linear", F.silu(emb)).chunk(2, dim=1)
    v = F.layer_norm(v, (d,), sd["norm_out.norm.weight"], sd["norm_out.norm.bias"], cfg["norm_eps"]) * (1 + scale[:, None, :]) + shift[:, None, :]
    v = linear(sd, "proj_out", v)
    p = cfg["patch_size"]
    out = v.reshape(B, Fr, H // p, W // p, -1, p, p)
    return out.permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)
