"""Oracle: Wan2.1-VAP MoT transformer (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Functional restatement of models/transformers/transformer_wan_mot.py on a flat state_dict.
All tensors keep the reference's dtypes, so on CPU with the same torch build the result is
bit-identical to the reference module (checked by oracle/gen_golden.py).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

from .common import SD, feed_forward, fp32_layer_norm, gelu_tanh, linear, rms_norm_across, sdpa, timestep_embedding

TEXT_CONTEXT_LEN = 512  # hardcoded in the reference, transformer_wan_mot.py:126-127


# ----------------------------------------------------------------------------------------------
# RoPE tables
# ----------------------------------------------------------------------------------------------
def _rope_1d(dim: int, pos: torch.Tensor, theta: float) -> torch.Tensor:
    """get_1d_rotary_pos_embed(use_real=False, freqs_dtype=float64), models/embeddings.py:1181-1205."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64)[: dim // 2] / dim))
    freqs = torch.outer(pos, freqs)
    return torch.polar(torch.ones_like(freqs), freqs)


def wan_rope(head_dim: int, patch_size, max_seq_len: int, latent_shape, ref: bool, theta: float = 10000.0) -> torch.Tensor:
    """WanRotaryPosEmbed.forward (transformer_wan_mot.py:390-409) and WanRotaryPosEmbedRef.forward
    (:429-464).  latent_shape = (frames, height, width) of the *unpatched* latent.  Returns complex128
    [1, 1, S, head_dim/2]; the ref table's temporal positions start at -frames (:437)."""
    num_frames, height, width = latent_shape
    p_t, p_h, p_w = patch_size
    ppf, pph, ppw = num_frames // p_t, height // p_h, width // p_w
    h_dim = w_dim = 2 * (head_dim // 6)
    t_dim = head_dim - h_dim - w_dim
    if ref:
        t_pos = torch.arange(-num_frames, max_seq_len)
    else:
        t_pos = torch.arange(max_seq_len)
    f_t = _rope_1d(t_dim, t_pos, theta)[:max_seq_len]
    f_h = _rope_1d(h_dim, torch.arange(max_seq_len), theta)
    f_w = _rope_1d(w_dim, torch.arange(max_seq_len), theta)
    f_t = f_t[:ppf].view(ppf, 1, 1, -1).expand(ppf, pph, ppw, -1)
    f_h = f_h[:pph].view(1, pph, 1, -1).expand(ppf, pph, ppw, -1)
    f_w = f_w[:ppw].view(1, 1, ppw, -1).expand(ppf, pph, ppw, -1)
    return torch.cat([f_t, f_h, f_w], dim=-1).reshape(1, 1, ppf * pph * ppw, -1)


def apply_rope_complex(x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """apply_rotary_emb inside WanAttnMOTProcessor2_0, transformer_wan_mot.py:229-233 (float64 complex)."""
    xr = torch.view_as_complex(x.to(torch.float64).unflatten(3, (-1, 2)))
    return torch.view_as_real(xr * freqs).flatten(3, 4).type_as(x)


# ----------------------------------------------------------------------------------------------
# Attention halves
# ----------------------------------------------------------------------------------------------
def self_attn_pre(sd: SD, p: str, x: torch.Tensor, heads: int, eps: float, freqs: Optional[torch.Tensor]):
    """WanAttnMOTProcessor2_0.__call__(is_before_attn=True), transformer_wan_mot.py:211-238."""
    q = linear(sd, p + ".to_q", x)
    k = linear(sd, p + ".to_k", x)
    v = linear(sd, p + ".to_v", x)
    q = rms_norm_across(q, sd[p + ".norm_q.weight"], eps)
    k = rms_norm_across(k, sd[p + ".norm_k.weight"], eps)
    q = q.unflatten(2, (heads, -1)).transpose(1, 2)
    k = k.unflatten(2, (heads, -1)).transpose(1, 2)
    v = v.unflatten(2, (heads, -1)).transpose(1, 2)
    if freqs is not None:
        q = apply_rope_complex(q, freqs)
        k = apply_rope_complex(k, freqs)
    return q, k, v


def self_attn_post(sd: SD, p: str, o: torch.Tensor) -> torch.Tensor:
    """WanAttnMOTProcessor2_0.__call__(is_before_attn=False), transformer_wan_mot.py:240-244."""
    return linear(sd, p + ".to_out.0", o.transpose(1, 2).flatten(2, 3))


def self_attn_plain(sd: SD, p: str, x: torch.Tensor, heads: int, eps: float, freqs: torch.Tensor) -> torch.Tensor:
    """WanAttnProcessor2_0 self-attention (non-MoT block), transformer_wan_mot.py:39-107 with
    encoder_hidden_states=None (no image branch)."""
    q, k, v = self_attn_pre(sd, p, x, heads, eps, freqs)
    o = sdpa(q, k, v).transpose(1, 2).flatten(2, 3).type_as(q)
    return linear(sd, p + ".to_out.0", o)


def cross_attn(sd: SD, p: str, x: torch.Tensor, ctx: torch.Tensor, heads: int, eps: float, num_mot_ref: int = 1) -> torch.Tensor:
    """WanAttnCrossMOTProcessor2_0.__call__, transformer_wan_mot.py:115-190 (also equals
    WanAttnProcessor2_0's I2V cross-attention :50-107 when num_mot_ref == 1).

    ctx = [image tokens | text tokens]; two independent softmaxes (image, text) summed in bf16."""
    n = num_mot_ref
    img_len = ctx.shape[1] - TEXT_CONTEXT_LEN * n
    ctx_img, ctx_txt = ctx[:, :img_len], ctx[:, img_len:]
    q = rms_norm_across(linear(sd, p + ".to_q", x), sd[p + ".norm_q.weight"], eps)
    k = rms_norm_across(linear(sd, p + ".to_k", ctx_txt), sd[p + ".norm_k.weight"], eps)
    v = linear(sd, p + ".to_v", ctx_txt)
    k_img = rms_norm_across(linear(sd, p + ".add_k_proj", ctx_img), sd[p + ".norm_added_k.weight"], eps)
    v_img = linear(sd, p + ".add_v_proj", ctx_img)

    def heads_split(t):  # 'b (n l) (h c) -> (b n) h l c'
        b, L, _ = t.shape
        t = t.unflatten(2, (heads, -1)).transpose(1, 2)  # b h (n l) c
        return t.reshape(b, heads, n, L // n, -1).permute(0, 2, 1, 3, 4).reshape(b * n, heads, L // n, -1)

    def heads_merge(t):  # '(b n) h l c -> b (n l) (h c)'
        bn, h, l, c = t.shape
        b = bn // n
        t = t.reshape(b, n, h, l, c).permute(0, 2, 1, 3, 4).reshape(b, h, n * l, c)
        return t.transpose(1, 2).flatten(2, 3)

    qh = heads_split(q)
    o_img = heads_merge(sdpa(qh, heads_split(k_img), heads_split(v_img))).type_as(q)
    o_txt = heads_merge(sdpa(qh, heads_split(k), heads_split(v))).type_as(q)
    return linear(sd, p + ".to_out.0", o_txt + o_img)


# ----------------------------------------------------------------------------------------------
# Block
# ----------------------------------------------------------------------------------------------
def _modulated_ln(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, eps: float) -> torch.Tensor:
    """(FP32LayerNorm_noaffine(x.float()) * (1 + scale) + shift).type_as(x), transformer_wan_mot.py:620-623."""
    return (fp32_layer_norm(x.float(), None, None, eps) * (1 + scale) + shift).type_as(x)


def wan_block(sd: SD, p: str, cfg: dict, with_mot_ref: bool, x: torch.Tensor, ctx: torch.Tensor, temb: torch.Tensor,
              freqs: torch.Tensor, x_ref: Optional[torch.Tensor] = None, ctx_ref: Optional[torch.Tensor] = None,
              temb_ref: Optional[torch.Tensor] = None, freqs_ref: Optional[torch.Tensor] = None, num_mot_ref: int = 1,
              trace: Optional[Callable[[str, torch.Tensor], None]] = None):
    """WanTransformerBlock.forward, transformer_wan_mot.py:566-699.  `p` = "blocks.<i>"."""
    heads, eps = cfg["num_attention_heads"], cfg["eps"]
    t = trace or (lambda name, tensor: None)
    mod = sd[p + ".scale_shift_table"] + temb.float()  # [B,6,d] fp32 (:606-608)
    shift, scale, gate, c_shift, c_scale, c_gate = mod.chunk(6, dim=1)

    if not with_mot_ref:  # :580-601
        xn = _modulated_ln(x, scale, shift, eps)
        a = self_attn_plain(sd, p + ".attn1", xn, heads, eps, freqs)
        x = (x.float() + a * gate).type_as(x)
        xc = fp32_layer_norm(x.float(), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps).type_as(x)
        x = x + cross_attn(sd, p + ".attn2", xc, ctx, heads, eps, 1)
        xf = _modulated_ln(x, c_scale, c_shift, eps)
        ff = feed_forward(sd, p + ".ffn", xf)
        x = (x.float() + ff.float() * c_gate).type_as(x)
        return x, x_ref

    assert num_mot_ref == 1  # :611
    mod_r = sd[p + ".scale_shift_table_mot_ref"] + temb_ref.float()  # [(b n),6,d]
    shift_r, scale_r, gate_r, c_shift_r, c_scale_r, c_gate_r = mod_r.chunk(6, dim=1)  # n == 1: (b n) t c == b t c

    # 1. joint self-attention (:620-663)
    xn = _modulated_ln(x, scale, shift, eps)
    xn_r = _modulated_ln(x_ref, scale_r, shift_r, eps)
    t("norm1", xn), t("norm1_ref", xn_r)
    q, k, v = self_attn_pre(sd, p + ".attn1", xn, heads, eps, freqs)
    q_r, k_r, v_r = self_attn_pre(sd, p + ".attn1_mot_ref", xn_r, heads, eps, freqs_ref)
    t("q", q), t("k", k), t("v", v), t("q_ref", q_r), t("k_ref", k_r), t("v_ref", v_r)
    o = sdpa(torch.cat([q, q_r], dim=-2), torch.cat([k, k_r], dim=-2), torch.cat([v, v_r], dim=-2)).type_as(q)
    t("attn_joint", o)
    o_t, o_r = torch.split(o, [q.shape[-2], q_r.shape[-2]], dim=-2)
    a = self_attn_post(sd, p + ".attn1", o_t)
    a_r = self_attn_post(sd, p + ".attn1_mot_ref", o_r)
    x = (x.float() + a * gate).type_as(x)
    x_ref = (x_ref.float() + a_r * gate_r).type_as(x_ref)
    t("after_attn1", x), t("after_attn1_ref", x_ref)

    # 2. per-stream cross-attention (:668-676)
    xc = fp32_layer_norm(x.float(), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps).type_as(x)
    xc_r = fp32_layer_norm(x_ref.float(), sd[p + ".norm2_mot_ref.weight"], sd[p + ".norm2_mot_ref.bias"], eps).type_as(x_ref)
    x = x + cross_attn(sd, p + ".attn2", xc, ctx, heads, eps, 1)
    x_ref = x_ref + cross_attn(sd, p + ".attn2_mot_ref", xc_r, ctx_ref, heads, eps, num_mot_ref)
    t("after_attn2", x), t("after_attn2_ref", x_ref)

    # 3. per-stream FFN (:680-697)
    xf = _modulated_ln(x, c_scale, c_shift, eps)
    x = (x.float() + feed_forward(sd, p + ".ffn", xf).float() * c_gate).type_as(x)
    xf_r = _modulated_ln(x_ref, c_scale_r, c_shift_r, eps)
    x_ref = (x_ref.float() + feed_forward(sd, p + ".ffn_mot_ref", xf_r).float() * c_gate_r).type_as(x_ref)
    return x, x_ref


# ----------------------------------------------------------------------------------------------
# Transformer shell
# ----------------------------------------------------------------------------------------------
def _condition_embedder(sd: SD, p: str, cfg: dict, timesteps: List[torch.Tensor], text: torch.Tensor, image: Optional[torch.Tensor]):
    """WanTimeTextImageEmbedding.forward / WanTimeTextImageEmbeddingRef.forward, transformer_wan_mot.py:293-365."""
    w_dtype = sd[p + ".time_embedder.linear_1.weight"].dtype
    tembs, projs = [], []
    for ts in timesteps:
        e = timestep_embedding(ts, cfg["freq_dim"], flip_sin_to_cos=True, downscale_freq_shift=0)
        if e.dtype != w_dtype:
            e = e.to(w_dtype)
        e = linear(sd, p + ".time_embedder.linear_2", F.silu(linear(sd, p + ".time_embedder.linear_1", e))).type_as(text)
        tembs.append(e)
        projs.append(linear(sd, p + ".time_proj", F.silu(e)))
    temb, proj = torch.cat(tembs, 0), torch.cat(projs, 0)
    text = linear(sd, p + ".text_embedder.linear_2", gelu_tanh(linear(sd, p + ".text_embedder.linear_1", text)))
    if image is not None:  # WanImageEmbedding.forward :259-268 (pos_embed None)
        q = p + ".image_embedder"
        image = fp32_layer_norm(image, sd[q + ".norm1.weight"], sd[q + ".norm1.bias"], 1e-5)
        image = feed_forward(sd, q + ".ff", image, approximate="none")
        image = fp32_layer_norm(image, sd[q + ".norm2.weight"], sd[q + ".norm2.bias"], 1e-5)
    return temb, proj, text, image


def wan_forward(sd: SD, cfg: dict, hidden_states: torch.Tensor, timestep: torch.Tensor, encoder_hidden_states: torch.Tensor,
                encoder_hidden_states_image: torch.Tensor, hidden_states_mot_ref: torch.Tensor,
                timestep_list_mot_ref: torch.Tensor, encoder_hidden_states_mot_ref: torch.Tensor,
                encoder_hidden_states_image_mot_ref: torch.Tensor, num_mot_ref: int = 1,
                block_io: Optional[Dict[int, dict]] = None) -> torch.Tensor:
    """WanTransformer3DMOTModel.forward, transformer_wan_mot.py:854-1000 (reference_train_mode None).
    If `block_io` is a dict it receives, per block index, that block's inputs and outputs."""
    B, C, Fr, Hh, Ww = hidden_states.shape
    p_t, p_h, p_w = cfg["patch_size"]
    D = cfg["attention_head_dim"]
    freqs = wan_rope(D, cfg["patch_size"], cfg["rope_max_seq_len"], hidden_states.shape[2:], ref=False)
    freqs_ref = wan_rope(D, cfg["patch_size"], cfg["rope_max_seq_len"], hidden_states_mot_ref.shape[2:], ref=True)

    x = F.conv3d(hidden_states, sd["patch_embedding.weight"], sd["patch_embedding.bias"], stride=tuple(cfg["patch_size"]))
    x = x.flatten(2).transpose(1, 2)
    x_ref = F.conv3d(hidden_states_mot_ref, sd["patch_embedding_mot_ref.weight"], sd["patch_embedding_mot_ref.bias"],
                     stride=tuple(cfg["patch_size"]))
    x_ref = x_ref.flatten(2).transpose(1, 2)

    temb, proj, ctx, ctx_img = _condition_embedder(sd, "condition_embedder", cfg, [timestep], encoder_hidden_states,
                                                   encoder_hidden_states_image)
    proj = proj.unflatten(1, (6, -1))
    temb_r, proj_r, ctx_r, ctx_img_r = _condition_embedder(sd, "condition_embedder_mot_ref", cfg, list(timestep_list_mot_ref),
                                                           encoder_hidden_states_mot_ref, encoder_hidden_states_image_mot_ref)
    proj_r = proj_r.unflatten(1, (6, -1))
    if ctx_img is not None:
        ctx = torch.cat([ctx_img, ctx], dim=1)
        ctx_r = torch.cat([ctx_img_r, ctx_r], dim=1)

    for i in range(cfg["num_layers"]):
        mot = i in cfg["block_idx_with_mot_ref"]
        if block_io is not None:
            block_io[i] = dict(x=x, x_ref=x_ref, ctx=ctx, ctx_ref=ctx_r, temb=proj, temb_ref=proj_r)
        x, x_ref = wan_block(sd, f"blocks.{i}", cfg, mot, x, ctx, proj, freqs, x_ref, ctx_r, proj_r, freqs_ref, num_mot_ref)
        if block_io is not None:
            block_io[i].update(out=x, out_ref=x_ref)

    shift, scale = (sd["scale_shift_table"] + temb.unsqueeze(1)).chunk(2, dim=1)  # :952
    x = (fp32_layer_norm(x.float(), None, None, cfg["eps"]) * (1 + scale) + shift).type_as(x)
    x = linear(sd, "proj_out", x)
    x = x.reshape(B, Fr // p_t, Hh // p_h, Ww // p_w, p_t, p_h, p_w, -1)
    x = x.permute(0, 7, 1, 4, 2, 5, 3, 6)
    return x.flatten(6, 7).flatten(4, 5).flatten(2, 3)
