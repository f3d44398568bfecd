"""CogVideoX-VAP: fused MoT block forward, drop-in attention processor and the stand-alone transformer shell.

Reference (paths relative to /root/reference/diffusers/src/diffusers):
  CogVideoXBlock.forward  models/transformers/cogvideox_transformer_3d_mot.py:156-515 (plain :171-203, MoT :375-513)
                          -> `cog_block_forward`  (boundary B3)
  CogVideoXAttnMOTProcessor2_0 / CogVideoXAttnProcessor2_0  models/attention_processor.py:2890-2959 / :2822-2888
                          -> same-named classes here (boundary B1, same __call__ kwargs)
  CogVideoXLayerNormZero  models/normalization.py:449-471 (all-bf16 tensor arithmetic -> ROUND_COG)
  CogVideoXTransformer3DMOTModel  cogvideox_transformer_3d_mot.py:517-1106 -> `CogVideoXTransformer3DMOTModel`

Joint sequence order of the MoT attention is the reference's: [text | video | text_ref | video_ref]
(attention_processor.py:2915, cogvideox_transformer_3d_mot.py:424-427); per stream the LayerNorm kernels write the
modulated text and video tokens straight into one [T+S, d] buffer that feeds the fused QKV GEMM, whose output lands in
the joint [J, 3d] buffer the attention kernel reads through strided TMA descriptors.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, ulysses
from . import streams as stream_pair
from .modules import AdaLayerNorm, Attention, CogVideoXLayerNormZero, FeedForward, TimestepEmbedding
from .rope import Tables, as_tables
from .wan import _f32, _heads_view, _joint_attention, _linear, _Output, _packed, _sp_p2p, _token_major


def _mods(norm: nn.Module, temb: torch.Tensor) -> torch.Tensor:
    """CogVideoXLayerNormZero modulation (normalization.py:467): Linear(silu(temb)) in the model dtype, with the two
    `1 + scale` additions done in that dtype too; returned as fp32 [nb, 6, d]:
    (shift, 1+scale, gate, enc_shift, 1+enc_scale, enc_gate)."""
    m = F.linear(F.silu(temb), norm.linear.weight, norm.linear.bias)
    m = m.view(m.shape[0], 6, -1).clone()
    m[:, 1] += 1
    m[:, 4] += 1
    return m.float()


def _ln_zero(norm: nn.Module, x: torch.Tensor, mods: torch.Tensor, text: bool, rows_per_mod: int, out: torch.Tensor) -> None:
    """LN_affine(x) * (1+scale) + shift with the reference's bf16 roundings (normalization.py:468-469).
    x [rows, d] of ONE batch element, mods fp32 [nmod, 6, d] (rows = nmod * rows_per_mod); video tokens use chunks
    (shift 0, 1+scale 1), text tokens chunks (3, 4).  `out` is a row-slice of the stream's [T+S, d] QKV input buffer."""
    i = 3 if text else 0
    ops.adaln_layernorm(x, eps=norm.norm.eps, rounding=ops.ROUND_COG, ln_w=_f32(norm.norm, "w", norm.norm.weight),
                        ln_b=_f32(norm.norm, "b", norm.norm.bias), scale1p=mods[:, i + 1], shift=mods[:, i], rows_per_batch=rows_per_mod, out=out)


def _qkv(attn: nn.Module, h: torch.Tensor, T: int, tables: Optional[Tables], out: torch.Tensor, scatter=None) -> None:
    """to_q/k/v + per-head LayerNorm + RoPE on the video tokens (attention_processor.py:2923-2945), one batch element.
    h [L, d] -> out [L, 3*inner] (row-slice of the joint buffer).
    scatter = (PeerExchange, row0): Ulysses peer-memory mode — the norm + RoPE kernel stores its results (and V) straight into the
    receive buffers of the ranks owning the heads; `out` then only holds the raw projection."""
    W, bvec = _packed(attn, "qkv", [attn.to_q, attn.to_k, attn.to_v])
    inner = W.shape[0] // 3
    heads = attn.heads
    ops.linear(h, W, bvec, out=out)
    if attn.norm_q is None:
        if tables is not None or scatter is not None:
            raise NotImplementedError("RoPE / sequence parallelism without qk LayerNorm is not a CogVideoX-VAP configuration")
        return
    norm_rope = dict(wq=_f32(attn.norm_q, "w", attn.norm_q.weight), bq=_f32(attn.norm_q, "b", attn.norm_q.bias),
                     wk=_f32(attn.norm_k, "w", attn.norm_k.weight), bk=_f32(attn.norm_k, "b", attn.norm_k.bias),
                     cos=tables[0] if tables else None, sin=tables[1] if tables else None, rows_per_batch=h.shape[0], rope_row0=T,
                     eps=attn.norm_q.eps, mode=ops.QK_COG)
    if scatter is not None:
        scatter[0].dispatch(out, scatter[1], batch_index=scatter[2] if len(scatter) > 2 else 0, **norm_rope)
    else:
        ops.qk_norm_rope_(out[:, :inner], out[:, inner:2 * inner], heads=heads, head_dim=inner // heads, **norm_rope)


def _gated(lin: nn.Linear, a: torch.Tensor, res: torch.Tensor, gate: torch.Tensor, rows_per_gate: int, out: torch.Tensor) -> None:
    """out = res + gate * Linear(a) with the reference's bf16 rounding points (cogvideox_transformer_3d_mot.py:445-446).
    a [rows, K], res/out [rows, d] (rows = ngate * rows_per_gate), gate fp32 [ngate, d]."""
    ops.linear(a, lin.weight, lin.bias, epilogue=ops.EPI_GATE_RES_BF16, residual=res, gate=gate, rows_per_batch=rows_per_gate, out=out)


class _Stream:
    """One token stream (target or reference) of a CogVideoX MoT block: its modules, tokens and modulation layout."""

    def __init__(self, block: nn.Module, sfx: str, v: torch.Tensor, e: torch.Tensor, temb: torch.Tensor, nmod: int, tables):
        self.norm1, self.attn, self.norm2, self.ff = (getattr(block, n + sfx) for n in ("norm1", "attn1", "norm2", "ff"))
        self.v, self.e = v.contiguous(), e.contiguous()  # [B, S, d], [B, T, d]
        self.temb = temb                                  # [B * nmod, time_embed_dim]
        self.nmod = nmod                                  # modulation vectors per batch element (n refs, or 1)
        self.S, self.T = v.shape[1], e.shape[1]
        self.sv, self.se = max(self.S // nmod, 1), max(self.T // nmod, 1)  # tokens per modulation vector
        self.tables = tables
        self.mods1 = _mods(self.norm1, temb)

    def mods_of(self, mods: torch.Tensor, b: int) -> torch.Tensor:
        return mods[b * self.nmod:(b + 1) * self.nmod]


def cog_block_forward(self: nn.Module, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor, temb: torch.Tensor,
                      image_rotary_emb=None, attention_kwargs: Optional[Dict[str, Any]] = None,
                      hidden_states_mot_ref: Optional[torch.Tensor] = None, encoder_hidden_states_mot_ref: Optional[torch.Tensor] = None,
                      temb_mot_ref: Optional[torch.Tensor] = None, temb_list_mot_ref: Optional[List[torch.Tensor]] = None,
                      image_rotary_emb_mot_ref=None):
    """Drop-in for CogVideoXBlock.forward (same signature, returns the same 4-tuple) on the sm_100a kernels."""
    if getattr(self, "ablation_single_encoder", False) or getattr(self, "ablation_residual_addition", False):
        raise NotImplementedError("the ablation branches (cogvideox_transformer_3d_mot.py:205-373) are outside the VAP hot path")
    if attention_kwargs:
        raise ValueError(f"attention_kwargs {list(attention_kwargs)} are not supported on the VAP path")
    B, S, d = hidden_states.shape
    heads = self.attn1.heads
    inner = self.attn1.to_q.weight.shape[0]
    hd = inner // heads
    dev = hidden_states.device
    streams = [_Stream(self, "", hidden_states, encoder_hidden_states, temb, 1, as_tables(image_rotary_emb, hd, dev))]
    if self.with_mot_ref:
        multi = temb_list_mot_ref is not None
        if multi and ulysses.current() is not None:
            raise NotImplementedError("Ulysses sequence parallelism with per-reference timesteps (temb_list_mot_ref) is not supported")
        if multi == (temb_mot_ref is not None):
            raise NotImplementedError("Not supprted for temb_list_mot_ref is not None and temb_mot_ref is not None or both are None")
        n = hidden_states_mot_ref.shape[1] // S if multi else 1
        # multi-ref: the reference reshapes the ref stream to (B*n, S, d) against cat(temb_list) (:393-401), i.e. reshaped
        # row r = b*n + i takes modulation row r of the concatenation
        temb_r = torch.cat(temb_list_mot_ref, dim=0) if multi else temb_mot_ref
        streams.append(_Stream(self, "_mot_ref", hidden_states_mot_ref, encoder_hidden_states_mot_ref, temb_r, n if multi else 1,
                               as_tables(image_rotary_emb_mot_ref, hd, dev)))

    # ---- norm1 + fused QKV of every stream into the joint buffer [text | video | text_ref | video_ref] ------------
    # Under Ulysses sequence parallelism the rows of every stream are this rank's shard of its [text | video] sequence (the shell
    # shards them: rank 0 holds the text rows, st.T may be 0 elsewhere); the joint attention then does the two exchanges.
    J = sum(st.T + st.S for st in streams)
    qkv = torch.empty((B, J, 3 * inner), dtype=torch.bfloat16, device=dev)
    px = _sp_p2p(hidden_states, J, heads, hd)  # PeerExchange (for the whole batch) in peer-memory mode, else None
    row0 = [0, streams[0].T + streams[0].S]    # first joint row of each stream
    # the expert's stream is issued on a side CUDA stream between the joint attentions (streams.py); None: CPU tensors / plain block / switched off
    ds = stream_pair.dual(dev, max(st.T + st.S for st in streams)) if len(streams) > 1 else None

    def pre(si: int) -> None:
        st = streams[si]
        L = st.T + st.S
        for b in range(B):
            h = torch.empty((L, d), dtype=torch.bfloat16, device=dev)
            m = st.mods_of(st.mods1, b)
            if st.T:
                _ln_zero(st.norm1, st.e[b], m, True, st.se, h[:st.T])
            if st.S:
                _ln_zero(st.norm1, st.v[b], m, False, st.sv, h[st.T:])
            _qkv(st.attn, h, st.T, st.tables, qkv[b, row0[si]:row0[si] + L], scatter=(px, row0[si], b) if px is not None else None)

    if ds is not None:
        ds.fork()
    if len(streams) > 1:
        with stream_pair.side(ds):
            pre(1)
    pre(0)
    if ds is not None:
        ds.join()

    # ---- joint attention ---------------------------------------------------------------------------------------
    if px is not None:
        o = px.attention()                  # [B, J, inner]: exchanges fused into the kernels over NVLink peer memory
    else:
        o = _joint_attention(qkv, heads)    # [B, J, inner]; NCCL all-to-alls around the kernel when Ulysses runs in "nccl" mode

    # ---- per stream: to_out + gated residual, norm2, FFN + gated residual ------------------------------------------
    def post(si: int):
        st = streams[si]
        row = row0[si]
        L = st.T + st.S
        v_out, e_out = torch.empty_like(st.v), torch.empty_like(st.e)
        mods2 = None
        for b in range(B):
            m1 = st.mods_of(st.mods1, b)
            e1 = torch.empty((st.T, d), dtype=torch.bfloat16, device=dev)
            v1 = torch.empty((st.S, d), dtype=torch.bfloat16, device=dev)
            if st.T:
                _gated(st.attn.to_out[0], o[b, row:row + st.T], st.e[b], m1[:, 5], st.se, e1)
            if st.S:
                _gated(st.attn.to_out[0], o[b, row + st.T:row + L], st.v[b], m1[:, 2], st.sv, v1)
            if mods2 is None:
                mods2 = _mods(st.norm2, st.temb)
            m2 = st.mods_of(mods2, b)
            h2 = torch.empty((L, d), dtype=torch.bfloat16, device=dev)
            if st.T:
                _ln_zero(st.norm2, e1, m2, True, st.se, h2[:st.T])
            if st.S:
                _ln_zero(st.norm2, v1, m2, False, st.sv, h2[st.T:])
            f1 = _linear(st.ff.net[0].proj, h2, epilogue=ops.EPI_BIAS_GELU)
            if st.T:
                _gated(st.ff.net[2], f1[:st.T], e1, m2[:, 5], st.se, e_out[b])
            if st.S:
                _gated(st.ff.net[2], f1[st.T:], v1, m2[:, 2], st.sv, v_out[b])
        return [v_out, e_out]

    if ds is not None:
        ds.fork()
    outs_r = []
    if len(streams) > 1:
        with stream_pair.side(ds):
            outs_r = post(1)
    outs = post(0) + outs_r
    if ds is not None:
        ds.join()  # `o` (read by the side stream) is still referenced here
    if not self.with_mot_ref:
        return outs[0], outs[1], hidden_states_mot_ref, encoder_hidden_states_mot_ref
    return outs[0], outs[1], outs[2], outs[3]


# ----------------------------------------------------------------------------------------------
# drop-in attention processors (boundary B1)
# ----------------------------------------------------------------------------------------------
class CogVideoXAttnMOTProcessor2_0:
    """Same two-phase protocol / kwarg names as the reference (attention_processor.py:2900-2959)."""

    def __call__(self, attn, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                 attention_mask: Optional[torch.Tensor] = None, image_rotary_emb=None, is_before_attn: bool = False,
                 is_ref_video: Optional[bool] = False, text_seq_length: Optional[int] = None):
        if is_before_attn:
            if attention_mask is not None:
                raise ValueError("attention masks are not supported on the VAP path (the reference block never passes one)")
            T = encoder_hidden_states.size(1)
            h = torch.cat([encoder_hidden_states, hidden_states], dim=1)
            B, L, _ = h.shape
            inner = attn.to_q.weight.shape[0]
            qkv = torch.empty((B, L, 3 * inner), dtype=torch.bfloat16, device=h.device)
            tables = as_tables(image_rotary_emb, inner // attn.heads, h.device)
            for b in range(B):
                _qkv(attn, h[b], T, tables, qkv[b])
            q, k, v = (_heads_view(qkv[..., i * inner:(i + 1) * inner], attn.heads) for i in range(3))
            return q, k, v, attention_mask
        out = _linear(attn.to_out[0], _token_major(hidden_states))
        e, v = out.split([text_seq_length, out.size(1) - text_seq_length], dim=1)
        return v, e


class CogVideoXAttnProcessor2_0:
    """Plain CogVideoX processor (attention_processor.py:2832-2888), used by blocks without the MoT branch."""

    def __call__(self, attn, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor,
                 attention_mask: Optional[torch.Tensor] = None, image_rotary_emb=None):
        mot = CogVideoXAttnMOTProcessor2_0()
        q, k, v, _ = mot(attn, hidden_states, encoder_hidden_states, attention_mask, image_rotary_emb, is_before_attn=True)
        return mot(attn, ops.attention(q, k, v), is_before_attn=False, text_seq_length=encoder_hidden_states.size(1))


# ----------------------------------------------------------------------------------------------
# stand-alone shell with the reference's module tree
# ----------------------------------------------------------------------------------------------
class CogVideoXBlock(nn.Module):
    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int, time_embed_dim: int, attention_bias: bool = False,
                 qk_norm: bool = True, norm_elementwise_affine: bool = True, norm_eps: float = 1e-5, ff_inner_dim: Optional[int] = None,
                 with_mot_ref: bool = False, _block_idx: int = 0):
        super().__init__()
        self.with_mot_ref = with_mot_ref
        self._block_idx = _block_idx
        self.ablation_single_encoder = False
        self.ablation_residual_addition = False

        def make(sfx: str):
            setattr(self, "norm1" + sfx, CogVideoXLayerNormZero(time_embed_dim, dim, norm_elementwise_affine, norm_eps, bias=True))
            setattr(self, "attn1" + sfx, Attention(dim, num_attention_heads, attention_head_dim, "layer_norm" if qk_norm else None, eps=1e-6,
                                                    bias=attention_bias, out_bias=True,
                                                    processor=CogVideoXAttnMOTProcessor2_0() if with_mot_ref else CogVideoXAttnProcessor2_0()))
            setattr(self, "norm2" + sfx, CogVideoXLayerNormZero(time_embed_dim, dim, norm_elementwise_affine, norm_eps, bias=True))
            setattr(self, "ff" + sfx, FeedForward(dim, inner_dim=ff_inner_dim, activation_fn="gelu-approximate", final_dropout=True))

        make("")
        if with_mot_ref:
            make("_mot_ref")

    forward = cog_block_forward


class CogVideoXPatchEmbed(nn.Module):
    """embeddings.py:626-757 with rotary positional embeddings (patch_size_t None): Conv2d patchify + text projection
    (+ the learned joint positional table of CogVideoX-5B-I2V).  Transformer-shell glue, torch."""

    def __init__(self, patch_size: int, in_channels: int, embed_dim: int, text_embed_dim: int, bias: bool, sample_width: int,
                 sample_height: int, sample_frames: int, temporal_compression_ratio: int, max_text_seq_length: int,
                 use_learned_positional_embeddings: bool):
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=(patch_size, patch_size), stride=patch_size, bias=bias)
        self.text_proj = nn.Linear(text_embed_dim, embed_dim)
        self.use_learned_positional_embeddings = use_learned_positional_embeddings
        self.sample_width, self.sample_height = sample_width, sample_height
        if use_learned_positional_embeddings:
            n = (sample_height // patch_size) * (sample_width // patch_size) * ((sample_frames - 1) // temporal_compression_ratio + 1)
            self.register_buffer("pos_embedding", torch.zeros(1, max_text_seq_length + n, embed_dim), persistent=True)

    def forward(self, text_embeds: torch.Tensor, image_embeds: torch.Tensor) -> torch.Tensor:
        text_embeds = self.text_proj(text_embeds)
        B, Fr, C, H, W = image_embeds.shape
        x = self.proj(image_embeds.reshape(-1, C, H, W))
        x = x.view(B, Fr, *x.shape[1:]).flatten(3).transpose(2, 3).flatten(1, 2)
        emb = torch.cat([text_embeds, x], dim=1).contiguous()
        if self.use_learned_positional_embeddings:
            if self.sample_width != W or self.sample_height != H:
                raise ValueError("It is currently not possible to generate videos at a different resolution that the defaults. "
                                 "This should only be the case with 'THUDM/CogVideoX-5b-I2V'.")
            emb = emb + self.pos_embedding.to(dtype=emb.dtype)
        return emb


def _shard_stream(seq: torch.Tensor, text_len: int, rotary, head_dim: int, sp, heads: int):
    """This rank's rows of one stream's [text | video] sequence [1, T + S, d] -> (text rows, video rows, RoPE tables of those video
    rows).  Rows are cut uniformly over the concatenation, as the reference's context-parallel split hook cuts its inputs
    (finetrainers/parallel/ptd.py:545-628)."""
    L = seq.shape[1]
    ulysses.check_divisible(L, heads, sp.world)
    n = L // sp.world
    r0 = sp.rank * n
    t_loc = min(max(text_len - r0, 0), n)   # text rows on this rank
    v0 = r0 + t_loc - text_len              # first video token on this rank
    loc = seq[:, r0:r0 + n]
    tables = as_tables(rotary, head_dim, seq.device)
    if tables is not None:  # a rank holding text rows only rotates nothing
        tables = tuple(t[v0:v0 + n - t_loc].contiguous() for t in tables) if n > t_loc else None
    return loc[:, :t_loc].contiguous(), loc[:, t_loc:].contiguous(), tables


class CogVideoXTransformer3DMOTModel(nn.Module):
    """Stand-alone mirror of the reference model (cogvideox_transformer_3d_mot.py:517-1106): same constructor kwargs (the
    subset the VAP checkpoints use), module names and forward signature."""

    def __init__(self, num_attention_heads: int = 30, attention_head_dim: int = 64, in_channels: int = 16, out_channels: Optional[int] = 16,
                 flip_sin_to_cos: bool = True, freq_shift: int = 0, time_embed_dim: int = 512, ofs_embed_dim: Optional[int] = None,
                 text_embed_dim: int = 4096, num_layers: int = 30, dropout: float = 0.0, attention_bias: bool = True, sample_width: int = 90,
                 sample_height: int = 60, sample_frames: int = 49, patch_size: int = 2, patch_size_t: Optional[int] = None,
                 temporal_compression_ratio: int = 4, max_text_seq_length: int = 226, activation_fn: str = "gelu-approximate",
                 timestep_activation_fn: str = "silu", norm_elementwise_affine: bool = True, norm_eps: float = 1e-5,
                 spatial_interpolation_scale: float = 1.875, temporal_interpolation_scale: float = 1.0,
                 use_rotary_positional_embeddings: bool = False, use_learned_positional_embeddings: bool = False, patch_bias: bool = True,
                 block_idx_with_mot_ref: List[int] = (0, 10, 20), attention_head_dim_mot_ref: Optional[int] = None,
                 supported_effect_types=None, num_ref_embeddings=None, reference_train_mode: Optional[str] = None,
                 ablation_single_encoder: bool = False, ablation_residual_addition: bool = False):
        super().__init__()
        if (not use_rotary_positional_embeddings or patch_size_t is not None or ofs_embed_dim or attention_head_dim_mot_ref is not None
                or supported_effect_types or num_ref_embeddings or reference_train_mode is not None or ablation_single_encoder
                or ablation_residual_addition or not flip_sin_to_cos or freq_shift != 0 or dropout != 0.0
                or activation_fn != "gelu-approximate" or timestep_activation_fn != "silu"):
            raise NotImplementedError("configuration outside the CogVideoX-VAP inference path (rotary embeddings, patch_size_t None, "
                                      "no ofs/effect/ref embeddings, no ablations)")
        inner = num_attention_heads * attention_head_dim
        self.config = dict(num_attention_heads=num_attention_heads, attention_head_dim=attention_head_dim, in_channels=in_channels,
                           out_channels=out_channels, time_embed_dim=time_embed_dim, text_embed_dim=text_embed_dim, num_layers=num_layers,
                           attention_bias=attention_bias, sample_width=sample_width, sample_height=sample_height, sample_frames=sample_frames,
                           patch_size=patch_size, patch_size_t=None, max_text_seq_length=max_text_seq_length,
                           norm_elementwise_affine=norm_elementwise_affine, norm_eps=norm_eps, use_rotary_positional_embeddings=True,
                           use_learned_positional_embeddings=use_learned_positional_embeddings,
                           block_idx_with_mot_ref=list(block_idx_with_mot_ref))
        pe = dict(patch_size=patch_size, in_channels=in_channels, embed_dim=inner, text_embed_dim=text_embed_dim, bias=patch_bias,
                  sample_width=sample_width, sample_height=sample_height, sample_frames=sample_frames,
                  temporal_compression_ratio=temporal_compression_ratio, max_text_seq_length=max_text_seq_length,
                  use_learned_positional_embeddings=use_learned_positional_embeddings)
        self.patch_embed = CogVideoXPatchEmbed(**pe)
        self.patch_embed_mot_ref = CogVideoXPatchEmbed(**pe)
        self.time_embedding = TimestepEmbedding(inner, time_embed_dim)
        self.time_embedding_mot_ref = TimestepEmbedding(inner, time_embed_dim)
        self.transformer_blocks = nn.ModuleList([
            CogVideoXBlock(inner, num_attention_heads, attention_head_dim, time_embed_dim, attention_bias=attention_bias,
                           norm_elementwise_affine=norm_elementwise_affine, norm_eps=norm_eps, with_mot_ref=i in block_idx_with_mot_ref,
                           _block_idx=i) for i in range(num_layers)])
        self.norm_final = nn.LayerNorm(inner, norm_eps, norm_elementwise_affine)
        self.norm_out = AdaLayerNorm(time_embed_dim, 2 * inner, norm_elementwise_affine, norm_eps)
        self.proj_out = nn.Linear(inner, patch_size * patch_size * out_channels)

    def _time(self, emb_mod: nn.Module, timestep: torch.Tensor, inner: int, dtype) -> torch.Tensor:
        half = inner // 2
        exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timestep.device) / half
        e = timestep[:, None].float() * torch.exp(exponent)[None, :]
        return emb_mod(torch.cat([torch.cos(e), torch.sin(e)], dim=-1).to(dtype))  # flip_sin_to_cos=True

    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor, timestep: torch.Tensor, timestep_cond=None, ofs=None,
                image_rotary_emb=None, attention_kwargs: Optional[Dict[str, Any]] = None, return_dict: bool = True, num_mot_ref: int = 1,
                hidden_states_mot_ref: Optional[torch.Tensor] = None, encoder_hidden_states_mot_ref: Optional[torch.Tensor] = None,
                image_rotary_emb_mot_ref=None, effect_types=None, reference_train_mode=None, timestep_list_mot_ref=None):
        cfg = self.config
        B, Fr, C, H, W = hidden_states.shape
        inner = cfg["num_attention_heads"] * cfg["attention_head_dim"]
        Ttok = encoder_hidden_states.shape[-2]
        dt = hidden_states.dtype
        emb = self._time(self.time_embedding, timestep, inner, dt)
        if timestep_list_mot_ref is not None:
            emb_list_r = [self._time(self.time_embedding_mot_ref, ts, inner, dt) for ts in timestep_list_mot_ref]
            emb_r = None
        else:
            emb_r, emb_list_r = self._time(self.time_embedding_mot_ref, timestep, inner, dt), None
        assert hidden_states_mot_ref.shape[1] // Fr == num_mot_ref, f"hidden_states_mot_ref.shape[1]: {hidden_states_mot_ref.shape}"

        sp = ulysses.current()
        h = self.patch_embed(encoder_hidden_states, hidden_states)
        vs, es = [], []
        for i in range(num_mot_ref):
            hi = self.patch_embed_mot_ref(encoder_hidden_states_mot_ref[:, i * Ttok:(i + 1) * Ttok], hidden_states_mot_ref[:, i * Fr:(i + 1) * Fr])
            es.append(hi[:, :Ttok]), vs.append(hi[:, Ttok:])
        rope, rope_r = image_rotary_emb, image_rotary_emb_mot_ref
        if sp is None:
            e, v = h[:, :Ttok], h[:, Ttok:]
            v_r, e_r = torch.cat(vs, dim=1), torch.cat(es, dim=1)
        else:
            # Ulysses: rank r owns rows [r n, (r+1) n) of each stream's [text | video] sequence (17 776 / 8 = 2 222), so
            # rank 0 holds the text rows in front of its video rows and the others hold video rows only; RoPE tables follow the rows
            if num_mot_ref != 1:
                raise NotImplementedError("Ulysses sequence parallelism supports one reference video (num_mot_ref == 1)")
            hd = cfg["attention_head_dim"]
            (e, v, rope), (e_r, v_r, rope_r) = (_shard_stream(t, Ttok, r, hd, sp, cfg["num_attention_heads"])
                                                for t, r in ((h, image_rotary_emb), (torch.cat([es[0], vs[0]], dim=1), image_rotary_emb_mot_ref)))

        for block in self.transformer_blocks:
            v, e, v_r, e_r = block(hidden_states=v, encoder_hidden_states=e, temb=emb, image_rotary_emb=rope,
                                   attention_kwargs=attention_kwargs, hidden_states_mot_ref=v_r, encoder_hidden_states_mot_ref=e_r,
                                   temb_mot_ref=emb_r, temb_list_mot_ref=emb_list_r, image_rotary_emb_mot_ref=rope_r)

        v = ops.adaln_layernorm(v, eps=self.norm_final.eps, rounding=ops.ROUND_COG, ln_w=_f32(self.norm_final, "w", self.norm_final.weight),
                                ln_b=_f32(self.norm_final, "b", self.norm_final.bias))
        # AdaLayerNorm(chunk_dim=1): shift first, then scale (normalization.py:72-77); bf16 tensor arithmetic
        m = F.linear(F.silu(emb), self.norm_out.linear.weight, self.norm_out.linear.bias).view(B, 2, -1).clone()
        m[:, 1] += 1
        m = m.float()
        v = ops.adaln_layernorm(v, eps=self.norm_out.norm.eps, rounding=ops.ROUND_COG, ln_w=_f32(self.norm_out.norm, "w", self.norm_out.norm.weight),
                                ln_b=_f32(self.norm_out.norm, "b", self.norm_out.norm.bias), scale1p=m[:, 1], shift=m[:, 0])
        v = self.proj_out(v)
        if sp is not None:  # all-gather the rank-local rows (rank 0's are preceded by its text rows, padded here) and drop the text rows
            pad = v.new_zeros((B, e.shape[1], v.shape[-1]))
            v = ulysses.gather_rows(torch.cat([pad, v], dim=1), sp)[:, Ttok:]
        p = cfg["patch_size"]
        out = v.reshape(B, Fr, H // p, W // p, -1, p, p).permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)
        if not return_dict:
            return (out,)
        return _Output(sample=out)
