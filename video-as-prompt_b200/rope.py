"""Temporally-biased 3D RoPE tables for the two MoT models, built on the device, in the layout the kernel wants.

The q/k kernel (vap_qk_norm_rope) rotates interleaved pairs (x[2i], x[2i+1]) by compact fp32 tables
``cos, sin : [tokens, head_dim/2]``.  This module
  * builds those tables for Wan from integer (t, h, w) grid positions — target t = 0..F-1, reference-video
    t = -F_ref..-1 (WanRotaryPosEmbed / WanRotaryPosEmbedRef, transformer_wan_mot.py:368-464) — in float64 on the
    device, instead of the reference's per-forward CPU build + H2D copy of a complex128 tensor;
  * converts what reference callers pass into that layout: Wan's complex128 ``rotary_emb`` [1,1,S,D/2]
    (transformer_wan_mot.py:408) and CogVideoX's ``(cos, sin)`` [S, D] repeat-interleaved tables (embeddings.py:1191-1193);
  * restates ``get_3d_rotary_pos_embed`` (embeddings.py:816-949, "linspace" grid, incl. ``mot_num`` /
    ``continous_negative`` / ``discrete_long_reference``) so pipelines and benches can make CogVideoX tables.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import torch

Tables = Tuple[torch.Tensor, torch.Tensor]
_WAN_CACHE: dict = {}


def _axis_angles(dim: int, pos: torch.Tensor, theta: float) -> torch.Tensor:
    """angles[p, i] = pos[p] * theta^(-2i/dim), float64 (get_1d_rotary_pos_embed, embeddings.py:1181-1187)."""
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64, device=pos.device)[: dim // 2] / dim))
    return torch.outer(pos.to(torch.float64), inv)


def wan_rope_tables(head_dim: int, patch_size: Sequence[int], latent_shape: Sequence[int], *, ref: bool, device,
                    max_seq_len: int = 1024, theta: float = 10000.0) -> Tables:
    """(cos, sin) fp32 [ppf*pph*ppw, head_dim/2] for a latent of (frames, height, width).

    Channel split t:h:w = D-4*(D//6) : 2*(D//6) : 2*(D//6) (transformer_wan_mot.py:378-379); token order
    (f, y, x) row-major (:405-408).  ref=True shifts the temporal positions to start at -frames (:437)."""
    key = (head_dim, tuple(patch_size), tuple(latent_shape), bool(ref), str(device), max_seq_len, theta)
    hit = _WAN_CACHE.get(key)
    if hit is not None:
        return hit
    frames, height, width = latent_shape
    p_t, p_h, p_w = patch_size
    ppf, pph, ppw = frames // p_t, height // p_h, width // p_w
    if max(ppf, pph, ppw) > max_seq_len:
        raise ValueError(f"grid {(ppf, pph, ppw)} exceeds rope_max_seq_len {max_seq_len}")
    h_dim = w_dim = 2 * (head_dim // 6)
    t_dim = head_dim - h_dim - w_dim
    t0 = -frames if ref else 0
    at = _axis_angles(t_dim, torch.arange(t0, t0 + ppf, device=device), theta)
    ah = _axis_angles(h_dim, torch.arange(pph, device=device), theta)
    aw = _axis_angles(w_dim, torch.arange(ppw, device=device), theta)
    ang = torch.cat([
        at.view(ppf, 1, 1, -1).expand(ppf, pph, ppw, -1),
        ah.view(1, pph, 1, -1).expand(ppf, pph, ppw, -1),
        aw.view(1, 1, ppw, -1).expand(ppf, pph, ppw, -1),
    ], dim=-1).reshape(ppf * pph * ppw, head_dim // 2)
    out = (ang.cos().float().contiguous(), ang.sin().float().contiguous())
    if len(_WAN_CACHE) > 16:
        _WAN_CACHE.clear()
    _WAN_CACHE[key] = out
    return out


def as_tables(rotary: Union[torch.Tensor, Tables, None], head_dim: int, device) -> Optional[Tables]:
    """Normalise whatever a caller hands a block into compact fp32 (cos, sin) [S, head_dim/2] on `device`.

    Accepts: None; Wan complex freqs [1,1,S,D/2]; CogVideoX (cos, sin) [S, D] with every value repeated twice
    (repeat_interleave(2)); or already compact tables [S, D/2].  The converted tables are cached on the source
    tensor object, because the reference hands the same object to all 40 blocks of a forward."""
    if rotary is None:
        return None
    if isinstance(rotary, torch.Tensor):
        if not rotary.is_complex():
            raise TypeError("a single-tensor rotary embedding must be complex (Wan freqs)")
        cached = getattr(rotary, "_vap_tables", None)
        if cached is None or cached[0].device != torch.device(device):
            fr = rotary.reshape(-1, rotary.shape[-1])
            if fr.shape[-1] != head_dim // 2:
                raise ValueError(f"rotary_emb last dim {fr.shape[-1]} != head_dim/2 = {head_dim // 2}")
            cached = (fr.real.to(device=device, dtype=torch.float32).contiguous(), fr.imag.to(device=device, dtype=torch.float32).contiguous())
            rotary._vap_tables = cached
        return cached
    cos, sin = rotary
    cached = getattr(cos, "_vap_tables", None)
    if cached is not None and cached[0].device == torch.device(device) and cached[2] is sin:
        return cached[0], cached[1]
    if cos.shape[-1] == head_dim:  # repeat-interleaved real tables -> compact
        c, s = cos[..., 0::2], sin[..., 0::2]
    elif cos.shape[-1] == head_dim // 2:
        c, s = cos, sin
    else:
        raise ValueError(f"rotary table last dim {cos.shape[-1]} matches neither head_dim {head_dim} nor head_dim/2")
    c = c.reshape(-1, head_dim // 2).to(device=device, dtype=torch.float32).contiguous()
    s = s.reshape(-1, head_dim // 2).to(device=device, dtype=torch.float32).contiguous()
    try:
        cos._vap_tables = (c, s, sin)
    except Exception:  # pragma: no cover - exotic tensor subclasses
        pass
    return c, s


def get_3d_rotary_pos_embed(embed_dim: int, crops_coords, grid_size, temporal_size: int, theta: float = 10000.0, *,
                            device=None, mot_num: int = 0, ref_type: str = "continous_negative", start_point: int = 50,
                            gap: int = 30) -> Tables:
    """CogVideoX 3D RoPE tables, same signature subset and output format ([T*H*W, embed_dim] cos and sin, values
    repeated pairwise) as the reference's get_3d_rotary_pos_embed(grid_type="linspace") (embeddings.py:816-949).

    mot_num > 0 gives the reference-video stream's tables: "continous_negative" places its frames at
    linspace(-mot_num*T, -1, mot_num*T) (:871-881); "discrete_long_reference" at start_point + gap*i + arange(T) (:886-890)."""
    (s0, s1), (e0, e1) = crops_coords
    gh, gw = grid_size
    f32 = dict(device=device, dtype=torch.float32)
    grid_h = torch.linspace(s0, e0 * (gh - 1) / gh, gh, **f32)
    grid_w = torch.linspace(s1, e1 * (gw - 1) / gw, gw, **f32)
    t_last = temporal_size * (temporal_size - 1) / temporal_size
    nt = temporal_size
    if mot_num <= 0:
        grid_t = torch.linspace(0, t_last, temporal_size, **f32)
    elif ref_type == "continous_negative":
        nt = temporal_size * mot_num
        grid_t = torch.linspace(-mot_num * (t_last + 1), -1, nt, **f32)
    elif ref_type == "discrete_long_reference":
        offs = start_point + torch.arange(mot_num, **f32) * gap
        grid_t = (offs.unsqueeze(1) + torch.arange(temporal_size, **f32)).flatten()
        if mot_num != 1:
            raise ValueError("discrete_long_reference only broadcasts for mot_num == 1 in the reference (embeddings.py:886-890, 932-935)")
    else:
        raise ValueError(f"Invalid {ref_type} passed for `ref_type`.")
    dim_t, dim_h, dim_w = embed_dim // 4, embed_dim // 8 * 3, embed_dim // 8 * 3

    def axis(dim, pos):  # fp32 angles like the reference (freqs_dtype float32)
        inv = 1.0 / (theta ** (torch.arange(0, dim, 2, **f32)[: dim // 2] / dim))
        a = torch.outer(pos, inv)
        return a.cos().repeat_interleave(2, dim=1), a.sin().repeat_interleave(2, dim=1)

    (tc, ts), (hc, hs), (wc, ws) = axis(dim_t, grid_t), axis(dim_h, grid_h), axis(dim_w, grid_w)

    def combine(a, b, c):
        return torch.cat([a[:, None, None, :].expand(-1, gh, gw, -1), b[None, :, None, :].expand(nt, -1, gw, -1),
                          c[None, None, :, :].expand(nt, gh, -1, -1)], dim=-1).reshape(nt * gh * gw, -1)

    return combine(tc, hc, wc), combine(ts, hs, ws)
