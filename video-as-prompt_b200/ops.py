"""Tensor-level wrappers over the C ABI (libvap_b200.so).  PyTorch only provides device memory and the stream.

Every function validates dtype / device / layout, allocates the output with torch and launches the sm_100a kernel
on ``torch.cuda.current_stream()``.  Nothing here computes on the CPU; a CPU tensor raises ``VapError``.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import VapError

EPI_BIAS = 0
EPI_BIAS_GELU = 1
EPI_GATE_RES_F32 = 2
EPI_RES_ADD = 3
EPI_GATE_RES_BF16 = 4

ROUND_WAN = 0
ROUND_COG = 1
QK_WAN = 0
QK_COG = 1


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda_bf16(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if t.requires_grad and torch.is_grad_enabled():
        # the kernels are forward-only and their outputs carry no grad_fn: refusing is better than silently cutting the graph of a trainer
        raise VapError(f"{name} requires grad inside a grad-enabled region: the VAP kernels are inference-only (no backward yet); "
                       "call under torch.no_grad() / torch.inference_mode()")
    if not t.is_cuda:
        raise VapError(f"{name} is on {t.device}: the VAP kernels are CUDA-only (sm_100a); there is no CPU fallback")
    if t.dtype != torch.bfloat16:
        raise TypeError(f"{name} must be torch.bfloat16, got {t.dtype}")
    if t.device.index != torch.cuda.current_device():
        # the launch goes to the CURRENT device's stream: a process that drives several GPUs (e.g. an accelerate device_map) must select the
        # tensor's device first, as torch's own kernels do internally
        raise VapError(f"{name} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                       f"wrap the call in `with torch.cuda.device({t.device.index}):`")


def _need_cuda_float(t: torch.Tensor, name: str) -> None:
    """A contiguous CUDA tensor that is either bf16 or fp32 (parameters a from_pretrained model keeps in fp32)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise VapError(f"{name} must be a CUDA tensor: the VAP kernels are CUDA-only (sm_100a); there is no CPU fallback")
    if t.dtype not in (torch.float32, torch.bfloat16) or not t.is_contiguous():
        raise TypeError(f"{name} must be a contiguous float32 or bfloat16 tensor, got {t.dtype}")


def _need_cuda_f32(t: Optional[torch.Tensor], name: str) -> int:
    if t is None:
        return 0
    if not t.is_cuda or t.dtype != torch.float32:
        raise TypeError(f"{name} must be a CUDA float32 tensor, got {t.dtype} on {t.device}")
    if t.stride(-1) != 1:
        raise ValueError(f"{name} must be contiguous in its last dimension")
    return t.data_ptr()


def _rows_view(t: torch.Tensor, name: str) -> Tuple[int, int, int]:
    """(rows, d, row_stride) of a tensor whose leading dims collapse to uniformly strided rows."""
    if t.stride(-1) != 1:
        raise ValueError(f"{name} must be contiguous in its last dimension")
    d = t.shape[-1]
    if t.dim() == 1:
        return 1, d, d
    if t.numel() == 0:  # no rows (a sequence-parallel rank holding text rows only): nothing to address, any pitch will do
        return 0, d, max(d, 1)
    rs = t.stride(-2)
    rows = t.shape[-2]
    for i in range(t.dim() - 3, -1, -1):  # leading dims must continue the same row pitch
        if t.shape[i] != 1 and t.stride(i) != rows * rs:
            raise ValueError(f"{name} leading dimensions are not uniformly strided rows: shape {tuple(t.shape)} strides {t.stride()}")
        rows *= t.shape[i]
    return rows, d, rs


def _mod_stride(t: Optional[torch.Tensor], d: int, name: str) -> int:
    """Modulation vectors come as [nbatch, d] or [nbatch, 1, d] (possibly a chunk view): return the batch stride."""
    if t is None:
        return 0
    if t.shape[-1] != d:
        raise ValueError(f"{name} last dim {t.shape[-1]} != {d}")
    lead = [i for i in range(t.dim() - 1) if t.shape[i] != 1]
    if len(lead) == 0:
        return 0
    if len(lead) > 1:
        raise ValueError(f"{name} must have at most one non-singleton leading dim, got {tuple(t.shape)}")
    return t.stride(lead[0])


def adaln_layernorm(x: torch.Tensor, *, eps: float, rounding: int, ln_w: Optional[torch.Tensor] = None,
                    ln_b: Optional[torch.Tensor] = None, scale1p: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
                    rows_per_batch: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(adaLN-modulated) LayerNorm, see vap_adaln_layernorm in include/vap_b200.h.  x: [..., rows, d] bf16."""
    _need_cuda_bf16(x, "x")
    rows, d, xs = _rows_view(x, "x")
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _need_cuda_bf16(out, "out")
    orows, od, os_ = _rows_view(out, "out")
    if (orows, od) != (rows, d):
        raise ValueError(f"out shape {tuple(out.shape)} does not match x {tuple(x.shape)}")
    ms = _mod_stride(scale1p, d, "scale1p")
    if shift is not None and _mod_stride(shift, d, "shift") != ms:
        raise ValueError("scale1p and shift must have the same batch stride")
    if rows_per_batch is None:
        rows_per_batch = x.shape[-2] if (x.dim() >= 3 and ms != 0) else rows
    if rows == 0:  # e.g. a sequence-parallel rank that holds text rows only: nothing to launch (an empty tensor has no storage)
        return out
    lib = _lib.load()
    rc = lib.vap_adaln_layernorm(x.data_ptr(), out.data_ptr(), rows, d, xs, os_, _need_cuda_f32(ln_w, "ln_w"), _need_cuda_f32(ln_b, "ln_b"),
                                 _need_cuda_f32(scale1p, "scale1p"), _need_cuda_f32(shift, "shift"), ms, max(int(rows_per_batch), 1),
                                 float(eps), int(rounding), _stream())
    _lib.check(rc, "vap_adaln_layernorm")
    return out


def qk_norm_rope_(q: torch.Tensor, k: Optional[torch.Tensor], *, heads: int, head_dim: int, wq: torch.Tensor,
                  wk: Optional[torch.Tensor] = None, bq: Optional[torch.Tensor] = None, bk: Optional[torch.Tensor] = None, cos: Optional[torch.Tensor] = None,
                  sin: Optional[torch.Tensor] = None, rows_per_batch: int, rope_row0: int = 0, eps: float, mode: int) -> None:
    """In-place q/k normalisation + RoPE, see vap_qk_norm_rope.  q, k: [..., rows, heads*head_dim] views with equal row stride."""
    _need_cuda_bf16(q, "q")
    rows, d, rs = _rows_view(q, "q")
    if d != heads * head_dim:
        raise ValueError(f"q last dim {d} != heads*head_dim {heads * head_dim}")
    if k is not None:
        _need_cuda_bf16(k, "k")
        if _rows_view(k, "k") != (rows, d, rs):
            raise ValueError(f"q {tuple(q.shape)}/{q.stride()} and k {tuple(k.shape)}/{k.stride()} must share shape and row stride")
    rope_rows = 0
    if cos is not None:
        if cos.shape != sin.shape or cos.dim() != 2 or cos.shape[1] != head_dim // 2 or not cos.is_contiguous() or not sin.is_contiguous():
            raise ValueError(f"cos/sin must be contiguous [rope_rows, {head_dim // 2}] tables, got {tuple(cos.shape)}")
        rope_rows = cos.shape[0]
        if rows_per_batch - rope_row0 > rope_rows:
            raise ValueError(f"RoPE table has {rope_rows} rows but {rows_per_batch - rope_row0} tokens per batch need rotating")
    if rows == 0:
        return
    lib = _lib.load()
    rc = lib.vap_qk_norm_rope(q.data_ptr(), k.data_ptr() if k is not None else 0, rows, heads, head_dim, rs, _need_cuda_f32(wq, "wq"), _need_cuda_f32(bq, "bq"),
                              _need_cuda_f32(wk, "wk"), _need_cuda_f32(bk, "bk"), _need_cuda_f32(cos, "cos"), _need_cuda_f32(sin, "sin"),
                              int(rows_per_batch), int(rope_row0), int(rope_rows), float(eps), int(mode), _stream())
    _lib.check(rc, "vap_qk_norm_rope")


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, scale: Optional[float] = None, out: Optional[torch.Tensor] = None,
              return_lse: bool = False, accumulate: bool = False):
    """Non-causal, unmasked attention.  q [B,H,Lq,D], k/v [B,H,Lkv,D] (any batch/head/token strides, D contiguous).

    Returns O with logical shape [B,H,Lq,D] laid out token-major ([B,Lq,H,D] memory), so that
    ``o.transpose(1, 2).flatten(2, 3)`` — what every reference processor does next — is a free view.
    accumulate=True (needs `out`): the result is ADDED to what `out` holds, as a bf16 tensor add (vap_attention_fwd_accumulate)."""
    if accumulate and out is None:
        raise ValueError("accumulate=True adds to an existing output: pass out=")
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _need_cuda_bf16(t, n)
        if t.dim() != 4 or t.stride(-1) != 1:
            raise ValueError(f"{n} must be [B,H,L,D] with contiguous D, got shape {tuple(t.shape)} strides {t.stride()}")
    B, H, Lq, D = q.shape
    Lkv = k.shape[2]
    if k.shape != (B, H, Lkv, D) or v.shape != (B, H, Lkv, D):
        raise ValueError(f"q {tuple(q.shape)}, k {tuple(k.shape)}, v {tuple(v.shape)} are inconsistent")
    if not return_lse and not accumulate and Lq > 0:
        splits = _auto_kv_splits(B, H, Lq, Lkv)
        if splits > 1:  # the work items are a poor multiple of the SM count: cut the KV sequence and merge (see attention_kv_splits)
            return attention_splitkv(q, k, v, splits, scale=scale, out=out)
    if out is None:
        out = torch.empty((B, Lq, H, D), dtype=torch.bfloat16, device=q.device).transpose(1, 2)
    _need_cuda_bf16(out, "out")
    if out.shape != (B, H, Lq, D) or out.stride(-1) != 1:
        raise ValueError(f"out must be [B,H,Lq,D] with contiguous D, got {tuple(out.shape)}")
    lse = torch.empty((B, H, Lq), dtype=torch.float32, device=q.device) if return_lse else None
    if scale is None:
        scale = D ** -0.5
    lib = _lib.load()
    fn = lib.vap_attention_fwd_accumulate if accumulate else lib.vap_attention_fwd
    rc = fn(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr() if lse is not None else 0, B, H, Lq,
            Lkv, D, q.stride(0), q.stride(1), q.stride(2), k.stride(0), k.stride(1), k.stride(2), v.stride(0),
            v.stride(1), v.stride(2), out.stride(0), out.stride(1), out.stride(2), float(scale), _stream())
    _lib.check(rc, "vap_attention_fwd_accumulate" if accumulate else "vap_attention_fwd")
    return (out, lse) if return_lse else out


def attention_bwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, lse: torch.Tensor, dout: torch.Tensor, *,
                  scale: Optional[float] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Gradients of `attention` (vap_attention_bwd): q / o / dout [B,H,Lq,D], k / v [B,H,Lkv,D] (any batch/head/token strides, D
    contiguous), lse [B,H,Lq] fp32 from ``attention(..., return_lse=True)``.  Returns dq, dk, dv with the logical shapes of q, k, v,
    laid out token-major like the forward's output.  Parity on a B200: tests/gpu_checks.py attn_bwd_* (max-abs relative 2-5e-3 per gradient)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (o, "o"), (dout, "dout")):
        _need_cuda_bf16(t, n)
        if t.dim() != 4 or t.stride(-1) != 1:
            raise ValueError(f"{n} must be [B,H,L,D] with contiguous D, got shape {tuple(t.shape)} strides {t.stride()}")
    B, H, Lq, D = q.shape
    Lkv = k.shape[2]
    if k.shape != (B, H, Lkv, D) or v.shape != (B, H, Lkv, D) or o.shape != q.shape or dout.shape != q.shape:
        raise ValueError(f"q {tuple(q.shape)}, k {tuple(k.shape)}, v {tuple(v.shape)}, o {tuple(o.shape)}, dout {tuple(dout.shape)} are inconsistent")
    if lse.shape != (B, H, Lq) or not lse.is_contiguous():
        raise ValueError(f"lse must be a contiguous [B,H,Lq] tensor, got {tuple(lse.shape)}")
    lse_ptr = _need_cuda_f32(lse, "lse")
    if scale is None:
        scale = D ** -0.5
    dq = torch.empty((B, Lq, H, D), dtype=torch.bfloat16, device=q.device).transpose(1, 2)
    dk = torch.empty((B, Lkv, H, D), dtype=torch.bfloat16, device=q.device).transpose(1, 2)
    dv = torch.empty((B, Lkv, H, D), dtype=torch.bfloat16, device=q.device).transpose(1, 2)
    if B * H * Lq * Lkv == 0:
        return dq.zero_(), dk.zero_(), dv.zero_()
    delta = torch.empty((B, H, Lq), dtype=torch.float32, device=q.device)
    strides = (ctypes.c_int64 * 24)(*[st for t in (q, k, v, o, dout, dq, dk, dv) for st in t.stride()[:3]])
    rc = _lib.load().vap_attention_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), dout.data_ptr(), lse_ptr, dq.data_ptr(), dk.data_ptr(),
                                       dv.data_ptr(), delta.data_ptr(), B, H, Lq, Lkv, D, strides, float(scale), _stream())
    _lib.check(rc, "vap_attention_bwd")
    return dq, dk, dv


_SPLIT_CACHE = {}


def _auto_kv_splits(B: int, H: int, Lq: int, Lkv: int) -> int:
    """Split count `attention` / `attention_scatter` use by themselves: VAP_ATTN_SPLITKV = "auto" (default: the wave model of
    attention_kv_splits), "0" / "1" (never split) or a number in [2, 8] (always split, for testing)."""
    key = (B, H, Lq, Lkv)
    s = _SPLIT_CACHE.get(key)
    if s is None:
        mode = os.environ.get("VAP_ATTN_SPLITKV", "auto")
        if mode == "auto":
            s = attention_kv_splits(B, H, Lq, Lkv)
        else:
            s = max(1, min(int(mode), 8, (Lkv + 127) // 128))
        _SPLIT_CACHE[key] = s
    return s


def attention_kv_splits(B: int, H: int, Lq: int, Lkv: int, min_gain: float = 0.03) -> int:
    """How many KV ranges to cut the attention into so that the (batch, head, 256 query rows) work items fill the SMs.
    The kernel runs one item per SM at a time, so T ~ ceil(items / SMs); with s ranges T ~ ceil(s * items / SMs) / s plus ~1 % for
    the extra prologues and the merge.  Returns 1 unless a split gains at least `min_gain` (e.g. 5 heads x 159 blocks = 795 items =
    5.37 waves on 148 SMs -> 6; two ranges: 10.74 -> 11 half-waves = 5.5)."""
    items = B * H * ((Lq + 255) // 256)
    sms = sm_count()
    kv_tiles = (Lkv + 127) // 128
    best, best_t = 1, float(-(-items // sms))
    for s in (2, 3, 4):
        if kv_tiles < 16 * s:  # keep every range long enough to amortise its prologue / epilogue
            break
        t = -(-(items * s) // sms) / s * 1.01
        if t < best_t * (1.0 - min_gain):
            best, best_t = s, t
    return best


def attention_splitkv(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, kv_splits: int, *, scale: Optional[float] = None,
                      out: Optional[torch.Tensor] = None, o_ptrs=None, rows_per_peer: int = 0, o_strides=None) -> Optional[torch.Tensor]:
    """Attention with the KV sequence cut into `kv_splits` ranges (vap_attention_fwd_splitkv) followed by the merge kernel
    (vap_attention_combine).  The merged result goes to `out` / a fresh token-major tensor like `attention`, or — with o_ptrs /
    rows_per_peer / o_strides as in `attention_scatter` — straight into the owning ranks' buffers."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _need_cuda_bf16(t, n)
        if t.dim() != 4 or t.stride(-1) != 1:
            raise ValueError(f"{n} must be [B,H,L,D] with contiguous D, got shape {tuple(t.shape)} strides {t.stride()}")
    B, H, Lq, D = q.shape
    Lkv = k.shape[2]
    if k.shape != (B, H, Lkv, D) or v.shape != (B, H, Lkv, D):
        raise ValueError(f"q {tuple(q.shape)}, k {tuple(k.shape)}, v {tuple(v.shape)} are inconsistent")
    if not 1 <= kv_splits <= 8:
        raise ValueError(f"kv_splits={kv_splits} must be in [1, 8]")
    if scale is None:
        scale = D ** -0.5
    o_part = torch.empty((kv_splits, B, Lq, H, D), dtype=torch.bfloat16, device=q.device)
    lse_part = torch.empty((kv_splits, B, H, Lq), dtype=torch.float32, device=q.device)
    lib = _lib.load()
    rc = lib.vap_attention_fwd_splitkv(q.data_ptr(), k.data_ptr(), v.data_ptr(), o_part.data_ptr(), lse_part.data_ptr(), int(kv_splits), B, H, Lq,
                                       Lkv, D, q.stride(0), q.stride(1), q.stride(2), k.stride(0), k.stride(1), k.stride(2), v.stride(0),
                                       v.stride(1), v.stride(2), float(scale), _stream())
    _lib.check(rc, "vap_attention_fwd_splitkv")
    if o_ptrs is not None:
        if len(o_ptrs) * rows_per_peer < Lq:
            raise ValueError(f"{len(o_ptrs)} peers x {rows_per_peer} rows do not cover Lq={Lq}")
        table = _ptr_table(o_ptrs)
        rc = lib.vap_attention_combine(o_part.data_ptr(), lse_part.data_ptr(), int(kv_splits), B, H, Lq, D, 0, table, len(o_ptrs), int(rows_per_peer), 0,
                                       int(o_strides[0]), int(o_strides[1]), int(o_strides[2]), _stream())
        _lib.check(rc, "vap_attention_combine")
        return None
    if out is None:
        out = torch.empty((B, Lq, H, D), dtype=torch.bfloat16, device=q.device).transpose(1, 2)
    _need_cuda_bf16(out, "out")
    if out.shape != (B, H, Lq, D) or out.stride(-1) != 1:
        raise ValueError(f"out must be [B,H,Lq,D] with contiguous D, got {tuple(out.shape)}")
    rc = lib.vap_attention_combine(o_part.data_ptr(), lse_part.data_ptr(), int(kv_splits), B, H, Lq, D, out.data_ptr(), 0, 0, 0, 0, out.stride(0),
                                   out.stride(1), out.stride(2), _stream())
    _lib.check(rc, "vap_attention_combine")
    return out


def _ptr_table(ptrs):
    """Host array of device pointers (void* const*) for the peer-table entry points."""
    arr = (ctypes.c_void_p * len(ptrs))(*[int(p) for p in ptrs])
    return arr


def qkv_scatter(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, heads: int, head_dim: int, wq: torch.Tensor, wk: torch.Tensor,
                bq: Optional[torch.Tensor] = None, bk: Optional[torch.Tensor] = None, cos: Optional[torch.Tensor] = None,
                sin: Optional[torch.Tensor] = None, rows_per_batch: int, rope_row0: int = 0, eps: float, mode: int, dst_ptrs, dst_slot: int,
                slot_rows: int, dst_row0: int) -> None:
    """q/k normalisation + RoPE + the Ulysses all-to-all #1 dispatch in ONE kernel (see vap_qkv_scatter): q, k, v are column
    views [rows, heads*head_dim] of the local QKV projection (left untouched); the results go to the receive buffers of the
    len(dst_ptrs) ranks (device pointers, peer memory), layout [slot, slot_rows, 3, heads/P*head_dim] each."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _need_cuda_bf16(t, n)
    rows, d, rs = _rows_view(q, "q")
    if d != heads * head_dim or _rows_view(k, "k") != (rows, d, rs) or _rows_view(v, "v") != (rows, d, rs):
        raise ValueError("q, k and v must be [rows, heads*head_dim] views with one common row stride")
    rope_rows = 0
    if cos is not None:
        if cos.shape != sin.shape or cos.dim() != 2 or cos.shape[1] != head_dim // 2 or not cos.is_contiguous() or not sin.is_contiguous():
            raise ValueError(f"cos/sin must be contiguous [rope_rows, {head_dim // 2}] tables, got {tuple(cos.shape)}")
        rope_rows = cos.shape[0]
        if rows_per_batch - rope_row0 > rope_rows:
            raise ValueError(f"RoPE table has {rope_rows} rows but {rows_per_batch - rope_row0} tokens per batch need rotating")
    if rows == 0:
        return
    lib = _lib.load()
    table = _ptr_table(dst_ptrs)
    rc = lib.vap_qkv_scatter(q.data_ptr(), k.data_ptr(), v.data_ptr(), rows, heads, head_dim, rs, _need_cuda_f32(wq, "wq"), _need_cuda_f32(bq, "bq"),
                             _need_cuda_f32(wk, "wk"), _need_cuda_f32(bk, "bk"), _need_cuda_f32(cos, "cos"), _need_cuda_f32(sin, "sin"),
                             int(rows_per_batch), int(rope_row0), int(rope_rows), float(eps), int(mode), table, len(dst_ptrs), int(dst_slot),
                             int(slot_rows), int(dst_row0), _stream())
    _lib.check(rc, "vap_qkv_scatter")


def attention_scatter(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, o_ptrs, rows_per_peer: int, o_strides, scale: Optional[float] = None) -> None:
    """Attention whose epilogue stores query row r into peer r // rows_per_peer's output buffer (device pointers o_ptrs, element
    strides o_strides = (batch, head, row) inside each peer buffer) — the Ulysses all-to-all #2 fused into the kernel."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _need_cuda_bf16(t, n)
        if t.dim() != 4 or t.stride(-1) != 1:
            raise ValueError(f"{n} must be [B,H,L,D] with contiguous D, got shape {tuple(t.shape)} strides {t.stride()}")
    B, H, Lq, D = q.shape
    Lkv = k.shape[2]
    if k.shape != (B, H, Lkv, D) or v.shape != (B, H, Lkv, D):
        raise ValueError(f"q {tuple(q.shape)}, k {tuple(k.shape)}, v {tuple(v.shape)} are inconsistent")
    if len(o_ptrs) * rows_per_peer < Lq:
        raise ValueError(f"{len(o_ptrs)} peers x {rows_per_peer} rows do not cover Lq={Lq}")
    splits = _auto_kv_splits(B, H, Lq, Lkv)
    if splits > 1:  # e.g. 5 heads per rank under 8-way Ulysses: the merge kernel does the peer stores
        attention_splitkv(q, k, v, splits, scale=scale, o_ptrs=o_ptrs, rows_per_peer=rows_per_peer, o_strides=o_strides)
        return
    if scale is None:
        scale = D ** -0.5
    lib = _lib.load()
    table = _ptr_table(o_ptrs)
    rc = lib.vap_attention_fwd_scatter(q.data_ptr(), k.data_ptr(), v.data_ptr(), table, len(o_ptrs), int(rows_per_peer), 0, B, H, Lq, Lkv, D,
                                       q.stride(0), q.stride(1), q.stride(2), k.stride(0), k.stride(1), k.stride(2), v.stride(0), v.stride(1),
                                       v.stride(2), int(o_strides[0]), int(o_strides[1]), int(o_strides[2]), float(scale), _stream())
    _lib.check(rc, "vap_attention_fwd_scatter")


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *, epilogue: int = EPI_BIAS,
           residual: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None, rows_per_batch: Optional[int] = None,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = epilogue(x @ weight.T + bias), see vap_gemm_bf16.  x [..., M, K], weight [N, K] (nn.Linear layout)."""
    _need_cuda_bf16(x, "x"), _need_cuda_bf16(weight, "weight")
    if x.dim() == 3 and x.shape[0] > 1 and (x.stride(0) != x.shape[1] * x.stride(1) or
                                            (out is not None and out.stride(0) != out.shape[1] * out.stride(1)) or
                                            (residual is not None and residual.stride(0) != residual.shape[1] * residual.stride(1))):
        # batches are row-slices of a larger (joint) buffer: one launch per batch element
        if out is None:
            out = torch.empty(x.shape[:-1] + (weight.shape[0],), dtype=torch.bfloat16, device=x.device)
        for b in range(x.shape[0]):
            gb = None
            if gate is not None:
                gb = gate.reshape(-1, gate.shape[-1]) if gate.dim() != 3 else gate[:, 0]
                gb = gb[b if gb.shape[0] > 1 else 0]
            linear(x[b], weight, bias, epilogue=epilogue, residual=None if residual is None else residual[b], gate=gb, out=out[b])
        return out
    M, K, lda = _rows_view(x, "x")
    if weight.dim() != 2 or weight.shape[1] != K or weight.stride(1) != 1:
        raise ValueError(f"weight must be [N, {K}] with contiguous rows, got {tuple(weight.shape)}")
    N = weight.shape[0]
    if out is None:
        out = torch.empty(x.shape[:-1] + (N,), dtype=torch.bfloat16, device=x.device)
    _need_cuda_bf16(out, "out")
    oM, oN, ldc = _rows_view(out, "out")
    if (oM, oN) != (M, N):
        raise ValueError(f"out shape {tuple(out.shape)} does not match [{M}, {N}]")
    if bias is not None:
        _need_cuda_bf16(bias, "bias")
        if bias.shape != (N,) or not bias.is_contiguous():
            raise ValueError(f"bias must be a contiguous [{N}] vector")
    r_ptr, ldr = 0, 0
    if residual is not None:
        _need_cuda_bf16(residual, "residual")
        rM, rN, ldr = _rows_view(residual, "residual")
        if (rM, rN) != (M, N):
            raise ValueError(f"residual shape {tuple(residual.shape)} does not match [{M}, {N}]")
        r_ptr = residual.data_ptr()
    gs = _mod_stride(gate, N, "gate") if gate is not None else 0
    if rows_per_batch is None:
        rows_per_batch = x.shape[-2] if (x.dim() >= 3 and gs != 0) else max(M, 1)
    if M == 0:
        return out
    lib = _lib.load()
    rc = lib.vap_gemm_bf16(x.data_ptr(), lda, weight.data_ptr(), weight.stride(0), out.data_ptr(), ldc, M, N, K,
                           bias.data_ptr() if bias is not None else 0, int(epilogue), r_ptr, ldr, _need_cuda_f32(gate, "gate"), gs,
                           max(int(rows_per_batch), 1), _stream())
    _lib.check(rc, "vap_gemm_bf16")
    return out


def cfg_flow_match_step(noise_cond: torch.Tensor, noise_uncond: Optional[torch.Tensor], sample: torch.Tensor, *, guidance_scale: float, dt: float,
                        out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Classifier-free guidance + FlowMatchEuler update in one pass (vap_cfg_flow_match_step):
    ``(sample.float() + dt * (u + g * (c - u))).to(bf16)`` with the reference's bf16 rounding points.  noise_cond / noise_uncond
    [B, ...] bf16 contiguous (noise_uncond None: no guidance), sample same shape, fp32 or bf16; `out` (optional) may be a channel
    slice of the next step's transformer input ([B, C_all, ...][:, :C]): every batch element must be one contiguous block."""
    _need_cuda_bf16(noise_cond, "noise_cond")
    if not noise_cond.is_contiguous() or noise_cond.dim() < 1:
        raise ValueError("noise_cond must be contiguous")
    if noise_uncond is not None:
        _need_cuda_bf16(noise_uncond, "noise_uncond")
        if noise_uncond.shape != noise_cond.shape or not noise_uncond.is_contiguous():
            raise ValueError("noise_uncond must be contiguous with the shape of noise_cond")
    if not isinstance(sample, torch.Tensor) or not sample.is_cuda:
        raise VapError("sample must be a CUDA tensor: the VAP kernels are CUDA-only (sm_100a); there is no CPU fallback")
    if sample.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"sample must be float32 or bfloat16, got {sample.dtype}")
    if sample.shape != noise_cond.shape or not sample.is_contiguous():
        raise ValueError("sample must be contiguous with the shape of noise_cond")
    B = noise_cond.shape[0]
    inner = noise_cond.numel() // max(B, 1)
    if inner % 8:
        raise ValueError(f"elements per batch ({inner}) must be a multiple of 8")
    if out is None:
        out = torch.empty(noise_cond.shape, dtype=torch.bfloat16, device=noise_cond.device)
    _need_cuda_bf16(out, "out")
    if out.shape != noise_cond.shape or (B > 0 and not out[0].is_contiguous()):
        raise ValueError("out must have the shape of noise_cond with every batch element contiguous")
    obs = out.stride(0) if B > 1 else max(inner, 8)
    if B * inner == 0:
        return out
    rc = _lib.load().vap_cfg_flow_match_step(noise_cond.data_ptr(), noise_uncond.data_ptr() if noise_uncond is not None else 0, sample.data_ptr(),
                                             int(sample.dtype == torch.float32), out.data_ptr(), B, inner, obs, float(guidance_scale), float(dt), _stream())
    _lib.check(rc, "vap_cfg_flow_match_step")
    return out


def wan_modulation(table: torch.Tensor, temb: torch.Tensor, plus_one_mask: int = 0b010010) -> torch.Tensor:
    """(table[1, C, d] + temb[B, C, d].float()) with 1 added to the chunks in `plus_one_mask`, fp32 [B, C, d] (vap_wan_modulation): the six
    adaLN vectors of a Wan block — shift, 1 + scale, gate, c_shift, 1 + c_scale, c_gate — in one launch."""
    _need_cuda_float(table, "table"), _need_cuda_float(temb, "temb")
    C, d = table.shape[-2], table.shape[-1]
    if table.numel() != C * d or temb.dim() != 3 or temb.shape[1:] != (C, d):
        raise ValueError(f"table {tuple(table.shape)} / temb {tuple(temb.shape)}: expected [1, C, d] and [B, C, d]")
    out = torch.empty(temb.shape, dtype=torch.float32, device=temb.device)
    rc = _lib.load().vap_wan_modulation(table.data_ptr(), int(table.dtype == torch.float32), temb.data_ptr(), int(temb.dtype == torch.float32), out.data_ptr(),
                                        temb.shape[0], C, d, int(plus_one_mask), _stream())
    _lib.check(rc, "vap_wan_modulation")
    return out


def ulysses_pack(src: torch.Tensor, nsplit: int, out: torch.Tensor) -> torch.Tensor:
    """out[s, l, :] = src[l, s*chunk:(s+1)*chunk].  src [L, nsplit*chunk] (row-strided view); out [nsplit, L, chunk] view whose
    rows / splits may be strided (e.g. the q, k or v slot of the all-to-all send buffer [P, L, 3, chunk])."""
    _need_cuda_bf16(src, "src"), _need_cuda_bf16(out, "out")
    L, width, rs = _rows_view(src, "src")
    if width % nsplit:
        raise ValueError(f"row width {width} is not divisible by nsplit {nsplit}")
    chunk = width // nsplit
    if out.shape != (nsplit, L, chunk) or out.stride(2) != 1:
        raise ValueError(f"out must be [{nsplit}, {L}, {chunk}] with contiguous last dim, got {tuple(out.shape)}")
    _lib.check(_lib.load().vap_ulysses_pack(src.data_ptr(), out.data_ptr(), L, nsplit, chunk, rs, out.stride(1), out.stride(0), _stream()),
               "vap_ulysses_pack")
    return out


def ulysses_unpack(src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[l, s*chunk:(s+1)*chunk] = src[s, l, :].  src [nsplit, L, chunk] (strided view ok) -> out [L, nsplit*chunk]."""
    _need_cuda_bf16(src, "src")
    if src.dim() != 3 or src.stride(2) != 1:
        raise ValueError("src must be [nsplit, L, chunk] with contiguous last dim")
    nsplit, L, chunk = src.shape
    if out is None:
        out = torch.empty((L, nsplit * chunk), dtype=torch.bfloat16, device=src.device)
    _need_cuda_bf16(out, "out")
    oL, ow, rs = _rows_view(out, "out")
    if (oL, ow) != (L, nsplit * chunk):
        raise ValueError(f"out shape {tuple(out.shape)} does not match [{L}, {nsplit * chunk}]")
    _lib.check(_lib.load().vap_ulysses_unpack(src.data_ptr(), out.data_ptr(), L, nsplit, chunk, src.stride(1), src.stride(0), rs, _stream()),
               "vap_ulysses_unpack")
    return out


def probe_umma(a: torch.Tensor, b: torch.Tensor, *, a_in_tmem: bool, b_mn_major: bool, lbo_b: int = -1, sbo_b: int = -1,
               kstep_b: int = -1, layout_type: int = -1, lane16_shapes: bool = False) -> torch.Tensor:
    """Bring-up probe (vap_probe_umma): D[128,N] fp32 = A[128,K] @ (B[N,K].T if not b_mn_major else B[K,N])."""
    _need_cuda_bf16(a, "a"), _need_cuda_bf16(b, "b")
    K = a.shape[1]
    N = b.shape[1] if b_mn_major else b.shape[0]
    if a.shape != (128, K) or not a.is_contiguous() or not b.is_contiguous():
        raise ValueError("probe_umma: a must be contiguous [128, K], b contiguous")
    d = torch.empty((128, N), dtype=torch.float32, device=a.device)
    rc = _lib.load().vap_probe_umma(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, int(a_in_tmem) | (int(lane16_shapes) << 1), int(b_mn_major), lbo_b, sbo_b, kstep_b,
                                    layout_type, _stream())
    _lib.check(rc, "vap_probe_umma")
    return d


def sm_count() -> int:
    n = _lib.load().vap_sm_count()
    if n < 0:
        _lib.check(n, "vap_sm_count")
    return n
