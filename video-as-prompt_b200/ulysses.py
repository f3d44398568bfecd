"""Ulysses sequence parallelism for the MoT joint attention (one process per GPU, NCCL all-to-all over NVLink/NVSwitch).

The reference scales long sequences with ring attention (finetrainers/models/attention_dispatch.py:686-773, hooks in
finetrainers/parallel/ptd.py:515-679); everything in a MoT block except the joint attention is token-local, so here both
streams are sharded along tokens (rank r owns rows [r*L/P, (r+1)*L/P) of each stream) and each block does exactly two
exchanges: all-to-all #1 turns the local rows x all heads of q|k|v into all rows x H/P heads, the attention kernel runs on
H/P heads over the full joint sequence, all-to-all #2 brings O back to local rows x all heads.

After exchange #1 the joint sequence is ordered rank-major ([tgt_0|ref_0|tgt_1|ref_1|...]) — a permutation of the
reference's [target|ref] order, which is irrelevant for unmasked attention as long as O rows map back to their tokens,
which exchange #2 does by construction.

Two transports:
  mode "p2p" (default on CUDA): both exchanges are FUSED into the producing kernels over NVLink peer memory (torch symmetric
      memory gives every rank the device pointers of all ranks' buffers).  The q/k-norm + RoPE kernel stores its results — and a
      copy of V — straight into the receive buffer of the rank that owns the head (vap_qkv_scatter: no local write, no pack
      kernel, no collective call), and the attention kernel's epilogue stores every O row straight into the output buffer of
      the rank that owns the row (vap_attention_fwd_scatter: no all-to-all, no unpack).  Ordering between ranks: one 7 us
      device-side barrier per exchange and two alternating buffer sets (see PeerExchange).
  mode "nccl": pack kernel -> dist.all_to_all_single -> attention -> all_to_all_single -> unpack kernel (the baseline; also the
      path the CPU/gloo tests drive with injected pack/unpack functions).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import ops


@dataclass
class SequenceParallel:
    group: Optional[dist.ProcessGroup]
    rank: int
    world: int
    mode: str = "nccl"  # "p2p" | "nccl"
    _exchanges: Optional[dict] = None


_CURRENT: Optional[SequenceParallel] = None


def enable(group: Optional[dist.ProcessGroup] = None, mode: Optional[str] = None) -> SequenceParallel:
    """Activate Ulysses over `group` (default: the world group).  world == 1 is a no-op context.
    mode: "p2p" (fused peer-memory exchange, default when the group runs on NCCL) or "nccl"; env VAP_ULYSSES overrides the default."""
    global _CURRENT
    import os
    if mode is None:
        mode = os.environ.get("VAP_ULYSSES") or ("p2p" if dist.get_backend(group) == "nccl" else "nccl")
    if mode not in ("p2p", "nccl"):
        raise ValueError(f"Ulysses mode must be 'p2p' or 'nccl', got {mode!r}")
    _CURRENT = SequenceParallel(group, dist.get_rank(group), dist.get_world_size(group), mode, {})
    return _CURRENT


def disable() -> None:
    global _CURRENT
    _CURRENT = None


def current() -> Optional[SequenceParallel]:
    return _CURRENT if (_CURRENT is not None and _CURRENT.world > 1) else None


def check_divisible(tokens: int, heads: int, world: int) -> None:
    if tokens % world:
        raise ValueError(f"Ulysses: {tokens} tokens per stream are not divisible by {world} ranks")
    if heads % world:
        raise ValueError(f"Ulysses: {heads} heads are not divisible by {world} ranks")


def shard_rows(x: torch.Tensor, sp: SequenceParallel, dim: int = 1) -> torch.Tensor:
    """Rank-local slice of a token-major tensor (what the reference's ContextParallelSplitHook does, ptd.py:545-628)."""
    n = x.shape[dim]
    if n % sp.world:
        raise ValueError(f"Ulysses: dim {dim} of size {n} is not divisible by {sp.world} ranks")
    return x.narrow(dim, sp.rank * (n // sp.world), n // sp.world)


def gather_rows(x: torch.Tensor, sp: SequenceParallel, dim: int = 1) -> torch.Tensor:
    """All-gather of rank-local rows back to the full sequence (the reference gathers at proj_out, ptd.py:675-679)."""
    x = x.movedim(dim, 0).contiguous()
    out = torch.empty((sp.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=sp.group)
    return out.movedim(0, dim)


def _pack_cuda(src: torch.Tensor, nsplit: int, out: torch.Tensor) -> None:
    ops.ulysses_pack(src, nsplit, out)


def _unpack_cuda(src: torch.Tensor, out: torch.Tensor) -> None:
    ops.ulysses_unpack(src, out)


def exchange_qkv(qkv: torch.Tensor, heads: int, sp: SequenceParallel, pack: Callable = _pack_cuda) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """All-to-all #1.  qkv [L_loc, 3*H*D] (local rows of both streams, columns q|k|v, heads in natural order) ->
    q, k, v views [1, H/P, P*L_loc, D] over the full (rank-major) joint sequence for this rank's heads."""
    P = sp.world
    L, width = qkv.shape
    inner = width // 3
    check_divisible(L * P, heads, P)
    hp = inner // P  # (H/P)*D columns per rank
    send = torch.empty((P, L, 3, hp), dtype=qkv.dtype, device=qkv.device)
    for w in range(3):
        pack(qkv[:, w * inner:(w + 1) * inner], P, send[:, :, w, :])
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=sp.group)
    joint = recv.view(1, P * L, 3, heads // P, hp // (heads // P))  # [1, J, 3, H/P, D]
    q, k, v = (joint[:, :, w].transpose(1, 2) for w in range(3))
    return q, k, v


def exchange_out(o: torch.Tensor, sp: SequenceParallel, unpack: Callable = _unpack_cuda) -> torch.Tensor:
    """All-to-all #2.  o [P*L_loc, (H/P)*D] (all joint rows, this rank's heads) -> [L_loc, H*D] (local rows, all heads)."""
    P = sp.world
    J, hp = o.shape
    L = J // P
    o = o.contiguous()
    recv = torch.empty((P, L, hp), dtype=o.dtype, device=o.device)
    dist.all_to_all_single(recv, o.view(P, L, hp), group=sp.group)
    out = torch.empty((L, P * hp), dtype=o.dtype, device=o.device)
    unpack(recv, out)
    return out


# ----------------------------------------------------------------------------------------------
# fused exchange over NVLink peer memory
# ----------------------------------------------------------------------------------------------
class PeerExchange:
    """Symmetric (peer-mapped) buffers of one joint-attention shape, shared by all MoT blocks.

      recv[f] [B, P, rows, 3, (H/P)*D]  slot s = what rank s dispatched to this rank (its local rows of both streams, q|k|v of this
                                        rank's heads); rows = local joint rows.  Viewed as [B, P*rows, 3, H/P, D] it is the
                                        rank-major joint sequence the attention kernel reads through strided TMA descriptors.
      out[f]  [B, rows, H*D]            this rank's rows of O for ALL heads; rank r's attention kernel fills columns
                                        [r*(H/P)*D, (r+1)*(H/P)*D).
    B = the batch of the forward (1 per guidance pass in the Wan pipeline, 2 = [conditional | unconditional] in the CogVideoX pipeline,
    pipeline_cogvideox_image2video_mot.py:972-1001): one exchange and ONE attention launch serve the whole batch.

    f alternates per block.  Why two sets suffice: block b writes set b%2 on the peers; a peer's last read of that set (block b-2's
    attention for recv, block b-2's output projections for out) is stream-ordered before the barrier that peer entered in
    block b-1, which this rank has passed before it launches block b's kernels."""

    def __init__(self, sp: SequenceParallel, rows: int, heads: int, head_dim: int, device: torch.device, batch: int = 1):
        import torch.distributed._symmetric_memory as symm
        P = sp.world
        self.sp, self.rows, self.heads, self.head_dim, self.batch = sp, rows, heads, head_dim, batch
        self.hp = heads // P * head_dim
        self.recv_batch_bytes = P * rows * 3 * self.hp * 2  # one batch element of a receive buffer
        group = sp.group if sp.group is not None else dist.group.WORLD
        self.recv, self.out, self.recv_ptrs, self.out_ptrs = [], [], [], []
        self._handles = []
        for _ in range(2):
            r = symm.empty((batch, P, rows, 3, self.hp), dtype=torch.bfloat16, device=device)
            h = symm.rendezvous(r, group)
            self.recv.append(r), self.recv_ptrs.append([int(p) for p in h.buffer_ptrs]), self._handles.append(h)
            o = symm.empty((batch, rows, heads * head_dim), dtype=torch.bfloat16, device=device)
            h = symm.rendezvous(o, group)
            # every rank writes its own head columns of the owner's rows
            self.out.append(o), self.out_ptrs.append([int(p) + sp.rank * self.hp * 2 for p in h.buffer_ptrs]), self._handles.append(h)
        self.flip = 1

    def next_block(self) -> None:
        self.flip ^= 1

    def barrier(self) -> None:
        self._handles[0].barrier(channel=0)  # stream-ordered: signals every peer and waits for every peer

    def dispatch(self, qkv: torch.Tensor, row0: int, batch_index: int = 0, **norm_rope) -> None:
        """Exchange #1 for one stream of one batch element: qkv [L, 3*H*D] = this rank's rows of the stream's QKV projection (pre-norm)."""
        inner = qkv.shape[-1] // 3
        if not 0 <= batch_index < self.batch:
            raise IndexError(f"batch element {batch_index} of a PeerExchange made for batch {self.batch}")
        ptrs = self.recv_ptrs[self.flip] if batch_index == 0 else [p + batch_index * self.recv_batch_bytes for p in self.recv_ptrs[self.flip]]
        ops.qkv_scatter(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], heads=self.heads, head_dim=self.head_dim,
                        dst_ptrs=ptrs, dst_slot=self.sp.rank, slot_rows=self.rows, dst_row0=row0, **norm_rope)

    def attention(self) -> torch.Tensor:
        """barrier -> joint attention of this rank's heads over the full sequence, O rows stored to their owners -> barrier.
        Returns this rank's rows of O for all heads, [B, rows, H*D]."""
        P, hp, D = self.sp.world, self.hp, self.head_dim
        self.barrier()
        joint = self.recv[self.flip].view(self.batch, P * self.rows, 3, self.heads // P, D)
        q, k, v = (joint[:, :, w].transpose(1, 2) for w in range(3))
        ops.attention_scatter(q, k, v, o_ptrs=self.out_ptrs[self.flip], rows_per_peer=self.rows,
                              o_strides=(self.rows * self.heads * D, D, self.heads * D))
        self.barrier()
        return self.out[self.flip]


def peer_exchange(sp: SequenceParallel, rows: int, heads: int, head_dim: int, device: torch.device, batch: int = 1) -> PeerExchange:
    """The (cached) PeerExchange of this shape.  Creating one is collective: every rank must get here in the same order."""
    key = (rows, heads, head_dim, batch)
    if key not in sp._exchanges:
        check_divisible(rows * sp.world, heads, sp.world)
        sp._exchanges[key] = PeerExchange(sp, rows, heads, head_dim, device, batch)
    return sp._exchanges[key]
