"""Ulysses sequence parallelism for the MoT joint attention (one process per GPU, NCCL all-to-all over NVLink/NVSwitch).

The reference scales long sequences with ring attention (finetrainers/models/attention_dispatch.py:686-773, hooks in
finetrainers/parallel/ptd.py:515-679); everything in a MoT block except the joint attention is token-local, so here both
streams are sharded along tokens (rank r owns rows [r*L/P, (r+1)*L/P) of each stream) and each block does exactly two
exchanges: all-to-all #1 turns the local rows x all heads of q|k|v into all rows x H/P heads, the attention kernel runs on
H/P heads over the full joint sequence, all-to-all #2 brings O back to local rows x all heads.

After exchange #1 the joint sequence is ordered rank-major ([tgt_0|ref_0|tgt_1|ref_1|...]) — a permutation of the
reference's [target|ref] order, which is irrelevant for unmasked attention as long as O rows map back to their tokens,
which exchange #2 does by construction.

The re-layout (pack / unpack) functions are injectable so the partitioning logic is testable on CPU with gloo; the
product path uses the CUDA kernels vap_ulysses_pack / vap_ulysses_unpack.
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


_CURRENT: Optional[SequenceParallel] = None


def enable(group: Optional[dist.ProcessGroup] = None) -> SequenceParallel:
    """Activate Ulysses over `group` (default: the world group).  world == 1 is a no-op context."""
    global _CURRENT
    _CURRENT = SequenceParallel(group, dist.get_rank(group), dist.get_world_size(group))
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
