"""CPU, world_size 2 and 4, gloo: the Ulysses partitioning / all-to-all plumbing of video-as-prompt_b200/ulysses.py.
The pack / unpack re-layouts are injected as torch permutes (the product path uses the CUDA kernels, covered by the
`ulysses_relayout` GPU check); attention itself is the oracle's SDPA.  Asserts the N-rank result equals the 1-rank one."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pack_ref(src, nsplit, out):
    L, width = src.shape
    out.copy_(src.view(L, nsplit, width // nsplit).permute(1, 0, 2))


def _unpack_ref(src, out):
    n, L, c = src.shape
    out.copy_(src.permute(1, 0, 2).reshape(L, n * c))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vap = importlib.import_module("video-as-prompt_b200")
        from oracle.common import sdpa
        uly = vap.ulysses
        sp = uly.enable()
        assert uly.current() is sp and sp.world == world
        H, D, S = 4, 8, 8 * world  # tokens per stream
        g = torch.Generator().manual_seed(0)
        qkv_t = torch.randn((S, 3 * H * D), generator=g).to(torch.bfloat16)  # target stream (already q/k-normed + RoPE'd)
        qkv_r = torch.randn((S, 3 * H * D), generator=g).to(torch.bfloat16)  # reference stream
        # single-process answer: joint attention over [target | ref]
        joint = torch.cat([qkv_t, qkv_r], 0)
        q, k, v = (joint[:, i * H * D:(i + 1) * H * D].view(1, 2 * S, H, D).transpose(1, 2) for i in range(3))
        full = sdpa(q, k, v).transpose(1, 2).reshape(2 * S, H * D)
        # N ranks: each owns S/P target rows and S/P ref rows
        loc = torch.cat([uly.shard_rows(qkv_t, sp, 0), uly.shard_rows(qkv_r, sp, 0)], 0)
        ql, kl, vl = uly.exchange_qkv(loc, H, sp, pack=_pack_ref)
        assert ql.shape == (1, H // world, 2 * S, D)
        o = sdpa(ql, kl, vl).transpose(1, 2).reshape(2 * S, (H // world) * D)
        o_loc = uly.exchange_out(o, sp, unpack=_unpack_ref)  # [2*S/P, H*D]: local target rows then local ref rows
        n = S // world
        want = torch.cat([full[rank * n:(rank + 1) * n], full[S + rank * n:S + (rank + 1) * n]], 0)
        err = (o_loc.float() - want.float()).abs().max().item()
        gathered = uly.gather_rows(o_loc[:n].unsqueeze(0), sp, dim=1)[0]
        gerr = (gathered.float() - full[:S].float()).abs().max().item()
        with pytest.raises(ValueError):
            uly.check_divisible(S + 1, H, world)
        ret[rank] = (err, gerr)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_ulysses_equals_single_rank(world):
    import random
    port = 29500 + random.randint(0, 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        err, gerr = ret[r]
        assert err < 2e-2 and gerr < 2e-2, (r, err, gerr)
