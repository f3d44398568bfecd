"""CUDA-graph replay of a transformer forward (fixed shapes): the whole step — every kernel of the 40 blocks, the Ulysses peer
barriers, the final all-gather — is captured once and replayed with one launch call.

Why: at 8 GPUs a Wan-14B 480p step is ~5 400 kernel launches of ~50 us each; round 1 suspected the Python / ctypes launch path of being the
bound there.  The reference has no counterpart (eager PyTorch); SURVEY.md §8(f) rank 2 lists graph capture of the step as a shell-level item.

    graphed = GraphedForward(model, example_kwargs)      # 2 eager warm-up forwards on a side stream, then capture
    out = graphed(**kwargs)[0]                           # copies tensor inputs into the captured buffers, replays

Status (round 2, B200s): capture (including the fork / join of the blocks' side stream, streams.py, the peer-memory exchange, the all-gather of the
sharded context projections and the final all-gather), replay and process exit are clean at 1, 2 and 8 GPUs.  The gain is what the launch path
costs: none on one GPU (GPU-bound), 1014.2 -> 1012.9 ms at 2 GPUs, 268.5 -> 266.0 ms at 8 (profiles/r02_bench_n8_v2.json) — the 8-GPU step is
GPU-bound too, so replay stays opt-in (`bench.py --graph on`).
"""
from __future__ import annotations

from typing import Any, Dict, Tuple

import torch


class GraphedForward:
    def __init__(self, model, example_kwargs: Dict[str, Any], warmup: int = 2):
        self.model = model
        self.static_in = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example_kwargs.items()}
        self._shapes = {k: (tuple(v.shape), v.dtype) for k, v in self.static_in.items() if torch.is_tensor(v)}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():  # warm-up off the default stream: lazy inits (weight packing, fp32 parameter
            for _ in range(warmup):                     # copies, cudaFuncSetAttribute, symmetric-memory rendezvous) happen here
                model(**self.static_in, return_dict=False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = model(**self.static_in, return_dict=False)[0]

    def __call__(self, return_dict: bool = False, **kwargs) -> Tuple[torch.Tensor]:
        for k, v in kwargs.items():
            if torch.is_tensor(v):
                if (tuple(v.shape), v.dtype) != self._shapes.get(k):
                    raise ValueError(f"GraphedForward was captured for {k} of shape/dtype {self._shapes.get(k)}, got {(tuple(v.shape), v.dtype)}")
                if v.data_ptr() != self.static_in[k].data_ptr():
                    self.static_in[k].copy_(v, non_blocking=True)
            elif v != self.static_in.get(k):
                raise ValueError(f"GraphedForward was captured with {k}={self.static_in.get(k)!r}, got {v!r}")
        self.graph.replay()
        return (self.static_out,)
