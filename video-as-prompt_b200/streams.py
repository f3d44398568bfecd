"""Two CUDA streams per MoT block: the frozen DiT's target stream and the VAP expert's reference stream are independent between two joint
attentions (adaLN-LayerNorm -> QKV -> q/k norm + RoPE before it; O-projection -> cross-attention -> FFN after it: transformer_wan_mot.py:620-635
and :649-697, cogvideox_transformer_3d_mot.py:407-422 and :445-511), so the expert's kernels are launched on a side stream and the hardware's
block scheduler fills the tail wave of one stream's persistent GEMM (and the launch gap behind its short kernels) with the other stream's CTAs.
Same kernels, same per-stream order: results are bit-identical to the one-stream schedule.  Matters most under Ulysses, where a rank's GEMMs
have M = 2 535 rows (8.1 / 2.7 / 7.3 waves of CTA pairs) and ~25 short kernels per block sit at their launch floors.

Ordering / memory rules the block forwards follow (so that no caching-allocator block is reused while the other stream still needs it):
  * fork(): the side stream waits for everything issued on the main stream so far;  join(): the main stream waits for the side stream.
  * every block that forks joins before it returns — callers (the reference's shell, forward hooks, tests) may read both outputs on the main stream;
  * a tensor allocated while the side stream is current is only ever freed after a later join; a main-stream tensor the side stream reads
    (the attention output, the contexts) stays referenced until the join that follows the read.
Capturable: inside a CUDA-graph capture the fork makes the side stream part of the capture and the join brings it back (graphs.GraphedForward).

Measured on a B200 (tools/dual_stream_ab.py, profiles/r02_dual_stream_ab.json; outputs bit-identical): 2 MoT blocks at the Wan-14B widths with
the 2 535 rows per stream ONE RANK owns under 8-way Ulysses 8.55 -> 8.02 ms (1.066x); at the full 20 280 rows 1.004x, CogVideoX-5B full size
0.993x (noise).  Default ("auto") therefore: on when a stream has at most AUTO_MAX_ROWS rows, i.e. whenever the sequence is sharded.
VAP_DUAL_STREAM=0 / 1 (or `dual_streams(False / True)`) forces it off / on.
"""
from __future__ import annotations

import contextlib
import os
from typing import Dict, Optional

import torch

AUTO_MAX_ROWS = 12288
_env = os.environ.get("VAP_DUAL_STREAM", "auto")
_ENABLED = [None if _env == "auto" else _env != "0"]  # None = auto (by rows), True / False = forced


class DualStream:
    def __init__(self, device: torch.device):
        self.side = torch.cuda.Stream(device=device)
        self._fork = torch.cuda.Event()
        self._join = torch.cuda.Event()

    def fork(self) -> None:
        self._fork.record(torch.cuda.current_stream())
        self.side.wait_event(self._fork)

    def join(self) -> None:
        self._join.record(self.side)
        torch.cuda.current_stream().wait_event(self._join)

    def on_side(self):
        return torch.cuda.stream(self.side)


_PER_DEVICE: Dict[int, DualStream] = {}


def dual(device: torch.device, rows: int = 0) -> Optional[DualStream]:
    """The DualStream of a CUDA device for a block whose streams have `rows` rows each, or None (CPU tensors, switched off, or — in the default
    "auto" mode — streams long enough to fill the GPU on their own)."""
    on = _ENABLED[0] if _ENABLED[0] is not None else rows <= AUTO_MAX_ROWS
    if not on or device.type != "cuda":
        return None
    idx = device.index if device.index is not None else torch.cuda.current_device()
    ds = _PER_DEVICE.get(idx)
    if ds is None:
        ds = _PER_DEVICE[idx] = DualStream(torch.device("cuda", idx))
    return ds


def side(ds: Optional[DualStream]):
    """Context manager: the side stream of `ds`, or nothing when dual-stream execution is off."""
    return ds.on_side() if ds is not None else contextlib.nullcontext()


class dual_streams:
    """Context manager / switch: `with dual_streams(False): ...` runs both token streams on the current stream, `dual_streams(True)` forces the
    two-stream schedule, `dual_streams(None)` restores the automatic choice."""

    def __init__(self, enabled: Optional[bool] = True):
        self.enabled = None if enabled is None else bool(enabled)

    def __enter__(self):
        self._prev = _ENABLED[0]
        _ENABLED[0] = self.enabled
        return self

    def __exit__(self, *exc):
        _ENABLED[0] = self._prev
        return False
