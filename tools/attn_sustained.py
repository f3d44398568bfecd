"""Joint attention under SUSTAINED load: our kernel and cuDNN's SDPA each run back to back for `--seconds` (the regime a denoise step is in:
the chip sits at its power cap), TFLOP/s of the last half of the window with the SM clock and board power sampled beside it.
    python tools/attn_sustained.py [--seconds 8] > gpurun_out/attn_sustained.json
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vap = importlib.import_module("video-as-prompt_b200")


class Sampler:
    def __init__(self):
        self.rows, self.stop = [], False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True, timeout=5).stdout
                mhz, watt = (float(x) for x in out.strip().split(","))
                self.rows.append((time.time(), mhz, watt))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.2)


def sustained(fn, seconds, flop):
    fn()
    torch.cuda.synchronize()
    s = Sampler()
    s.t.start()
    t0 = time.time()
    marks = []
    while time.time() - t0 < seconds:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            fn()
        e1.record()
        torch.cuda.synchronize()
        marks.append((time.time() - t0, e0.elapsed_time(e1) / 4))
    s.stop = True
    s.t.join(2)
    late = [ms for t, ms in marks if t > seconds / 2]
    first = [ms for t, ms in marks if t < 1.0]
    rows = [r for r in s.rows if r[0] - t0 > seconds / 2]
    med = lambda v: sorted(v)[len(v) // 2] if v else None  # noqa: E731
    return dict(first_second_tflops=round(flop / med(first) / 1e9, 1), sustained_tflops=round(flop / med(late) / 1e9, 1), sustained_ms=round(med(late), 3),
                sm_mhz=med([r[1] for r in rows]), power_w=med([r[2] for r in rows]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=8.0)
    a = ap.parse_args()
    H, J, D = 40, 40560, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn((1, J, 3 * H * D), generator=g, device="cuda").to(torch.bfloat16)
    q, k, v = (qkv[..., i * H * D:(i + 1) * H * D].unflatten(2, (H, D)).transpose(1, 2) for i in range(3))
    qc, kc, vc = q.contiguous(), k.contiguous(), v.contiguous()
    flop = 4.0 * H * J * J * D
    from torch.nn.attention import SDPBackend, sdpa_kernel
    res = {"shape": dict(H=H, J=J, D=D)}
    for rnd in range(2):
        res[f"vap_round{rnd}"] = sustained(lambda: vap.ops.attention(q, k, v), a.seconds, flop)
        with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
            res[f"cudnn_round{rnd}"] = sustained(lambda: F.scaled_dot_product_attention(qc, kc, vc), a.seconds, flop)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
