"""A/B of the two-CUDA-stream block schedule (video-as-prompt_b200/streams.py) against the one-stream schedule, same process, alternating.

    python tools/dual_stream_ab.py [--iters 8] > gpurun_out/dual_stream_ab.json

Models: 2 MoT blocks at the Wan-14B widths (a) at the full 480p token count (20 280 per stream: what one GPU runs) and (b) at 2 535 tokens per
stream — the rows ONE RANK owns under 8-way Ulysses, i.e. the GEMM / norm shapes whose wave quantisation and launch floors limit the 8-GPU step
(the attention of that proxy is small; the exchange is not part of it) — and 2 MoT blocks at the CogVideoX-5B widths.  Outputs of the two
schedules must be bit-identical (same kernels, same per-stream order).
"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vap = importlib.import_module("video-as-prompt_b200")
synth = vap.synth


def build(cls, cfg):
    with torch.device("meta"):
        m = cls(**cfg)
    m = m.to(torch.bfloat16).to_empty(device="cuda")
    synth.fill_module_(m, seed=1234, num_layers=cfg["num_layers"])
    return m.eval()


def time_forward(model, inp, iters):
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = model(**inp, return_dict=False)[0]
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0], out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--case", default="", help="substring selecting the cases to run")
    ap.add_argument("--profile", action="store_true", help="one forward per case (one-stream schedule) between cudaProfilerStart/Stop, for an ncu launch list")
    a = ap.parse_args()
    cases = [
        ("wan14b_2blocks_480p_full", "wan", dict(synth.WAN_14B, num_layers=2, block_idx_with_mot_ref=[0, 1]), (13, 60, 104)),
        ("wan14b_2blocks_rank_of_8_rows", "wan", dict(synth.WAN_14B, num_layers=2, block_idx_with_mot_ref=[0, 1]), (13, 30, 26)),
        ("cog5b_2blocks_480p_full", "cog", dict(synth.COG_5B, num_layers=2, block_idx_with_mot_ref=[0, 1]), (13, 60, 90)),
    ]
    with torch.no_grad():
        for name, fam, cfg, (f, h, w) in cases:
            if a.case and a.case not in name:
                continue
            model = build(vap.WanTransformer3DMOTModel if fam == "wan" else vap.CogVideoXTransformer3DMOTModel, cfg)
            inp = synth.wan_inputs(cfg, f, h, w, device="cuda") if fam == "wan" else synth.cog_inputs(cfg, f, h, w, device="cuda")
            res = {"case": name}
            outs = {}
            if a.profile:
                with vap.dual_streams(False):
                    time_forward(model, inp, 2)
                    torch.cuda.cudart().cudaProfilerStart()
                    model(**inp, return_dict=False)
                    torch.cuda.synchronize()
                    torch.cuda.cudart().cudaProfilerStop()
                print(json.dumps(dict(case=name, profiled=True)), flush=True)
                continue
            for rnd in range(2):
                for mode in (False, True):
                    with vap.dual_streams(mode):
                        if rnd == 0:
                            time_forward(model, inp, 2)  # warm-up (packs weights, builds tables, fills both allocator pools)
                        med, best, out = time_forward(model, inp, a.iters)
                    key = "dual" if mode else "single"
                    outs[key] = out
                    res.setdefault(key + "_ms", []).append(round(med, 3))
                    res.setdefault(key + "_best_ms", []).append(round(best, 3))
            res["bit_identical"] = bool(torch.equal(outs["dual"], outs["single"]))
            res["speedup_median"] = round(min(res["single_ms"]) / min(res["dual_ms"]), 4)
            print(json.dumps(res), flush=True)
            del model, inp, outs
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
