"""Host-side launch-path profiler (runs WITHOUT a GPU): how long does Python take to issue one forward's launches?

At 8 GPUs the 480p step did not get faster with faster kernels: how much of it is the Python / ctypes launch path?  This tool swaps libvap_b200.so for a stub whose
entry points have the same names and return 0 at once (built here with gcc from the SIGNATURES table), lets CPU bf16 tensors through
the ops' device checks, and runs the stand-alone Wan / CogVideoX shell at small shapes — so that what remains is exactly the host
work per launch: argument validation, torch.empty, view arithmetic, ctypes marshalling, and the torch glue ops.

    python tools/host_path_profile.py [--family wan|cog] [--blocks 8] [--iters 20] [--record calls.json] [--profile]

--record dumps the sequence of C-ABI calls (name + every non-pointer argument + the aliasing pattern of the pointer arguments), which
tests/test_host_logic.py uses to pin the launch sequence: host-path optimisations must not change what is launched.
Developer tool only: nothing here computes anything, no number it prints is a benchmark value.
"""
from __future__ import annotations

import argparse
import cProfile
import ctypes
import importlib
import json
import os
import pstats
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def build_stub(sigs) -> ctypes.CDLL:
    """A shared library exporting every SIGNATURES name as `int f() { return 0; }` (cdecl ignores the extra arguments)."""
    src = "\n".join(f"int {name}() {{ return {300 if name == 'vap_version' else (148 if name == 'vap_sm_count' else 0)}; }}"
                    for name in sigs if name != "vap_last_error")
    src += '\nconst char* vap_last_error() { return "stub"; }\n'
    d = os.path.join(ROOT, "video-as-prompt_b200", "csrc", "build", "stub")  # git-ignored build directory
    os.makedirs(d, exist_ok=True)
    c, so = os.path.join(d, "stub.c"), os.path.join(d, f"libvap_stub_{os.getpid()}.so")
    open(c, "w").write(src)
    subprocess.run(["gcc", "-shared", "-fPIC", "-O1", "-w", "-o", so, c], check=True)
    lib = ctypes.CDLL(so)
    os.unlink(so)  # the mapping stays valid; nothing is left behind
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


class Recorder:
    """Wraps the stub: logs (name, scalar arguments, pointer-alias pattern) of every call."""

    def __init__(self, lib, sigs):
        self.calls = []
        self._ptr_ids = {}
        for name, (_, argtypes) in sigs.items():
            setattr(self, name, self._wrap(name, getattr(lib, name), argtypes))

    def _wrap(self, name, fn, argtypes):
        is_ptr = [t is ctypes.c_void_p for t in argtypes]

        def call(*args):
            rec = [name]
            for a, p in zip(args, is_ptr):
                if p:
                    if isinstance(a, ctypes.Array):  # peer pointer table
                        rec.append(["table", len(a)])
                    elif not a:
                        rec.append(None)
                    else:
                        rec.append("p")
                else:
                    rec.append(round(a, 9) if isinstance(a, float) else a)
            self.calls.append(rec)
            return fn(*args)
        return call


def patch(vap, lib):
    ops = vap.ops

    def need_bf16(t, name):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if t.dtype != torch.bfloat16:
            raise TypeError(f"{name} must be torch.bfloat16, got {t.dtype}")

    def need_f32(t, name):
        if t is None:
            return 0
        if t.dtype != torch.float32:
            raise TypeError(f"{name} must be float32")
        if t.stride(-1) != 1:
            raise ValueError(f"{name} must be contiguous in its last dimension")
        return t.data_ptr()

    ops._need_cuda_bf16 = need_bf16
    ops._need_cuda_f32 = need_f32
    ops._need_cuda_float = lambda t, name: None
    ops._stream = lambda: 0
    if hasattr(ops, "_dev_check"):
        ops._dev_check = lambda *a, **k: None
    vap._lib._lib = lib
    vap._lib.load = lambda: lib


def make_model(vap, family: str, blocks: int):
    torch.manual_seed(0)
    if family == "wan":
        cfg = dict(vap.synth.WAN_TINY, num_layers=blocks, block_idx_with_mot_ref=list(range(blocks)))
        model = vap.WanTransformer3DMOTModel(**cfg).to(torch.bfloat16).eval()
        inputs = vap.synth.wan_inputs(cfg, 2, 8, 8)
    else:
        cfg = dict(vap.synth.COG_TINY, num_layers=blocks, block_idx_with_mot_ref=list(range(max(blocks - 1, 1))))
        model = vap.CogVideoXTransformer3DMOTModel(**cfg).to(torch.bfloat16).eval()
        inputs = vap.synth.cog_inputs(cfg, 2, 8, 8)
    return model, inputs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--family", default="wan", choices=["wan", "cog"])
    ap.add_argument("--blocks", type=int, default=8)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--record", default=None)
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    torch.set_num_threads(1)
    vap = importlib.import_module("video-as-prompt_b200")
    stub = build_stub(vap._lib.SIGNATURES)
    rec = Recorder(stub, vap._lib.SIGNATURES)
    patch(vap, rec)
    model, inputs = make_model(vap, a.family, a.blocks)
    with torch.no_grad():
        model(**inputs)  # packs weights, fills the caches
        rec.calls.clear()
        model(**inputs)
        ncalls = len(rec.calls)
        if a.record:
            json.dump(rec.calls, open(a.record, "w"))
        patch(vap, stub)  # time without the recorder
        for _ in range(3):
            model(**inputs)
        t0 = time.perf_counter()
        for _ in range(a.iters):
            model(**inputs)
        dt = (time.perf_counter() - t0) / a.iters
        print(json.dumps({"family": a.family, "blocks": a.blocks, "c_abi_calls_per_forward": ncalls, "host_ms_per_forward": round(dt * 1e3, 3),
                          "host_us_per_block": round(dt * 1e6 / a.blocks, 1), "host_us_per_call": round(dt * 1e6 / max(ncalls, 1), 2)}))
        if a.profile:
            pr = cProfile.Profile()
            pr.enable()
            for _ in range(a.iters):
                model(**inputs)
            pr.disable()
            pstats.Stats(pr).sort_stats("tottime").print_stats(30)


if __name__ == "__main__":
    main()
