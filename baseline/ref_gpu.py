"""The reference's OWN transformer classes on the GPU (stock PyTorch path: F.scaled_dot_product_attention, nn.Linear, eager glue) —
the anchor SURVEY §8(d)(i) asks for ("reference modules on the same B200 with stock SDPA").  Test / bench infrastructure only.

  build_reference(family, cfg, seed | share_with=our_model)   reference model in bf16 on the device, synthetic weights (or OUR model's
                                                              very tensors, assigned — no second copy of 65 GB at 14B)
  record_blocks(model, inputs)                                 stock forward + every block's (kwargs, outputs) via forward hooks
  time_forward(model, inputs, steps, warmup)                   CUDA-event timed forwards
  sdpa_kernels(model, inputs)                                  names of the attention kernels the stock SDPA dispatched to (torch.profiler)
"""
from __future__ import annotations

import importlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_loader  # noqa: E402


def _vap():
    return importlib.import_module("video-as-prompt_b200")


def reference_class(family: str):
    d = ref_loader.load()
    return d.WanTransformer3DMOTModel if family == "wan" else d.CogVideoXTransformer3DMOTModel


def build_reference(family: str, cfg: dict, seed: int = 1234, device="cuda", share_with=None):
    """Constructed on the meta device (no fp32 host copy of a 14B model), materialised in bf16 on `device`.
    transformer_wan_mot.py:368-388: WanRotaryPosEmbed keeps its frequency table as a plain attribute built in __init__, which the
    meta construction leaves without storage — that one module is rebuilt normally."""
    cls = reference_class(family)
    with torch.device("meta"):
        model = cls(**cfg)
    model = model.to(torch.bfloat16)
    if share_with is not None:
        sd = share_with.state_dict()
        missing = set(model.state_dict()) ^ set(sd)
        if missing:
            raise KeyError(f"state_dict keys differ between the reference class and the mirror shell: {sorted(missing)[:6]}")
        model.load_state_dict(sd, assign=True)
    else:
        model = model.to_empty(device=device)
        _vap().synth.fill_module_(model, seed=seed, num_layers=cfg["num_layers"])
    if family == "wan":
        rp = model.rope
        model.rope = type(rp)(rp.attention_head_dim, rp.patch_size, rp.max_seq_len)
    for p in model.parameters():
        p.requires_grad_(False)
    return model.eval()


def blocks_of(model):
    return list(model.blocks) if hasattr(model, "blocks") else list(model.transformer_blocks)


@torch.no_grad()
def record_blocks(model, inputs):
    """-> (final output, [(kwargs_i, outputs_i)]) of one stock forward; tensors are kept by reference (the reference blocks do not
    write into their inputs)."""
    rec = []
    hooks = [blk.register_forward_hook(lambda m, a, kw, out: rec.append((dict(kw), out)), with_kwargs=True) for blk in blocks_of(model)]
    try:
        final = model(**inputs, return_dict=False)[0]
    finally:
        for h in hooks:
            h.remove()
    return final, rec


@torch.no_grad()
def time_forward(model, inputs, steps: int, warmup: int, after=None):
    """ms per forward (CUDA events on the current stream, after `warmup` untimed forwards); after(out, i) runs inside the timed
    region after each forward (e.g. the scheduler update)."""
    out = None
    for i in range(warmup):
        out = model(**inputs, return_dict=False)[0]
        if after is not None:
            after(out, i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = model(**inputs, return_dict=False)[0]
        if after is not None:
            after(out, warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / max(steps, 1), out


@torch.no_grad()
def sdpa_kernels(model, inputs, top: int = 4):
    """Device kernels of one stock forward, by total time: the top ones and every attention-looking name (which SDPA backend fired)."""
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model(**inputs, return_dict=False)
        torch.cuda.synchronize()
    rows = sorted(((e.key, getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0.0), e.count) for e in prof.key_averages()),
                  key=lambda r: -r[1])
    total = sum(r[1] for r in rows) or 1.0
    pick = rows[:top] + [r for r in rows[top:] if any(s in r[0].lower() for s in ("attn", "attention", "fmha", "flash", "sdpa"))]
    return [dict(kernel=k[:120], share=round(t / total, 4), launches=c) for k, t, c in pick]
