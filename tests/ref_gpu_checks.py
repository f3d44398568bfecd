"""GPU parity against the REFERENCE'S OWN CLASSES running on the same device (stock PyTorch path), at the real model widths.

The reference (bytedance/Video-As-Prompt's vendored diffusers) is staged, unmodified, under the git-ignored baseline/_ref
(baseline/ref_loader.py; it travels to the GPU box with the snapshot).  Each check builds the reference's
WanTransformer3DMOTModel / CogVideoXTransformer3DMOTModel with synthetic weights at Wan-14B / CogVideoX-5B widths, runs it stock,
then `vap_b200.install(model, level=...)` ON THAT SAME INSTANCE and runs it again:

  * per block, teacher-forced: block i of the installed path gets the stock run's recorded inputs of block i and
    must reproduce its recorded outputs within max-abs <= 2e-2 relative (north_star's bf16 gate); alongside, both paths' error
    against an fp32 evaluation of the same (bf16-rounded) weights on the same inputs is reported;
  * whole forward and 4-step denoise with classifier-free guidance: cosine >= 0.999;
  * `uninstall` restores the stock output bit for bit.
"""
from __future__ import annotations

import copy
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "baseline")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_gpu  # noqa: E402
import ref_loader  # noqa: E402

vap = importlib.import_module("video-as-prompt_b200")
synth = vap.synth
DEV = "cuda"


def available() -> bool:
    return ref_loader.available()


def rel_err(a, b):
    a, b = a.detach().float(), b.detach().float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def cosine(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm())).item()


def _to(x, fn):
    if torch.is_tensor(x):
        return fn(x)
    if isinstance(x, (list, tuple)):
        return type(x)(_to(t, fn) for t in x)
    if isinstance(x, dict):
        return {k: _to(v, fn) for k, v in x.items()}
    return x


def _f32(x):
    return _to(x, lambda t: t.float() if t.dtype == torch.bfloat16 else t)


def _inputs(family, cfg, latent, batch=1):
    f, h, w = latent
    make = synth.wan_inputs if family == "wan" else synth.cog_inputs
    return make(cfg, f, h, w, seed=0, device=DEV, batch=batch)


@torch.no_grad()
def check_blocks(family: str, cfg: dict, latent, tol: float = 2e-2, with_fp32: bool = True, levels=("block", "processor")):
    """Teacher-forced per-block parity + whole forward, install() on the reference instance itself."""
    model = ref_gpu.build_reference(family, cfg, seed=7, device=DEV)
    inp = _inputs(family, cfg, latent)
    keys = list(model.state_dict())
    final, rec = ref_gpu.record_blocks(model, inp)
    res = {"tokens_per_stream": int(rec[0][0]["hidden_states"].shape[1]), "blocks": len(rec)}

    truth = None
    if with_fp32:  # the same bf16-rounded weights evaluated in fp32 by the reference's own block code, on each block's recorded inputs
        truth = []
        try:
            for blk, (kw, _) in zip(ref_gpu.blocks_of(model), rec):
                b32 = copy.deepcopy(blk).float()
                truth.append(b32(**_f32(kw)))
                del b32
        except Exception as exc:  # noqa: BLE001 — the fp32 evaluation is a report, not a gate
            res["fp32_truth_unavailable"] = f"{type(exc).__name__}: {str(exc)[:160]}"
            truth = None
        torch.cuda.empty_cache()

    outs = lambda o: [t for t in (o if isinstance(o, (tuple, list)) else (o,)) if torch.is_tensor(t)]  # noqa: E731
    vap.install(model, level="block")
    try:
        per_block, vs32 = {}, {}
        for i, (blk, (kw, out)) in enumerate(zip(ref_gpu.blocks_of(model), rec)):
            got = blk(**kw)
            for j, (g, r) in enumerate(zip(outs(got), outs(out))):
                if r is kw.get("hidden_states_mot_ref"):
                    continue  # plain block: the reference stream is passed through untouched
                per_block[f"block{i}.{j}"] = rel_err(g, r)
                if truth is not None:
                    t = outs(truth[i])[j]
                    vs32[f"block{i}.{j}"] = (rel_err(g, t), rel_err(r, t))  # (ours vs fp32, reference bf16 vs fp32)
        out_block = model(**inp, return_dict=False)[0]
    finally:
        vap.uninstall(model)
    res["per_block_max"] = max(per_block.values())
    res["per_block"] = {k: round(v, 5) for k, v in per_block.items()}
    if vs32:
        res["ours_vs_fp32_max"] = max(v[0] for v in vs32.values())
        res["reference_bf16_vs_fp32_max"] = max(v[1] for v in vs32.values())
    res["forward_block_level"] = dict(err=rel_err(out_block, final), cosine=cosine(out_block, final))
    if "processor" in levels:
        vap.install(model, level="processor")
        try:
            out_proc = model(**inp, return_dict=False)[0]
        finally:
            vap.uninstall(model)
        res["forward_processor_level"] = dict(err=rel_err(out_proc, final), cosine=cosine(out_proc, final))
    restored = model(**inp, return_dict=False)[0]
    res["uninstall_restores_bit_exact"] = bool(torch.equal(restored, final))
    res["state_dict_keys_unchanged"] = list(model.state_dict()) == keys
    assert res["per_block_max"] <= tol, f"{family} blocks vs the reference on the GPU: {res}"
    assert res["forward_block_level"]["cosine"] >= 0.999, res
    assert "forward_processor_level" not in res or res["forward_processor_level"]["cosine"] >= 0.999, res
    assert res["uninstall_restores_bit_exact"] and res["state_dict_keys_unchanged"], res
    return res


@torch.no_grad()
def check_wan_denoise(cfg: dict, latent, steps: int = 4, time_it: bool = False, model=None):
    """north_star gate: final latents of a 4-step denoise (CFG 5.0, FlowMatchEuler shift 3) — stock reference vs install() on the same
    instance — cosine >= 0.999.  The loop (vap.denoise.wan_denoise) restates pipeline_wan_i2v_mot.py:801-877 and is the same code for
    both runs; only the transformer's block forwards differ."""
    model = model if model is not None else ref_gpu.build_reference("wan", cfg, seed=7, device=DEV)
    f, h, w = latent
    inp = _inputs("wan", cfg, latent)
    neg = synth.wan_inputs(cfg, f, h, w, seed=5, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(11)
    lat0 = torch.randn((1, 16, f, h, w), generator=g, device=DEV)
    lat_ref = torch.randn((1, 16, f, h, w), generator=g, device=DEV)
    kw = {k: inp[k] for k in ("encoder_hidden_states", "encoder_hidden_states_image", "encoder_hidden_states_mot_ref",
                              "encoder_hidden_states_image_mot_ref", "num_mot_ref")}
    kw_u = dict(kw, encoder_hidden_states=neg["encoder_hidden_states"], encoder_hidden_states_mot_ref=neg["encoder_hidden_states_mot_ref"])
    cond, cond_r = inp["hidden_states"][:, 16:].float(), inp["hidden_states_mot_ref"][:, 16:].float()

    def run():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lat = vap.denoise.wan_denoise(model, lat0, cond, lat_ref, cond_r, kw, kw_u, steps, 3.0, 5.0, cache_context=False)
        e1.record()
        torch.cuda.synchronize()
        return lat, e0.elapsed_time(e1)

    stock, ms_stock = run()
    vap.install(model, level="block")
    try:
        run() if time_it else None  # warm (weight packing, fp32 parameter copies) before the timed run
        ours, ms_ours = run()
    finally:
        vap.uninstall(model)
    res = dict(cosine=cosine(ours, stock), err=rel_err(ours, stock), steps=steps, guided=True)
    if time_it:
        res.update(stock_ms_per_guided_step=ms_stock / steps, installed_ms_per_guided_step=ms_ours / steps)
    assert res["cosine"] >= 0.999, f"wan {steps}-step denoise, install() vs the stock reference on the GPU: {res}"
    return res


@torch.no_grad()
def check_cog_denoise(cfg: dict, latent, steps: int = 4):
    """CogVideoX loop (pipeline_cogvideox_image2video_mot.py:964-1057): B=2 CFG forward per step, CogVideoXDPMScheduler."""
    model = ref_gpu.build_reference("cog", cfg, seed=7, device=DEV)
    f, h, w = latent
    inp = _inputs("cog", cfg, latent, batch=2)
    kw2 = {k: inp[k] for k in ("encoder_hidden_states", "encoder_hidden_states_mot_ref", "image_rotary_emb", "image_rotary_emb_mot_ref", "num_mot_ref")}
    g = torch.Generator(device=DEV).manual_seed(12)
    lat0, img, lat_ref, img_ref = (torch.randn((1, f, 16, h, w), generator=g, device=DEV) for _ in range(4))
    stock = vap.denoise.cog_denoise(model, lat0, img, lat_ref, img_ref, kw2, steps, 6.0, True, 3)
    vap.install(model, level="block")
    try:
        ours = vap.denoise.cog_denoise(model, lat0, img, lat_ref, img_ref, kw2, steps, 6.0, True, 3)
    finally:
        vap.uninstall(model)
    res = dict(cosine=cosine(ours, stock), err=rel_err(ours, stock), steps=steps)
    assert res["cosine"] >= 0.999, f"cog {steps}-step denoise, install() vs the stock reference on the GPU: {res}"
    return res


def check_train_block_level(family: str, cfg: dict, latent):
    """Trainer seam at block level (training.py): only the "_mot_ref" parameters train (sft_trainer/trainer.py:154-164); the loss back-propagates
    (a) through the stock reference (torch autograd everywhere, cuDNN SDPA) and (b) through install(level="block", trainable=True) — fused
    forward, reference block recomputed in the backward with vap_attention_fwd / vap_attention_bwd in the SDPA slot.  Every trainable tensor
    must get a gradient with cosine > 0.99 against the stock one; forward + backward times of both are reported."""
    model = ref_gpu.build_reference(family, cfg, seed=7, device=DEV).train()
    for name, prm in model.named_parameters():
        prm.requires_grad_("_mot_ref" in name)
    inp = _inputs(family, cfg, latent)
    with torch.no_grad():
        shape = model(**inp, return_dict=False)[0].shape
    target = torch.randn(shape, generator=torch.Generator(device=DEV).manual_seed(1), device=DEV)

    def grads():
        model.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = model(**inp, return_dict=False)[0].float()
        torch.nn.functional.mse_loss(out, target).backward()
        e1.record()
        torch.cuda.synchronize()
        return {n: p.grad.float().clone() for n, p in model.named_parameters() if p.grad is not None}, e0.elapsed_time(e1)

    grads()  # warm-up (cuDNN plans, allocator)
    ref, ms_stock = grads()
    vap.install(model, level="block", trainable=True)
    try:
        grads()
        got, ms_ours = grads()
    finally:
        vap.uninstall(model)
    # Per-tensor cosine for every tensor that carries a real gradient (norm >= 1 % of the largest one), and the relative error of the whole
    # gradient.  Tensors below that floor are rounding noise in BOTH runs — e.g. the cross-attention key bias, to which a softmax is invariant
    # up to the RMSNorm that follows it: its cosine flips sign from run to run — and are only required to stay small.
    norms = {n: ref[n].double().norm().item() for n in ref}
    floor = 1e-2 * max(norms.values())
    cos, small = {}, {}
    for n in ref:
        a, b = ref[n].double().flatten(), got[n].double().flatten()
        if norms[n] >= floor:
            cos[n] = (torch.dot(a, b) / (a.norm() * b.norm())).item()
        else:
            small[n] = b.norm().item()
    num = sum((got[n].double() - ref[n].double()).pow(2).sum().item() for n in ref) ** 0.5
    den = sum(ref[n].double().pow(2).sum().item() for n in ref) ** 0.5
    res = dict(trainable_tensors=len(ref), compared_by_cosine=len(cos), same_names=sorted(ref) == sorted(got), worst_cosine=min(cos.values()),
               worst_name=min(cos, key=cos.get), whole_gradient_rel_err=num / den, noise_level_tensors_stay_small=all(v < 2 * floor for v in small.values()),
               stock_fwd_bwd_ms=ms_stock, installed_fwd_bwd_ms=ms_ours)
    assert res["same_names"] and res["worst_cosine"] > 0.99 and res["whole_gradient_rel_err"] < 0.1 and res["noise_level_tensors_stay_small"], res
    return res


# widths of BASELINE.json configs[2] / configs[1]; 3 layers (MoT, plain, MoT), reduced token count
WAN_14B_3L = dict(synth.WAN_14B, num_layers=3, block_idx_with_mot_ref=[0, 2])
COG_5B_3L = dict(synth.COG_5B, num_layers=3, block_idx_with_mot_ref=[0, 1])

CHECKS = {
    # 3 latent frames of the 480x832 grid: 4 680 tokens per stream, J = 9 360
    "ref_wan14b_blocks": lambda: check_blocks("wan", WAN_14B_3L, (3, 60, 104)),
    "ref_wan14b_denoise": lambda: check_wan_denoise(WAN_14B_3L, (3, 60, 104)),
    # CogVideoX-5B widths; the learned positional embedding of the 5B-I2V config pins the latent grid to 13 x 60 x 90
    "ref_cog5b_blocks": lambda: check_blocks("cog", COG_5B_3L, (13, 60, 90)),
    "ref_cog5b_denoise": lambda: check_cog_denoise(COG_5B_3L, (13, 60, 90)),
    # block-level trainer seam: gradients of the expert's parameters, stock autograd vs fused forward + recomputed backward
    "ref_wan14b_train_block": lambda: check_train_block_level("wan", WAN_14B_3L, (3, 60, 104)),
    "ref_cog5b_train_block": lambda: check_train_block_level("cog", dict(synth.COG_5B, num_layers=2, block_idx_with_mot_ref=[0, 1]), (13, 60, 90)),
    # BASELINE.json configs[2] itself: all 40 MoT blocks, 49 frames 480x832 (J = 40 560), one forward + the 4-step guided denoise
    "ref_wan14b_full_cfg3": lambda: check_wan_full(),
}


@torch.no_grad()
def check_wan_full(cfg=None, latent=(13, 60, 104)):
    cfg = cfg or synth.WAN_14B
    model = ref_gpu.build_reference("wan", cfg, seed=1234, device=DEV)
    inp = _inputs("wan", cfg, latent)
    ms_stock, final = ref_gpu.time_forward(model, inp, steps=2, warmup=1)
    kernels = ref_gpu.sdpa_kernels(model, inp)
    vap.install(model, level="block")
    try:
        ms_ours, out = ref_gpu.time_forward(model, inp, steps=2, warmup=1)
    finally:
        vap.uninstall(model)
    res = dict(forward=dict(err=rel_err(out, final), cosine=cosine(out, final)), stock_ms=ms_stock, installed_ms=ms_ours, speedup=ms_stock / ms_ours,
               stock_kernels=kernels)
    assert res["forward"]["cosine"] >= 0.999, res
    res["denoise"] = check_wan_denoise(cfg, latent, time_it=True, model=model)
    return res
