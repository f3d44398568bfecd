"""Generate tests/golden/*.pt from the REFERENCE modules (run in the authoring container only; needs /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

For each tiny configuration it (1) builds the reference's own WanTransformer3DMOTModel / CogVideoXTransformer3DMOTModel,
(2) overwrites every parameter with the deterministic synthetic value of its name (video-as-prompt_b200/synth.py),
(3) runs it in bf16 on CPU with hooks recording each block's inputs and outputs, (4) asserts that the oracle restatement
reproduces every recorded tensor BIT-EXACTLY, and (5) stores inputs-by-seed + expected tensors as small fixtures.
The fixtures are what pins the oracle (the reference has no tests or golden vectors for the MoT path).
"""
from __future__ import annotations

import importlib
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VAP_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "diffusers", "src"))
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")

import torch  # noqa: E402

torch.set_grad_enabled(False)

from oracle import cog_oracle, denoise, wan_oracle  # noqa: E402

synth = importlib.import_module("video-as-prompt_b200.synth")
GOLD = os.path.join(ROOT, "tests", "golden")


def _hook_blocks(blocks, names_in, names_out):
    rec = {}

    def pre(i):
        def f(mod, args, kwargs):
            rec[i] = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in kwargs.items() if k in names_in}
        return f

    def post(i):
        def f(mod, args, kwargs, out):
            for n, t in zip(names_out, out):
                rec[i][n] = None if t is None else t.clone()
        return f

    handles = []
    for i, b in enumerate(blocks):
        handles.append(b.register_forward_pre_hook(pre(i), with_kwargs=True))
        handles.append(b.register_forward_hook(post(i), with_kwargs=True))
    return rec, handles


def _eq(a, b, what):
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    if not torch.equal(a, b):
        d = (a.float() - b.float()).abs().max().item()
        raise AssertionError(f"oracle != reference for {what}: max abs diff {d}")


def gen_wan():
    from diffusers import WanTransformer3DMOTModel
    from diffusers.schedulers import FlowMatchEulerDiscreteScheduler

    cfg = dict(synth.WAN_TINY, num_layers=3, block_idx_with_mot_ref=[0, 2])  # MoT, plain, MoT
    model = WanTransformer3DMOTModel(**cfg).to(torch.bfloat16).eval()
    synth.fill_module_(model, seed=1, num_layers=cfg["num_layers"])
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    shapes = {k: list(v.shape) for k, v in sd.items()}
    with open(os.path.join(GOLD, "wan_tiny_keys.json"), "w") as f:
        json.dump(dict(config={k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}, shapes=shapes), f, indent=0)

    frames, h, w = 3, 16, 24  # S = 3*8*12 = 288 tokens per stream (not a multiple of 128)
    inp = synth.wan_inputs(cfg, frames, h, w, seed=3)
    rec, handles = _hook_blocks(model.blocks, {"hidden_states", "encoder_hidden_states", "temb", "hidden_states_mot_ref",
                                               "encoder_hidden_states_mot_ref", "temb_mot_ref"}, ["out", "out_ref"])
    ref_out = model(**inp, return_dict=False)[0]
    for hd in handles:
        hd.remove()
    bio = {}
    ora_out = wan_oracle.wan_forward(sd, cfg, **{k: v for k, v in inp.items()}, block_io=bio)
    _eq(ora_out, ref_out, "wan final output")
    for i in rec:
        _eq(bio[i]["out"], rec[i]["out"], f"wan block {i} out")
        _eq(bio[i]["out_ref"], rec[i]["out_ref"], f"wan block {i} out_ref")
        _eq(bio[i]["x"], rec[i]["hidden_states"], f"wan block {i} in")
    # rope tables: reference modules vs oracle
    _eq(wan_oracle.wan_rope(128, cfg["patch_size"], 1024, (frames, h, w), ref=False), model.rope(inp["hidden_states"]), "wan rope")
    _eq(wan_oracle.wan_rope(128, cfg["patch_size"], 1024, (frames, h, w), ref=True), model.rope_mot_ref(inp["hidden_states_mot_ref"]), "wan rope ref")
    shared = ("encoder_hidden_states", "encoder_hidden_states_mot_ref", "temb", "temb_mot_ref")  # identical for every block
    fixture = dict(cfg=cfg, weight_seed=1, input_seed=3, latent=(frames, h, w), final=ref_out, shared={k: rec[0][k] for k in shared},
                   blocks={i: {k: v for k, v in rec[i].items() if k not in shared} for i in rec})

    # 4-step denoise with CFG through the reference scheduler class
    sched = FlowMatchEulerDiscreteScheduler(shift=3.0)
    sched.set_timesteps(4)
    ts_o, sig_o = denoise.flow_match_schedule(4, 3.0)
    _eq(ts_o, sched.timesteps, "timesteps"), _eq(sig_o, sched.sigmas, "sigmas")
    g = torch.Generator().manual_seed(11)
    lat0 = torch.randn((1, 16, frames, h, w), generator=g)
    lat_ref = torch.randn((1, 16, frames, h, w), generator=g)
    cond = inp["hidden_states"][:, 16:].float()
    cond_ref = inp["hidden_states_mot_ref"][:, 16:].float()
    kw = {k: inp[k] for k in ("encoder_hidden_states", "encoder_hidden_states_image", "encoder_hidden_states_mot_ref",
                              "encoder_hidden_states_image_mot_ref", "num_mot_ref")}
    neg = synth.wan_inputs(cfg, frames, h, w, seed=4)
    kw_u = dict(kw, encoder_hidden_states=neg["encoder_hidden_states"], encoder_hidden_states_mot_ref=neg["encoder_hidden_states_mot_ref"])
    lat = lat0.clone()
    for i, t in enumerate(sched.timesteps):  # the reference pipeline's loop body (pipeline_wan_i2v_mot.py:801-877)
        x_in = torch.cat([lat, cond], dim=1).to(torch.bfloat16)
        x_ref = torch.cat([lat_ref, cond_ref], dim=1).to(torch.bfloat16)
        ts_ref = (sched.timesteps[-1] * 0 + 1).unsqueeze(0).unsqueeze(0).repeat(1, 1)
        n_c = model(hidden_states=x_in, timestep=t.expand(1), hidden_states_mot_ref=x_ref, timestep_list_mot_ref=ts_ref, return_dict=False, **kw)[0]
        n_u = model(hidden_states=x_in, timestep=t.expand(1), hidden_states_mot_ref=x_ref, timestep_list_mot_ref=ts_ref, return_dict=False, **kw_u)[0]
        noise = n_u + 5.0 * (n_c - n_u)
        lat = sched.step(noise, t, lat, return_dict=False)[0]
    fwd = lambda **k: wan_oracle.wan_forward(sd, cfg, **k)  # noqa: E731
    lat_o, _ = denoise.wan_denoise(fwd, lat0.clone(), cond, lat_ref, cond_ref, kw, kw_u, 4, 3.0, 5.0)
    _eq(lat_o, lat, "wan 4-step latents")
    fixture["denoise"] = dict(seed=11, steps=4, shift=3.0, guidance=5.0, neg_seed=4, final_latents=lat)
    torch.save(fixture, os.path.join(GOLD, "wan_tiny.pt"))
    print("wan_tiny.pt ok:", {i: list(rec[i]) for i in rec})


def gen_cog():
    from diffusers import CogVideoXTransformer3DMOTModel
    from diffusers.models.embeddings import get_3d_rotary_pos_embed

    def rope_fn(D, crops, grid, T, device=None, **kw):
        return get_3d_rotary_pos_embed(D, crops, grid, T, device=device, **kw)

    cfg = dict(synth.COG_TINY, num_layers=3, block_idx_with_mot_ref=[0, 1])  # MoT, MoT, plain (like the 5B: last block plain)
    model = CogVideoXTransformer3DMOTModel(**cfg).to(torch.bfloat16).eval()
    synth.fill_module_(model, seed=2, num_layers=cfg["num_layers"])
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    with open(os.path.join(GOLD, "cog_tiny_keys.json"), "w") as f:
        json.dump(dict(config=cfg, shapes={k: list(v.shape) for k, v in sd.items()}), f, indent=0)
    names_in = {"hidden_states", "encoder_hidden_states", "temb", "hidden_states_mot_ref", "encoder_hidden_states_mot_ref", "temb_mot_ref",
                "temb_list_mot_ref"}
    names_out = ["out_v", "out_e", "out_v_ref", "out_e_ref"]
    fixtures = {}
    for tag, (frames, h, w, nref, multi) in {"small": (3, 12, 16, 1, False), "multi": (2, 12, 16, 2, True)}.items():
        inp = synth.cog_inputs(cfg, frames, h, w, seed=5, num_mot_ref=nref, rope_fn=rope_fn)
        # our own table builder must agree with the reference's
        ours = importlib.import_module("video-as-prompt_b200.rope").get_3d_rotary_pos_embed
        mine = synth.cog_inputs(cfg, frames, h, w, seed=5, num_mot_ref=nref, rope_fn=ours)
        for k in ("image_rotary_emb", "image_rotary_emb_mot_ref"):
            _eq(mine[k][0], inp[k][0], k + " cos"), _eq(mine[k][1], inp[k][1], k + " sin")
            _eq(cog_oracle.cog_rope_3d(64, ((0, 0), (h // 2, w // 2)), (h // 2, w // 2), frames, mot_num=nref if "ref" in k else 0)[0], inp[k][0], k)
        if multi:
            inp["timestep_list_mot_ref"] = [torch.full((1,), 300.0 + 100 * i) for i in range(nref)]
        rec, handles = _hook_blocks(model.transformer_blocks, names_in, names_out)
        ref_out = model(**inp, return_dict=False)[0]
        for hd in handles:
            hd.remove()
        bio = {}
        ora_out = cog_oracle.cog_forward(sd, cfg, **inp, block_io=bio)
        _eq(ora_out, ref_out, f"cog[{tag}] final")
        for i in rec:
            for n in names_out:
                if rec[i][n] is not None:
                    _eq(bio[i][n], rec[i][n], f"cog[{tag}] block {i} {n}")
        shared = ("temb", "temb_mot_ref", "temb_list_mot_ref")
        fixtures[tag] = dict(latent=(frames, h, w), num_mot_ref=nref, multi=multi, input_seed=5, final=ref_out,
                             shared={k: rec[0].get(k) for k in shared},
                             blocks={i: {k: v for k, v in rec[i].items() if k not in shared} for i in rec},
                             timestep_list=[300.0 + 100 * i for i in range(nref)] if multi else None)
    # BASELINE.json config #1: 4 latent frames, 30x45 patch grid (60x90 latent) — final output only (block tensors are too big to commit)
    inp = synth.cog_inputs(cfg, 4, 60, 90, seed=6, rope_fn=rope_fn)
    ref_out = model(**inp, return_dict=False)[0]
    _eq(cog_oracle.cog_forward(sd, cfg, **inp), ref_out, "cog config#1 final")
    fixtures["config1"] = dict(latent=(4, 60, 90), num_mot_ref=1, multi=False, input_seed=6, final=ref_out)
    # 4-step denoise with classifier-free guidance (one B=2 forward per step), dynamic guidance scale and the reference's own
    # CogVideoXDPMScheduler in the configuration convert_cogvideox_to_diffusers.py:312-326 writes for the 5B model
    from diffusers.schedulers import CogVideoXDPMScheduler
    sched = CogVideoXDPMScheduler(snr_shift_scale=1.0, beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085, clip_sample=False,
                                  num_train_timesteps=1000, prediction_type="v_prediction", rescale_betas_zero_snr=True, set_alpha_to_one=True,
                                  timestep_spacing="trailing")
    steps, gscale, noise_seed = 4, 6.0, 21
    sched.set_timesteps(steps)
    ac_o, ts_o = denoise.cog_dpm_tables(steps)
    _eq(ts_o, sched.timesteps, "cog dpm timesteps"), _eq(ac_o, sched.alphas_cumprod, "cog dpm alphas_cumprod")
    frames, h, w = 3, 12, 16
    inp = synth.cog_inputs(cfg, frames, h, w, seed=7, batch=2, rope_fn=rope_fn)  # batch 0 = negative prompt, 1 = prompt
    kw2 = {k: inp[k] for k in ("encoder_hidden_states", "encoder_hidden_states_mot_ref", "image_rotary_emb", "image_rotary_emb_mot_ref", "num_mot_ref")}
    g = torch.Generator().manual_seed(12)
    lat0, img, lat_ref, img_ref = (torch.randn((1, frames, 16, h, w), generator=g) for _ in range(4))
    lat, old = lat0.to(torch.bfloat16), None
    gen = torch.Generator().manual_seed(noise_seed)
    for i, t in enumerate(sched.timesteps):  # the reference pipeline's loop body (pipeline_cogvideox_image2video_mot.py:964-1057)
        x = torch.cat([torch.cat([lat] * 2), torch.cat([img] * 2)], dim=2).to(torch.bfloat16)
        xr = torch.cat([torch.cat([lat_ref] * 2), torch.cat([img_ref] * 2)], dim=2).to(torch.bfloat16)
        noise = model(hidden_states=x, hidden_states_mot_ref=xr, timestep=t.expand(2), return_dict=False, **kw2)[0].float()
        gs = 1 + gscale * ((1 - math.cos(math.pi * ((steps - t.item()) / steps) ** 5.0)) / 2)
        n_u, n_c = noise.chunk(2)
        noise = n_u + gs * (n_c - n_u)
        lat, old = sched.step(noise, old, t, sched.timesteps[i - 1] if i > 0 else None, lat, generator=gen, return_dict=False)
        lat = lat.to(torch.bfloat16)
    fwd = lambda **k: cog_oracle.cog_forward(sd, cfg, **k)  # noqa: E731
    lat_o, _ = denoise.cog_denoise(fwd, lat0.clone(), img, lat_ref, img_ref, kw2, steps, gscale, True, noise_seed)
    _eq(lat_o, lat, "cog 4-step latents")
    fixtures["denoise"] = dict(latent=(frames, h, w), input_seed=7, latent_seed=12, noise_seed=noise_seed, steps=steps, guidance=gscale, dynamic_cfg=True,
                               final_latents=lat)
    torch.save(dict(cfg=cfg, weight_seed=2, cases=fixtures), os.path.join(GOLD, "cog_tiny.pt"))
    print("cog_tiny.pt ok")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    gen_wan()
    gen_cog()
