"""CPU: the oracle restatement against the committed golden fixtures that were recorded from the reference modules
(oracle/gen_golden.py).  Same torch build => bit-exact; a different CPU/torch may differ in the last bf16 ulp, so the
assertion is `rel_err <= 4e-3` with exact equality reported when it holds."""
import importlib
import os
import sys

import pytest
import torch

from conftest import rel_err

from oracle import cog_oracle, denoise, wan_oracle

synth = importlib.import_module("video-as-prompt_b200.synth")
TOL = 4e-3


def _wan_sd(g):
    import json
    keys = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wan_tiny_keys.json")))
    return synth.synth_state_dict(keys["shapes"], seed=g["weight_seed"], num_layers=g["cfg"]["num_layers"])


def _cog_sd(g):
    import json
    keys = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cog_tiny_keys.json")))
    return synth.synth_state_dict(keys["shapes"], seed=g["weight_seed"], num_layers=g["cfg"]["num_layers"])


def test_wan_forward_and_blocks(wan_golden):
    g = wan_golden
    sd, cfg = _wan_sd(g), g["cfg"]
    inp = synth.wan_inputs(cfg, *g["latent"], seed=g["input_seed"])
    bio = {}
    with torch.no_grad():
        out = wan_oracle.wan_forward(sd, cfg, **inp, block_io=bio)
    assert rel_err(out, g["final"]) <= TOL
    for i, blk in g["blocks"].items():
        assert rel_err(bio[i]["out"], blk["out"]) <= TOL
        assert rel_err(bio[i]["out_ref"], blk["out_ref"]) <= TOL


def test_wan_blocks_teacher_forced(wan_golden):
    """Each block alone on the recorded inputs of that block (the protocol of SURVEY.md §8c)."""
    g = wan_golden
    sd, cfg = _wan_sd(g), g["cfg"]
    f, h, w = g["latent"]
    fr = wan_oracle.wan_rope(128, cfg["patch_size"], 1024, (f, h, w), ref=False)
    fr_r = wan_oracle.wan_rope(128, cfg["patch_size"], 1024, (f, h, w), ref=True)
    sh = g["shared"]
    with torch.no_grad():
        for i, blk in g["blocks"].items():
            x, xr = wan_oracle.wan_block(sd, f"blocks.{i}", cfg, i in cfg["block_idx_with_mot_ref"], blk["hidden_states"],
                                         sh["encoder_hidden_states"], sh["temb"], fr, blk["hidden_states_mot_ref"],
                                         sh["encoder_hidden_states_mot_ref"], sh["temb_mot_ref"], fr_r, 1)
            assert rel_err(x, blk["out"]) <= TOL and rel_err(xr, blk["out_ref"]) <= TOL


def test_wan_ref_rope_is_temporally_shifted():
    """Reference-video tokens sit at temporal positions -F..-1 (transformer_wan_mot.py:437): their t-angles are the
    target's shifted by -F, the h/w angles are identical."""
    F_, h, w = 3, 4, 6
    a = wan_oracle.wan_rope(128, (1, 2, 2), 1024, (F_, h, w), ref=False)[0, 0]
    b = wan_oracle.wan_rope(128, (1, 2, 2), 1024, (F_, h, w), ref=True)[0, 0]
    assert a.dtype == torch.complex128 and a.shape == (F_ * 2 * 3, 64)
    t_dim = 128 - 4 * (128 // 6)
    inv = 1.0 / (10000.0 ** (torch.arange(0, t_dim, 2, dtype=torch.float64) / t_dim))
    shift = torch.polar(torch.ones_like(inv), -F_ * inv)
    assert torch.allclose(b[:, : t_dim // 2], a[:, : t_dim // 2] * shift, atol=1e-12)
    assert torch.equal(b[:, t_dim // 2:], a[:, t_dim // 2:])


def test_flow_match_schedule_known_values():
    ts, sig = denoise.flow_match_schedule(4, shift=1.0)
    assert torch.allclose(ts, torch.tensor([1000.0, 667.0, 334.0, 1.0]))
    assert sig[-1] == 0 and sig.shape == (5,)
    ts3, sig3 = denoise.flow_match_schedule(4, shift=3.0)
    assert torch.all(sig3[:-1] >= sig[:-1] - 1e-6) and torch.all(sig3[:-1] <= 1.0)


def test_wan_four_step_denoise(wan_golden):
    g = wan_golden
    sd, cfg, dn = _wan_sd(g), g["cfg"], g["denoise"]
    f, h, w = g["latent"]
    inp = synth.wan_inputs(cfg, f, h, w, seed=g["input_seed"])
    neg = synth.wan_inputs(cfg, f, h, w, seed=dn["neg_seed"])
    gen = torch.Generator().manual_seed(dn["seed"])
    lat0 = torch.randn((1, 16, f, h, w), generator=gen)
    lat_ref = torch.randn((1, 16, f, h, w), generator=gen)
    kw = {k: inp[k] for k in ("encoder_hidden_states", "encoder_hidden_states_image", "encoder_hidden_states_mot_ref",
                              "encoder_hidden_states_image_mot_ref", "num_mot_ref")}
    kw_u = dict(kw, encoder_hidden_states=neg["encoder_hidden_states"], encoder_hidden_states_mot_ref=neg["encoder_hidden_states_mot_ref"])
    with torch.no_grad():
        lat, _ = denoise.wan_denoise(lambda **k: wan_oracle.wan_forward(sd, cfg, **k), lat0, inp["hidden_states"][:, 16:].float(), lat_ref,
                                     inp["hidden_states_mot_ref"][:, 16:].float(), kw, kw_u, dn["steps"], dn["shift"], dn["guidance"])
    assert rel_err(lat, dn["final_latents"]) <= 2e-2


@pytest.mark.parametrize("case", ["small", "multi", "config1"])
def test_cog_forward_and_blocks(cog_golden, case):
    g = cog_golden
    c = g["cases"][case]
    sd, cfg = _cog_sd(g), g["cfg"]
    inp = synth.cog_inputs(cfg, *c["latent"], seed=c["input_seed"], num_mot_ref=c["num_mot_ref"])
    if c["multi"]:
        inp["timestep_list_mot_ref"] = [torch.full((1,), t) for t in c["timestep_list"]]
    bio = {}
    with torch.no_grad():
        out = cog_oracle.cog_forward(sd, cfg, **inp, block_io=bio)
    assert rel_err(out, c["final"]) <= TOL
    for i, blk in c.get("blocks", {}).items():
        for n in ("out_v", "out_e", "out_v_ref", "out_e_ref"):
            if blk[n] is not None:
                assert rel_err(bio[i][n], blk[n]) <= TOL, (i, n)


def test_cog_ref_rope_negative_positions():
    """continous_negative: reference frames at linspace(-n*T, -1, n*T) (embeddings.py:871-881)."""
    T, gh, gw = 3, 2, 2
    cos_t, sin_t = cog_oracle.cog_rope_3d(64, ((0, 0), (gh, gw)), (gh, gw), T)
    cos_r, sin_r = cog_oracle.cog_rope_3d(64, ((0, 0), (gh, gw)), (gh, gw), T, mot_num=1)
    assert cos_t.shape == (T * gh * gw, 64) and cos_r.shape == cos_t.shape
    # spatial part identical, temporal part differs; first ref frame is at t=-3 -> angle -3*theta_0 = -3 on channel 0
    assert torch.equal(cos_t[:, 16:], cos_r[:, 16:])
    assert torch.allclose(sin_r[0, 0], torch.sin(torch.tensor(-3.0)))
    assert torch.allclose(sin_r[-1, 0], torch.sin(torch.tensor(-1.0)))


def test_cog_four_step_dpm_denoise(cog_golden):
    """CogVideoXDPMScheduler loop restatement (oracle/denoise.py) against the latents recorded from the reference scheduler +
    reference transformer (4 steps, CFG as one B=2 forward, dynamic guidance, two noise draws per 2nd-order step)."""
    g = cog_golden
    sd, cfg, dn = _cog_sd(g), g["cfg"], g["cases"]["denoise"]
    f, h, w = dn["latent"]
    inp = synth.cog_inputs(cfg, f, h, w, seed=dn["input_seed"], batch=2)
    kw2 = {k: inp[k] for k in ("encoder_hidden_states", "encoder_hidden_states_mot_ref", "image_rotary_emb", "image_rotary_emb_mot_ref", "num_mot_ref")}
    gen = torch.Generator().manual_seed(dn["latent_seed"])
    lat0, img, lat_ref, img_ref = (torch.randn((1, f, 16, h, w), generator=gen) for _ in range(4))
    with torch.no_grad():
        lat, preds = denoise.cog_denoise(lambda **k: cog_oracle.cog_forward(sd, cfg, **k), lat0, img, lat_ref, img_ref, kw2, dn["steps"], dn["guidance"],
                                         dn["dynamic_cfg"], dn["noise_seed"])
    assert len(preds) == dn["steps"]
    assert rel_err(lat, dn["final_latents"]) <= TOL


def test_cog_dpm_schedule_matches_product():
    """The product-side schedule / step (video-as-prompt_b200/denoise.py) and the oracle's are independent restatements."""
    vd = importlib.import_module("video-as-prompt_b200.denoise")
    for n in (4, 10, 50):
        ac_o, ts_o = denoise.cog_dpm_tables(n)
        ac_p, ts_p = vd.cog_dpm_schedule(n)
        assert torch.equal(ts_o, ts_p) and torch.allclose(ac_o, ac_p, rtol=0, atol=0)
    ac, ts = denoise.cog_dpm_tables(4)
    g = torch.Generator().manual_seed(0)
    x = torch.randn((1, 2, 16, 4, 4), generator=g).to(torch.bfloat16)
    v = torch.randn((1, 2, 16, 4, 4), generator=g)
    old = None
    xo, xp, oo, op = x, x, None, None
    for i, t in enumerate(ts.tolist()):
        go, gp = torch.Generator().manual_seed(5 + i), torch.Generator().manual_seed(5 + i)
        xo, oo = denoise.cog_dpm_step(ac, 4, v, oo, t, ts[i - 1].item() if i > 0 else None, xo, go)
        xp, op = vd.cog_dpm_step(ac, 4, v, op, t, ts[i - 1].item() if i > 0 else None, xp, gp)
        assert torch.equal(xo, xp) and torch.equal(oo, op)
        xo, xp = xo.to(torch.bfloat16), xp.to(torch.bfloat16)
