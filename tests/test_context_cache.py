"""CPU: the context cache of the Wan denoise loop (SURVEY §8f rank 1 — "cache K/V of the constant context across steps and CFG
passes") is pure host logic: with it the loop must launch fewer projections and return the SAME latents bit for bit.  Runs in a
spawned process on the torch stand-in ops (tests/cpu_standin_ops.py), which must not leak into the other CPU tests."""
import importlib
import os
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    vap = importlib.import_module("video-as-prompt_b200")
    import cpu_standin_ops
    cpu_standin_ops.install(vap)
    torch.set_num_threads(2)
    cfg = dict(vap.synth.WAN_TINY, num_layers=3, block_idx_with_mot_ref=[0, 2])
    model = vap.WanTransformer3DMOTModel(**cfg).to(torch.bfloat16).eval()
    vap.synth.fill_module_(model, seed=7, num_layers=3)
    f, h, w = 2, 8, 8
    inp = vap.synth.wan_inputs(cfg, f, h, w, seed=0)
    neg = vap.synth.wan_inputs(cfg, f, h, w, seed=1)
    keys = ("encoder_hidden_states", "encoder_hidden_states_image", "encoder_hidden_states_mot_ref", "encoder_hidden_states_image_mot_ref", "num_mot_ref")
    kw = {k: inp[k] for k in keys}
    kw_u = dict(kw, encoder_hidden_states=neg["encoder_hidden_states"], encoder_hidden_states_mot_ref=neg["encoder_hidden_states_mot_ref"])
    g = torch.Generator().manual_seed(3)
    lat0, lat_ref = torch.randn((1, 16, f, h, w), generator=g), torch.randn((1, 16, f, h, w), generator=g)
    cond, cond_ref = inp["hidden_states"][:, 16:].float(), inp["hidden_states_mot_ref"][:, 16:].float()

    counts = {}
    real_linear = vap.ops.linear

    def counting_linear(x, weight, *a, **k):
        counts[tuple(weight.shape) + (x.shape[-2],)] = counts.get(tuple(weight.shape) + (x.shape[-2],), 0) + 1
        return real_linear(x, weight, *a, **k)

    vap.ops.linear = counting_linear
    res = {}
    outs = []
    for cache in (False, True):
        counts.clear()
        outs.append(vap.denoise.wan_denoise(model, lat0.clone(), cond, lat_ref, cond_ref, kw, kw_u, 3, 3.0, 5.0, cache_context=cache))
        # context K/V projections: packed [2 * inner, inner] weights applied to 512 text rows / 257 image rows
        res["kv_launches_cache_%d" % cache] = sum(n for (N, K, M), n in counts.items() if M in (512, 257))
    res["bit_exact"] = bool(torch.equal(outs[0], outs[1]))
    # classifier-free guidance as one B = 2 forward per step == two B = 1 forwards (same arithmetic per sample), with and without the cache
    for cache in (False, True):
        counts.clear()
        lat = vap.denoise.wan_denoise(model, lat0.clone(), cond, lat_ref, cond_ref, kw, kw_u, 3, 3.0, 5.0, cache_context=cache, batch_cfg=True)
        res["batch_cfg_exact_cache_%d" % cache] = bool(torch.equal(lat, outs[0]))
        res["batch_cfg_launches_cache_%d" % cache] = sum(counts.values())
    counts.clear()
    vap.denoise.wan_denoise(model, lat0.clone(), cond, lat_ref, cond_ref, kw, kw_u, 3, 3.0, 5.0, cache_context=False)
    res["sequential_launches"] = sum(counts.values())
    # the fused CFG + scheduler step (stand-in of vap_cfg_flow_match_step's contract) reproduces the torch expression loop bit for bit
    fused = vap.denoise.wan_denoise(model, lat0.clone(), cond, lat_ref, cond_ref, kw, kw_u, 3, 3.0, 5.0, fused_step=True)
    res["fused_step_exact"] = bool(torch.equal(fused, outs[0]))
    fused1 = vap.denoise.wan_denoise(model, lat0.clone(), cond, lat_ref, cond_ref, kw, None, 2, 3.0, 5.0, fused_step=True)
    plain1 = vap.denoise.wan_denoise(model, lat0.clone(), cond, lat_ref, cond_ref, kw, None, 2, 3.0, 5.0)
    res["fused_step_no_cfg_exact"] = bool(torch.equal(fused1, plain1))
    # dead reference-stream work in the last MoT block: skipping it must not change the model output
    with torch.no_grad():
        full = model(**inp, return_dict=False)[0]
        counts.clear()
        model(**inp, return_dict=False)
        n_full = sum(counts.values())
        model.skip_dead_reference_work = True
        counts.clear()
        lean = model(**inp, return_dict=False)[0]
        n_lean = sum(counts.values())
        model.skip_dead_reference_work = False
    res["dead_skip_exact"] = bool(torch.equal(full, lean))
    res["dead_skip_linears"] = (n_full, n_lean)
    res["left_over_entries"] = sum(1 for m in model.modules() if "_vap_ctx_cache" in m.__dict__)
    # the cache keys on tensor identity AND version: an in-place edit of the conditioning must miss
    with vap.wan.context_cache():
        a = model(**inp, return_dict=False)[0]
        inp["encoder_hidden_states"].mul_(0.5)
        b = model(**inp, return_dict=False)[0]
    with torch.no_grad():
        vap.wan.clear_context_cache(model)
        b_ref = model(**inp, return_dict=False)[0]
    res["version_miss"] = bool(torch.equal(b, b_ref)) and not bool(torch.equal(a, b))
    q.put(res)


def test_context_cache_is_bit_exact_and_skips_the_constant_projections():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_worker, args=(q,))
    p.start()
    res = q.get(timeout=600)
    p.join(60)
    assert res["bit_exact"], res
    # 3 steps x 2 guidance passes x (3 target-stream + 2 expert-stream cross-attentions) x (text + image) = 60 without the cache;
    # with it every (module, context) pair is projected once: 5 modules x 2 contexts... the image tokens are shared by both passes but
    # the key is the concatenated [image | text] context, so 2 contexts x 2 projections x 5 modules = 20
    assert res["kv_launches_cache_0"] == 60 and res["kv_launches_cache_1"] == 20, res
    assert res["left_over_entries"] == 0 and res["version_miss"], res
    assert res["batch_cfg_exact_cache_0"] and res["batch_cfg_exact_cache_1"], res
    assert res["fused_step_exact"] and res["fused_step_no_cfg_exact"], res
    assert res["batch_cfg_launches_cache_0"] < res["sequential_launches"], res
    # the expert stream of the last MoT block loses its O-projection, 4 cross-attention projections and 2 FFN GEMMs
    assert res["dead_skip_exact"] and res["dead_skip_linears"][0] - res["dead_skip_linears"][1] == 7, res


def test_cache_entries_die_with_their_inputs():
    """An entry holds only weak references to its input tensors: once the input is gone (a shell that builds a fresh context tensor every forward,
    like the reference's) the entry is purged at the next lookup — nothing stays pinned, and a new tensor at the same address cannot hit it."""
    import gc
    import importlib
    vap = importlib.import_module("video-as-prompt_b200")
    wan = vap.wan
    owner = torch.nn.Linear(2, 2)
    calls = []

    def compute():
        calls.append(1)
        return torch.zeros(1)

    with wan.context_cache():
        a = torch.randn(4, 8)
        wan._cached(owner, "kv", (a, None), compute)
        wan._cached(owner, "kv", (a, None), compute)
        assert len(calls) == 1 and len(owner.__dict__["_vap_ctx_cache"]["kv"]) == 1
        del a
        gc.collect()
        b = torch.randn(4, 8)  # may or may not reuse a's storage: either way it must miss
        wan._cached(owner, "kv", (b, None), compute)
        assert len(calls) == 2 and len(owner.__dict__["_vap_ctx_cache"]["kv"]) == 1  # the dead entry was purged, not kept beside the new one
