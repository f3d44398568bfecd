"""CPU, authoring container only (skipped where /root/reference does not exist, e.g. on the GPU box): the drop-in boundary on the
REFERENCE's own classes.  The reference's WanTransformer3DMOTModel / CogVideoXTransformer3DMOTModel are imported from
/root/reference, filled with synthetic weights and run on CPU as they are; then `vap_b200.install(model, level=...)` rebinds the
block forwards (B3) or swaps the attention processors + the SDPA slot (B1 + B2) on those very instances — kernels replaced by the
torch stand-ins of tests/cpu_standin_ops.py — and the SAME pipeline-facing call must give the same output within the bf16 gate, and
`uninstall` must restore the reference's behaviour bit for bit.  This is the host-side half of the drop-in claim; the kernels' half
is `-m gpu`."""
import importlib
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VAP_REFERENCE", "/root/reference")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "diffusers", "src")), reason="the reference tree is not present here")


def _worker(family, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        sys.path.insert(0, os.path.join(REF, "diffusers", "src"))
        os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
        sys.dont_write_bytecode = True
        torch.set_grad_enabled(False)
        torch.set_num_threads(4)
        vap = importlib.import_module("video-as-prompt_b200")
        import cpu_standin_ops
        cpu_standin_ops.install(vap)
        if family == "wan":
            from diffusers import WanTransformer3DMOTModel as RefModel
            cfg = dict(vap.synth.WAN_TINY, num_layers=3, block_idx_with_mot_ref=[0, 2])
            inp = vap.synth.wan_inputs(cfg, 3, 16, 24, seed=0)
        else:
            from diffusers import CogVideoXTransformer3DMOTModel as RefModel
            cfg = dict(vap.synth.COG_TINY, num_layers=3, block_idx_with_mot_ref=[0, 2])
            inp = vap.synth.cog_inputs(cfg, 2, 12, 20, seed=0, batch=2)
        model = RefModel(**cfg).to(torch.bfloat16).eval()
        vap.synth.fill_module_(model, seed=7, num_layers=cfg["num_layers"])
        keys = list(model.state_dict())

        def run():
            return model(**inp, return_dict=False)[0].float()

        ref = run()
        res = {}
        for level in ("block", "processor"):
            vap.install(model, level=level)
            if family == "wan":  # the shell's per-forward CPU RoPE build + H2D copy is replaced by the cached device tables
                res[level + "_rope_swapped"] = isinstance(model.rope(inp["hidden_states"]), tuple) and isinstance(model.rope_mot_ref(inp["hidden_states_mot_ref"]), tuple)
            out = run()
            vap.uninstall(model)
            res[level] = ((out - ref).abs().max() / ref.abs().max()).item()
            res[level + "_restored"] = bool(torch.equal(run(), ref))
        res["keys_unchanged"] = list(model.state_dict()) == keys
        q.put(res)
    except Exception:  # noqa: BLE001
        import traceback
        q.put({"error": traceback.format_exc()[-3000:]})


@pytest.mark.parametrize("family", ["wan", "cog"])
def test_install_on_the_reference_model_matches_the_reference(family):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_worker, args=(family, q))
    p.start()
    res = q.get(timeout=600)
    p.join(60)
    assert "error" not in res, res.get("error")
    assert res["block"] < 2e-2 and res["processor"] < 2e-2, res
    assert res["block_restored"] and res["processor_restored"] and res["keys_unchanged"], res
    if family == "wan":
        assert res["block_rope_swapped"] and res["processor_rope_swapped"], res


def _train_worker(family, q, mode="sdpa"):
    """mode "sdpa": the trainer's contract at the SDPA seam (finetrainers/trainer/sft_trainer/trainer.py:154-164, 674-714): only parameters with
    "_mot_ref" in their name train; the loss back-propagates through the reference's own block code and OUR attention.
    mode "block": install(level="block", trainable=True) — the fused forward as the first pass of activation checkpointing, the reference's block
    code recomputed in the backward (video-as-prompt_b200/training.py); also under torch.utils.checkpoint, as the trainer calls the blocks when
    gradient checkpointing is on (positional arguments), and without gradient tracking (plain fused forward)."""
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        sys.path.insert(0, os.path.join(REF, "diffusers", "src"))
        sys.dont_write_bytecode = True
        torch.set_num_threads(4)
        vap = importlib.import_module("video-as-prompt_b200")
        import cpu_standin_ops
        cpu_standin_ops.install(vap)
        if family == "wan":
            from diffusers import WanTransformer3DMOTModel as RefModel
            cfg = dict(vap.synth.WAN_TINY, num_layers=2, block_idx_with_mot_ref=[0, 1])
            inp = vap.synth.wan_inputs(cfg, 2, 8, 8, seed=0)
        else:
            from diffusers import CogVideoXTransformer3DMOTModel as RefModel
            cfg = dict(vap.synth.COG_TINY, num_layers=2, block_idx_with_mot_ref=[0, 1])
            inp = vap.synth.cog_inputs(cfg, 2, 8, 12, seed=0)
        model = RefModel(**cfg).to(torch.bfloat16).train()
        vap.synth.fill_module_(model, seed=7, num_layers=cfg["num_layers"])
        for name, prm in model.named_parameters():
            prm.requires_grad_("_mot_ref" in name)
        target = torch.randn(model(**inp, return_dict=False)[0].shape, generator=torch.Generator().manual_seed(1))

        def grads():
            model.zero_grad(set_to_none=True)
            out = model(**inp, return_dict=False)[0].float()
            torch.nn.functional.mse_loss(out, target).backward()
            return {n: p.grad.float().clone() for n, p in model.named_parameters() if p.grad is not None}

        ref = grads()
        res = {}
        if mode == "sdpa":
            vap.install(model, level="sdpa")
            got = grads()
            vap.uninstall(model)
        else:
            with torch.no_grad():
                out_ref = model(**inp, return_dict=False)[0].float()
            vap.install(model, level="block", trainable=True)
            got = grads()
            with torch.no_grad():  # no gradient tracking: the wrapper is the fused forward itself
                out = model(**inp, return_dict=False)[0].float()
            res["forward_err"] = ((out - out_ref).abs().max() / out_ref.abs().max()).item()
            res["sdpa_slot_restored"] = torch.nn.functional.scaled_dot_product_attention is not vap.sdpa.joint_sdpa
            vap.uninstall(model)
        res.update({"n_ref": len(ref), "n_got": len(got), "same_names": sorted(ref) == sorted(got)})
        # cosine per parameter (bf16 training noise makes max-abs a poor gate for gradients), worst case over all trainable tensors
        cos = {}
        floor = 1e-2 * max(g.double().norm().item() for g in ref.values())  # below it a gradient is rounding noise in both runs (e.g. the key biases)
        for n in ref:
            a, b = ref[n].double().flatten(), got[n].double().flatten()
            if a.norm() >= floor:
                cos[n] = (torch.dot(a, b) / (a.norm() * b.norm())).item()
        res["worst_cosine"] = min(cos.values())
        res["worst_name"] = min(cos, key=cos.get)
        res["attn_grads_present"] = any("attn1_mot_ref.to_q" in n for n in got)
        q.put(res)
    except Exception:  # noqa: BLE001
        import traceback
        q.put({"error": traceback.format_exc()[-3000:]})


@pytest.mark.parametrize("family", ["wan", "cog"])
def test_fused_block_forward_with_recompute_backward_trains_the_reference_model(family):
    """install(level="block", trainable=True): every trainable tensor gets a gradient whose cosine with the stock reference's is > 0.99 (the
    gradients are the reference block's own, evaluated at the saved inputs; the inputs of later blocks differ by the fused path's bf16 noise)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_train_worker, args=(family, q, "block"))
    p.start()
    res = q.get(timeout=600)
    p.join(60)
    assert "error" not in res, res.get("error")
    assert res["same_names"] and res["n_got"] > 0 and res["attn_grads_present"], res
    assert res["worst_cosine"] > 0.99, res
    assert res["forward_err"] < 2e-2 and res["sdpa_slot_restored"], res


@pytest.mark.parametrize("family", ["wan", "cog"])
def test_sdpa_seam_trains_the_reference_model(family):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_train_worker, args=(family, q))
    p.start()
    res = q.get(timeout=600)
    p.join(60)
    assert "error" not in res, res.get("error")
    assert res["same_names"] and res["n_got"] > 0 and res["attn_grads_present"], res
    assert res["worst_cosine"] > 0.99, res
