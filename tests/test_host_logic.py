"""CPU: the C-ABI library loads and exports every declared symbol, the stand-alone shells mirror the reference's
state_dict, RoPE table builders agree with the oracle, argument validation / error behaviour, install() seams.
No kernel is launched here (there is no GPU)."""
import importlib
import json
import os
import re
import sys

import pytest
import torch

from oracle import cog_oracle, wan_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vap = importlib.import_module("video-as-prompt_b200")


def test_library_exports_every_header_symbol():
    header = open(os.path.join(ROOT, "include", "vap_b200.h")).read()
    declared = set(re.findall(r"\b(vap_[a-z0-9_]+)\s*\(", header))
    assert {"vap_attention_fwd", "vap_gemm_bf16", "vap_adaln_layernorm", "vap_qk_norm_rope", "vap_ulysses_pack"} <= declared
    lib = vap._lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vap_b200.h but not exported by libvap_b200.so"
    assert declared == set(vap._lib.SIGNATURES), "ctypes SIGNATURES out of sync with the header"
    assert lib.vap_version() == int(re.search(r"#define VAP_B200_VERSION (\d+)", header).group(1))
    assert lib.vap_last_error() is not None


def test_alias_import():
    import vap_b200
    assert vap_b200 is vap and hasattr(vap_b200, "install")


def test_c_abi_rejects_bad_arguments_without_a_gpu():
    lib = vap._lib.load()
    assert lib.vap_gemm_bf16(0, 0, 0, 0, 0, 0, 1, 8, 8, 0, 0, 0, 0, 0, 0, 1, 0) == -1
    assert b"null tensor" in lib.vap_last_error()
    assert lib.vap_attention_fwd(16, 16, 16, 16, 0, 1, 1, 8, 8, 96, *([8] * 12), 1.0, 0) == -1
    assert b"head_dim" in lib.vap_last_error()
    assert lib.vap_adaln_layernorm(16, 16, 4, 12, 16, 16, 0, 0, 0, 0, 0, 1, 1e-6, 0, 0) == -1
    assert lib.vap_qk_norm_rope(16, 16, 4, 2, 48, 96, 16, 0, 16, 0, 0, 0, 4, 0, 0, 1e-6, 0, 0) == -1


def test_ops_fail_loudly_on_cpu_tensors():
    x = torch.zeros(4, 256, dtype=torch.bfloat16)
    with pytest.raises(vap.VapError, match="no CPU fallback"):
        vap.ops.linear(x, torch.zeros(256, 256, dtype=torch.bfloat16))
    with pytest.raises(vap.VapError):
        vap.ops.adaln_layernorm(x, eps=1e-6, rounding=0)
    with pytest.raises(vap.VapError):
        vap.ops.attention(torch.zeros(1, 1, 8, 128, dtype=torch.bfloat16), torch.zeros(1, 1, 8, 128, dtype=torch.bfloat16),
                          torch.zeros(1, 1, 8, 128, dtype=torch.bfloat16))
    with pytest.raises(TypeError):
        vap.ops._need_cuda_bf16("nope", "x")
    # forward-only kernels: a tensor that wants gradients is refused instead of silently losing its graph (trainer drop-in is open)
    w = torch.zeros(256, 256, dtype=torch.bfloat16, requires_grad=True)
    with pytest.raises(vap.VapError, match="inference-only"):
        vap.ops._need_cuda_bf16(w, "weight")
    with pytest.raises(vap.VapError, match="inference-only"):
        vap.ops.linear(w, w)
    with torch.no_grad(), pytest.raises(vap.VapError, match="no CPU fallback"):
        vap.ops.linear(w, w)


def test_joint_sdpa_constraints():
    q = torch.zeros(1, 2, 8, 128, dtype=torch.bfloat16)
    for kw in (dict(attn_mask=torch.zeros(8, 8)), dict(dropout_p=0.1), dict(is_causal=True), dict(enable_gqa=True)):
        with pytest.raises(ValueError):
            vap.joint_sdpa(q, q, q, **kw)
    with pytest.raises(ValueError, match="bfloat16"):
        vap.joint_sdpa(q.float(), q.float(), q.float())
    with pytest.raises(ValueError, match="head_dim"):
        vap.joint_sdpa(q[..., :96], q[..., :96], q[..., :96])
    # non-strict global patch (trainer use): calls outside the envelope reach the ORIGINAL torch SDPA, calls inside never do
    import torch.nn.functional as F
    orig = F.scaled_dot_product_attention
    vap.sdpa.patch_scaled_dot_product_attention(strict=False)
    try:
        x = torch.randn(1, 2, 8, 96)
        assert torch.allclose(F.scaled_dot_product_attention(x, x, x, is_causal=True), orig(x, x, x, is_causal=True))
        with pytest.raises(vap.VapError, match="no CPU fallback"):
            F.scaled_dot_product_attention(q, q, q)  # inside the envelope: the kernel or nothing
    finally:
        vap.sdpa.unpatch_scaled_dot_product_attention()
    assert F.scaled_dot_product_attention is orig


@pytest.mark.parametrize("family", ["wan", "cog"])
def test_shell_state_dict_matches_reference(family):
    """Key-for-key and shape-for-shape equal to the reference model's state_dict (recorded by oracle/gen_golden.py)."""
    spec = json.load(open(os.path.join(ROOT, "tests", "golden", f"{family}_tiny_keys.json")))
    cfg = spec["config"]
    if family == "wan":
        cfg["patch_size"] = tuple(cfg["patch_size"])
        model = vap.WanTransformer3DMOTModel(**cfg)
    else:
        model = vap.CogVideoXTransformer3DMOTModel(**cfg)
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert mine.keys() == spec["shapes"].keys(), (set(mine) ^ set(spec["shapes"]))
    assert mine == spec["shapes"]
    # the trainer's trainable filter (finetrainers/trainer/sft_trainer/trainer.py:154-164) selects the expert by name
    assert any("_mot_ref" in k for k in mine)


def test_wan_rope_tables_match_oracle():
    for ref in (False, True):
        fr = wan_oracle.wan_rope(128, (1, 2, 2), 1024, (3, 8, 12), ref=ref)[0, 0]
        cos, sin = vap.rope.wan_rope_tables(128, (1, 2, 2), (3, 8, 12), ref=ref, device="cpu")
        assert cos.dtype == torch.float32 and cos.shape == (3 * 4 * 6, 64)
        assert torch.allclose(cos.double(), fr.real, atol=1e-7) and torch.allclose(sin.double(), fr.imag, atol=1e-7)
        c2, s2 = vap.rope.as_tables(fr.view(1, 1, -1, 64), 128, "cpu")
        assert torch.equal(c2, cos) or torch.allclose(c2, cos, atol=1e-7)


def test_cog_rope_matches_oracle_and_compacts():
    for kw in (dict(), dict(mot_num=1), dict(mot_num=2), dict(mot_num=1, ref_type="discrete_long_reference")):
        a = vap.rope.get_3d_rotary_pos_embed(64, ((0, 0), (6, 8)), (6, 8), 3, **kw)
        b = cog_oracle.cog_rope_3d(64, ((0, 0), (6, 8)), (6, 8), 3, **kw)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
        c, s = vap.rope.as_tables(a, 64, "cpu")
        assert c.shape == (a[0].shape[0], 32) and torch.equal(c, a[0][:, ::2].contiguous())
    with pytest.raises(ValueError):
        vap.rope.get_3d_rotary_pos_embed(64, ((0, 0), (6, 8)), (6, 8), 3, mot_num=1, ref_type="bogus")


def test_synth_is_deterministic_and_name_keyed():
    a = vap.synth.synth_tensor("blocks.0.attn1.to_q.weight", (8, 8), seed=1)
    assert torch.equal(a, vap.synth.synth_tensor("blocks.0.attn1.to_q.weight", (8, 8), seed=1))
    assert not torch.equal(a, vap.synth.synth_tensor("blocks.0.attn1.to_k.weight", (8, 8), seed=1))
    assert abs(vap.synth.synth_tensor("blocks.0.norm2.weight", (4096,), 0).mean().item() - 1.0) < 0.02


def test_install_rebinds_block_forward_and_keeps_state_dict():
    cfg = dict(vap.synth.WAN_TINY)
    model = vap.WanTransformer3DMOTModel(**cfg)
    keys = list(model.state_dict())
    vap.install(model, level="block")
    assert all(b.forward.__func__ is vap.wan_block_forward for b in model.blocks)
    assert list(model.state_dict()) == keys
    vap.uninstall(model)
    vap.install(model, level="processor")
    import torch.nn.functional as F
    assert F.scaled_dot_product_attention is vap.joint_sdpa
    assert type(model.blocks[0].attn1.processor).__name__ == "WanAttnMOTProcessor2_0"
    vap.uninstall(model)
    assert F.scaled_dot_product_attention is not vap.joint_sdpa
    with pytest.raises(ValueError):
        vap.install(model, level="nope")
    with pytest.raises(TypeError):
        vap.install(torch.nn.Linear(2, 2))


def test_attention_module_filters_unknown_kwargs(caplog):
    """Attention.forward drops kwargs the processor does not declare, with a warning (attention_processor.py:593-602)."""
    seen = {}

    class Proc:
        def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, rotary_emb=None):
            seen["rotary"] = rotary_emb
            return hidden_states

    a = vap.modules.Attention(16, 2, 8, "rms_norm_across_heads", processor=Proc())
    with caplog.at_level("WARNING", logger="vap_b200"):
        out = a(torch.zeros(1, 2, 16), rotary_emb="r", bogus=1)
    assert seen["rotary"] == "r" and out.shape == (1, 2, 16)
    assert any("bogus" in r.message for r in caplog.records)


def test_packed_qkv_views_keep_state_dict_roundtrip():
    a = vap.modules.Attention(16, 2, 8, "rms_norm_across_heads", bias=True)
    before = {k: v.clone() for k, v in a.state_dict().items()}
    W, b = vap.wan._packed(a, "qkv", [a.to_q, a.to_k, a.to_v])
    assert W.shape == (48, 16) and a.to_k.weight.data_ptr() == W[16:32].data_ptr()
    after = a.state_dict()
    assert all(torch.equal(before[k], after[k]) for k in before)
    a.load_state_dict({k: v + 1 for k, v in before.items()})
    assert torch.equal(W[:16], before["to_q.weight"] + 1), "load_state_dict must write through to the packed storage"
    assert vap.wan._packed(a, "qkv", [a.to_q, a.to_k, a.to_v])[0] is W


def test_flow_match_schedule_matches_oracle():
    from oracle import denoise as od
    for n, sh in ((4, 3.0), (50, 5.0), (4, 1.0)):
        t1, s1 = vap.denoise.flow_match_schedule(n, sh)
        t2, s2 = od.flow_match_schedule(n, sh)
        assert torch.allclose(t1, t2, rtol=1e-6) and torch.allclose(s1, s2, rtol=1e-6)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's own CPU path timed on the host cores) prints ONE JSON line with the contract's keys; run here on
    the tiny workload so it takes seconds."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "wan_tiny", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "steps/s" and line["value"] > 0
    # kind "reference": the reference's own module from baseline/_ref (staged here) or /root/reference; "port": the oracle, where neither exists
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    # the CPU arm times a bounded sample and says so: the step it reports is flagged as extrapolated, with the measured seconds and the factor
    assert line["extrapolated"] is True and line["measured_ms_per_sample"] > 0 and line["extrapolation_factor"] >= 1.0
    assert line["e2e"] == {"value": line["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_attention_kv_split_heuristic(monkeypatch):
    """The wave model behind the automatic split-KV attention: only shapes whose (batch, head, 256-row) work items are a poor
    multiple of the 148 SMs are split — the 5-heads-per-rank shape of 8-way Ulysses at 480p — and never a short KV sequence."""
    monkeypatch.setattr(vap.ops, "sm_count", lambda: 148)
    ks = vap.ops.attention_kv_splits
    J = 40560
    assert ks(1, 5, J, J) == 2            # Ulysses 8 ranks: 795 items = 5.37 waves -> 6; two ranges -> 5.5
    assert ks(1, 40, J, J) == 1           # single GPU: 42.97 waves
    assert ks(1, 20, J, J) == 1 and ks(1, 10, J, J) == 1  # 2 and 4 ranks
    assert ks(1, 5, 151200, 151200) == 1  # 720p at 8 ranks: 19.97 waves
    assert ks(1, 48, 35552, 35552) == 1   # CogVideoX-5B
    assert ks(1, 40, 20280, 512) == 1 and ks(1, 2, 300, 300) == 1  # cross-attention / tiny: too few KV tiles to cut
    monkeypatch.setenv("VAP_ATTN_SPLITKV", "0")
    vap.ops._SPLIT_CACHE.clear()
    assert vap.ops._auto_kv_splits(1, 5, J, J) == 1
    monkeypatch.setenv("VAP_ATTN_SPLITKV", "3")
    vap.ops._SPLIT_CACHE.clear()
    assert vap.ops._auto_kv_splits(1, 5, J, J) == 3 and vap.ops._auto_kv_splits(1, 1, 64, 200) == 2  # clamped to the KV tiles
    vap.ops._SPLIT_CACHE.clear()


def test_split_kv_entry_points_validate_arguments():
    lib = vap._lib.load()
    assert lib.vap_attention_fwd_splitkv(16, 16, 16, 16, 16, 9, 1, 1, 8, 8, 128, *([8] * 9), 1.0, 0) == -1
    assert b"kv_splits" in lib.vap_last_error()
    assert lib.vap_attention_fwd_splitkv(16, 16, 16, 16, 16, 4, 1, 1, 8, 300, 128, *([8] * 9), 1.0, 0) == -1  # 3 KV tiles < 4 ranges
    assert b"KV tiles" in lib.vap_last_error()
    assert lib.vap_attention_combine(16, 16, 2, 1, 1, 8, 128, 0, 0, 0, 0, 0, 128, 128, 128, 0) == -1  # neither o nor o_peers
    assert b"either o or o_peers" in lib.vap_last_error()


# ------------------------------------------------------------------------------------------------------------------
# launch sequence of the fused blocks: which C-ABI entry points one forward calls, in which order, on which shapes
# (tools/host_path_profile.py swaps the library for a no-op stub in a SUBPROCESS; nothing is computed)
# ------------------------------------------------------------------------------------------------------------------
def _recorded_calls(family: str, blocks: int, tmp_path):
    import subprocess
    out = tmp_path / f"{family}_calls.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "host_path_profile.py"), "--family", family, "--blocks", str(blocks),
                        "--iters", "1", "--record", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    summary = json.loads(r.stdout.strip().splitlines()[-1])
    calls = json.load(open(out))
    assert summary["c_abi_calls_per_forward"] == len(calls)
    return calls


def test_wan_mot_block_launch_sequence(tmp_path):
    """One Wan MoT block = 37 launches (SURVEY §8a W1-W9 on both streams): the two streams' adaLN modulation vectors (one launch each), per
    stream LN -> fused-QKV GEMM -> q/k norm + RoPE into the
    JOINT buffer, ONE attention over [target | ref], per stream O-projection with the gated-residual epilogue, then per stream the
    cross-attention (q / text kv / image kv projections, three norms, two attentions — the second accumulating into the first's output —,
    ungated residual epilogue) and the FFN (GELU and
    gated-residual epilogues); + the output head's LayerNorm.  No pack / cat / transpose kernels of ours in between."""
    calls = _recorded_calls("wan", 1, tmp_path)
    names = [c[0].replace("vap_", "") for c in calls]
    ln, gemm, qk, att, mod, acc = "adaln_layernorm", "gemm_bf16", "qk_norm_rope", "attention_fwd", "wan_modulation", "attention_fwd_accumulate"
    tail = [ln, gemm, qk, gemm, qk, att, gemm, qk, acc, gemm, ln, gemm, gemm]  # the image softmax is added to the text one in its epilogue
    # issue order: the expert's stream (side CUDA stream on a GPU, streams.py) first, then the target's, in each of the two phases around the attention
    assert names == [mod] + [mod, ln, gemm, qk] + [ln, gemm, qk] + [att] + [gemm] + tail + [gemm] + tail + [ln]
    # the joint attention reads q, k, v of BOTH streams as strided views of one [J, 3d] buffer: J = 2 x 32 rows, row stride 3 * 256
    joint = [c for c in calls if c[0] == "vap_attention_fwd"][0]
    scal = [x for x in joint[1:] if x not in ("p", None)]
    B, H, Lq, Lkv, D = scal[:5]
    assert (B, H, Lq, Lkv, D) == (1, 2, 64, 64, 128)
    q_strides, k_strides, v_strides, o_strides = scal[5:8], scal[8:11], scal[11:14], scal[14:17]
    assert q_strides == k_strides == v_strides == [64 * 768, 128, 768] and o_strides == [64 * 256, 128, 256]
    # GEMM epilogues in order: QKV x2 (bias), O-proj x2 (fp32 gated residual), then per stream q / kv / kv_img (bias), to_out (residual add),
    # FFN up (GELU), FFN down (fp32 gated residual)
    epi = [[x for x in c[1:] if x not in ("p", None)][6] for c in calls if c[0] == "vap_gemm_bf16"]
    assert epi == [0, 0] + [2, 0, 0, 0, 3, 1, 2] * 2


def test_cog_mot_block_launch_sequence(tmp_path):
    """CogVideoX: block 0 carries the expert (18 launches), block 1 is a plain block (9); each stream's LayerNormZero writes text and
    video rows into ONE [T + S, d] buffer feeding the QKV / FFN GEMMs, the joint attention runs over [text | video | text_ref | video_ref]."""
    calls = _recorded_calls("cog", 2, tmp_path)
    names = [c[0].replace("vap_", "") for c in calls]
    ln, gemm, qk, att = "adaln_layernorm", "gemm_bf16", "qk_norm_rope", "attention_fwd"
    pre = [ln, ln, gemm, qk]
    ffn = [ln, ln, gemm, gemm, gemm]
    assert names == pre + pre + [att, gemm, gemm] + ffn + [gemm, gemm] + ffn + pre + [att, gemm, gemm] + ffn + [ln, ln]  # per phase: expert stream, then target
    att_calls = [[x for x in c[1:] if x not in ("p", None)] for c in calls if c[0] == "vap_attention_fwd"]
    assert att_calls[0][:5] == [1, 4, 516, 516, 64] and att_calls[1][:5] == [1, 4, 258, 258, 64]  # J = 2 (226 + 32); plain block: 226 + 32
    qk_calls = [[x for x in c[1:] if x not in ("p", None)] for c in calls if c[0] == "vap_qk_norm_rope"]
    assert all(c[:4] == [258, 4, 64, 768] and c[5] == 226 for c in qk_calls)  # RoPE skips the 226 text rows (rope_row0)


def test_cfg_flow_match_step_contract_equals_the_reference_expression():
    """The rounding points vap_cfg_flow_match_step implements (restated by the CPU stand-in) are those of the reference's tensor ops:
    noise_uncond + g * (noise - noise_uncond) in bf16 (pipeline_wan_i2v_mot.py:874), then FlowMatchEulerDiscreteScheduler.step
    (scheduling_flow_match_euler_discrete.py:433-467).  torch casts the scheduler's 0-dim fp32 dt to the tensor operand's dtype (bf16)
    before the multiply, so the kernel's caller hands over the bf16-rounded dt — as this comparison does."""
    import cpu_standin_ops as so
    g = torch.Generator().manual_seed(5)
    c, u = torch.randn(2, 4096, generator=g).bfloat16(), torch.randn(2, 4096, generator=g).bfloat16()
    sigma, sigma_next = torch.tensor(0.9375), torch.tensor(0.712345)
    dt_b = float((sigma_next - sigma).bfloat16())
    for sample in (torch.randn(2, 4096, generator=g), torch.randn(2, 4096, generator=g).bfloat16()):
        noise = u + 5.0 * (c - u)
        ref = (sample.to(torch.float32) + (sigma_next - sigma) * noise).to(noise.dtype)
        assert torch.equal(so.cfg_flow_match_step(c, u, sample, guidance_scale=5.0, dt=dt_b), ref)
        ref1 = (sample.to(torch.float32) + (sigma_next - sigma) * c).to(c.dtype)
        assert torch.equal(so.cfg_flow_match_step(c, None, sample, guidance_scale=1.0, dt=dt_b), ref1)
    with pytest.raises(vap.VapError):
        vap.ops.cfg_flow_match_step(c, u, torch.zeros(2, 4096), guidance_scale=5.0, dt=-0.1)
    lib = vap._lib.load()
    assert lib.vap_cfg_flow_match_step(16, 0, 16, 1, 16, 1, 12, 16, 5.0, -0.1, 0) == -1 and b"multiple of 8" in lib.vap_last_error()


def test_bench_flop_accounting_matches_the_survey():
    """bench.py's algorithmic-FLOP model (roofline numerators, model_tflops) reproduces SURVEY.md §8's figures for the three expert placements
    of Wan2.1-14B at 49f 480x832 (2.349e15 / 1.594e15 / 1.216e15 per forward, 3.369e13 per joint attention) and for 81f 720p (2.244e16, 4.682e14)."""
    sys.path.insert(0, ROOT)
    import bench
    w = bench.workloads(vap.synth)
    want = {"wan14b": (2.349e15, 3.369e13), "wan14b_d20": (1.594e15, 3.369e13), "wan14b_d10": (1.216e15, 3.369e13), "wan14b_720p": (2.244e16, 4.682e14)}
    for name, (total, attn) in want.items():
        f, h, wd = w[name]["latent"]
        S = f * (h // 2) * (wd // 2)
        got_total, got_attn = bench.wan_flops(w[name]["cfg"], S, S)
        assert abs(got_total / total - 1) < 1e-3 and abs(got_attn / attn - 1) < 1e-3, (name, got_total, got_attn)
    assert [i for i in w["wan14b_d20"]["cfg"]["block_idx_with_mot_ref"]] == list(range(0, 40, 2))


def test_dual_stream_schedule_is_off_for_cpu_tensors_and_automatic_by_row_count(monkeypatch):
    """streams.dual(): None for CPU tensors (the gloo / stand-in tests run both token streams in issue order on one queue); in the default
    "auto" mode the two-stream schedule is chosen by the rows per stream (measured: +6.6 % at the 2 535 rows a rank owns under 8-way Ulysses,
    neutral at the full 20 280); dual_streams(True / False / None) forces / restores it."""
    st = vap.streams
    assert st.dual(torch.device("cpu"), 100) is None
    created = []
    monkeypatch.setattr(st, "DualStream", lambda dev: created.append(dev) or object())
    monkeypatch.setattr(st, "_PER_DEVICE", {})
    with vap.dual_streams(None):
        assert st.dual(torch.device("cuda", 0), 2535) is not None and st.dual(torch.device("cuda", 0), 20280) is None
        assert st.dual(torch.device("cuda", 0), st.AUTO_MAX_ROWS) is not None
    with vap.dual_streams(True):
        assert st.dual(torch.device("cuda", 0), 20280) is not None
    with vap.dual_streams(False):
        assert st.dual(torch.device("cuda", 0), 100) is None
    assert created == [torch.device("cuda", 0)]  # one DualStream per device, cached


def test_rows_view_accepts_an_empty_batched_tensor():
    """A sequence-parallel CogVideoX rank may hold text rows only: its video tensor is [B, 0, d] (B = 2 under CFG), whose batch stride cannot
    'continue the row pitch' — there is nothing to address, the wrappers must see 0 rows instead of raising."""
    t = torch.empty((2, 280, 512), dtype=torch.bfloat16)[:, 140:140]
    assert vap.ops._rows_view(t, "x")[0] == 0
    with pytest.raises(ValueError):  # a non-empty slice whose batch stride breaks the row pitch is still refused
        vap.ops._rows_view(torch.empty((2, 3, 512))[:, 1:3], "x")
