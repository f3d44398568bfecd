"""Deterministic synthetic weights, configurations and inputs (there is no network for checkpoints or datasets).

Weights are a pure function of (parameter name, shape, seed): every tensor is drawn from its own CPU generator seeded
with crc32(name) ^ seed, so the reference model, the oracle and the B200 path can be given bit-identical weights from
the parameter names alone (tests/golden fixtures only store inputs and expected outputs).  On a CUDA device (benchmarks
at 14B scale) the same rule is applied with a device generator — fast, but a different stream of numbers.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable, Optional, Tuple

import torch

# ---------------------------------------------------------------------------------------------------------------
# configurations (BASELINE.json `configs`; shapes pinned by the reference's conversion scripts)
# ---------------------------------------------------------------------------------------------------------------
WAN_TINY = dict(patch_size=(1, 2, 2), num_attention_heads=2, attention_head_dim=128, in_channels=36, out_channels=16, text_dim=64,
                freq_dim=256, ffn_dim=512, num_layers=2, cross_attn_norm=True, qk_norm="rms_norm_across_heads", eps=1e-6, image_dim=32,
                added_kv_proj_dim=256, rope_max_seq_len=1024, block_idx_with_mot_ref=[0])
# Wan2.1-I2V-14B (diffusers/scripts/convert_wan_to_diffusers.py:116-135) + VAP expert in every block (config_ori.json)
WAN_14B = dict(patch_size=(1, 2, 2), num_attention_heads=40, attention_head_dim=128, in_channels=36, out_channels=16, text_dim=4096,
               freq_dim=256, ffn_dim=13824, num_layers=40, cross_attn_norm=True, qk_norm="rms_norm_across_heads", eps=1e-6, image_dim=1280,
               added_kv_proj_dim=5120, rope_max_seq_len=1024, block_idx_with_mot_ref=list(range(40)))
COG_TINY = dict(num_attention_heads=4, attention_head_dim=64, in_channels=32, out_channels=16, time_embed_dim=64, text_embed_dim=128,
                num_layers=2, sample_width=90, sample_height=60, sample_frames=49, patch_size=2, max_text_seq_length=226,
                use_rotary_positional_embeddings=True, use_learned_positional_embeddings=False, norm_eps=1e-5, block_idx_with_mot_ref=[0])
# CogVideoX-5B-I2V (diffusers/scripts/convert_cogvideox_to_diffusers.py:148-157, 202-212, 249-252) + VAP expert in blocks 0..40
COG_5B = dict(num_attention_heads=48, attention_head_dim=64, in_channels=32, out_channels=16, time_embed_dim=512, text_embed_dim=4096,
              num_layers=42, sample_width=90, sample_height=60, sample_frames=49, patch_size=2, max_text_seq_length=226,
              use_rotary_positional_embeddings=True, use_learned_positional_embeddings=True, norm_eps=1e-5, block_idx_with_mot_ref=list(range(41)))


def _seed_of(name: str, seed: int) -> int:
    return (zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF


def synth_tensor(name: str, shape: Tuple[int, ...], seed: int = 0, device="cpu", num_layers: int = 1) -> torch.Tensor:
    """fp32 value of parameter `name`: weights ~ N(0, 1/fan_in) (residual-branch outputs further scaled by 1/sqrt(2L) so
    40 random blocks keep O(1) activations), biases ~ N(0, 0.02^2), norm weights 1 + N(0, 0.1^2), norm biases N(0, 0.05^2),
    scale_shift_table ~ N(0, 1/d) (the reference's own init, transformer_wan_mot.py:523)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(_seed_of(name, seed))
    r = torch.randn(shape, generator=g, dtype=torch.float32, device=dev)
    leaf = name.split(".")[-1]
    if "scale_shift_table" in name:
        return r / math.sqrt(shape[-1])
    if leaf == "pos_embedding":
        return r * 0.02
    is_norm = any(part.startswith("norm") for part in name.split(".")[:-1]) and len(shape) == 1 and ".linear." not in name
    if is_norm:
        return 1.0 + 0.1 * r if leaf == "weight" else 0.05 * r
    if leaf == "bias":
        return 0.02 * r
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    w = r / math.sqrt(max(fan_in, 1))
    if ".to_out." in name or ".net.2." in name:  # residual-branch output projections
        w = w / math.sqrt(2.0 * max(num_layers, 1))
    return w


def synth_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0, dtype=torch.bfloat16, device="cpu", num_layers: int = 1) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, tuple(s), seed, device, num_layers).to(dtype) for k, s in shapes.items()}


@torch.no_grad()
def fill_module_(module: torch.nn.Module, seed: int = 0, num_layers: int = 1) -> torch.nn.Module:
    """Overwrite every parameter and persistent buffer of `module` (ours or the reference's) with its synthetic value."""
    for name, t in list(module.named_parameters()) + [(n, b) for n, b in module.named_buffers() if n in module.state_dict()]:
        t.copy_(synth_tensor(name, tuple(t.shape), seed, t.device, num_layers).to(t.dtype))
    return module


# ---------------------------------------------------------------------------------------------------------------
# inputs
# ---------------------------------------------------------------------------------------------------------------
def _randn(name: str, shape, seed: int, device) -> torch.Tensor:
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(_seed_of("input:" + name, seed))
    return torch.randn(shape, generator=g, dtype=torch.float32, device=dev)


def wan_inputs(cfg: dict, frames: int, height: int, width: int, seed: int = 0, device="cpu", dtype=torch.bfloat16, timestep: float = 500.0,
               batch: int = 1) -> dict:
    """Synthetic forward kwargs of WanTransformer3DMOTModel: latents [B,36,F,h,w] whose channels 16-19 are the I2V mask
    (1 on frame 0, pipeline_wan_i2v_mot.py:437-447), UMT5-like text [B,512,text_dim] with a zero-padded tail (:210-214),
    CLIP-like image tokens [B,257,image_dim], reference stream at the fixed timestep 1 (:812-813)."""
    C = cfg["in_channels"]
    def latent(tag):
        x = _randn(tag, (batch, C, frames, height, width), seed, device)
        if C >= 20:
            x[:, 16:20] = 0
            x[:, 16:20, 0] = 1
        return x.to(dtype)
    def text(tag):
        t = _randn(tag, (batch, 512, cfg["text_dim"]), seed, device) * 0.1
        t[:, 384:] = 0
        return t.to(dtype)
    return dict(
        hidden_states=latent("wan.latent"), hidden_states_mot_ref=latent("wan.latent_ref"),
        timestep=torch.full((batch,), timestep, dtype=torch.float32, device=device),
        timestep_list_mot_ref=torch.ones((1, batch), dtype=torch.float32, device=device),
        encoder_hidden_states=text("wan.text"), encoder_hidden_states_mot_ref=text("wan.text_ref"),
        encoder_hidden_states_image=_randn("wan.clip", (batch, 257, cfg["image_dim"]), seed, device).to(dtype),
        encoder_hidden_states_image_mot_ref=_randn("wan.clip_ref", (batch, 257, cfg["image_dim"]), seed, device).to(dtype),
        num_mot_ref=1,
    )


def cog_inputs(cfg: dict, frames: int, height: int, width: int, seed: int = 0, device="cpu", dtype=torch.bfloat16, timestep: float = 500.0,
               batch: int = 1, num_mot_ref: int = 1, rope_fn=None) -> dict:
    """Synthetic forward kwargs of CogVideoXTransformer3DMOTModel: latents [B,F,32,h,w], T5-like text [B,226,text_dim],
    RoPE tables from get_3d_rotary_pos_embed for the target and (mot_num, continous_negative) for the reference stream."""
    if rope_fn is None:
        from .rope import get_3d_rotary_pos_embed as rope_fn
    C, p, D = cfg["in_channels"], cfg["patch_size"], cfg["attention_head_dim"]
    gh, gw = height // p, width // p
    T = cfg["max_text_seq_length"]
    return dict(
        hidden_states=_randn("cog.latent", (batch, frames, C, height, width), seed, device).to(dtype),
        hidden_states_mot_ref=_randn("cog.latent_ref", (batch, frames * num_mot_ref, C, height, width), seed, device).to(dtype),
        encoder_hidden_states=(_randn("cog.text", (batch, T, cfg["text_embed_dim"]), seed, device) * 0.1).to(dtype),
        encoder_hidden_states_mot_ref=(_randn("cog.text_ref", (batch, T * num_mot_ref, cfg["text_embed_dim"]), seed, device) * 0.1).to(dtype),
        timestep=torch.full((batch,), timestep, dtype=torch.float32, device=device),
        image_rotary_emb=rope_fn(D, ((0, 0), (gh, gw)), (gh, gw), frames, device=device),
        image_rotary_emb_mot_ref=rope_fn(D, ((0, 0), (gh, gw)), (gh, gw), frames, device=device, mot_num=num_mot_ref,
                                         ref_type="continous_negative"),
        num_mot_ref=num_mot_ref,
    )
