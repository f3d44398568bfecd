"""`install(model)`: make a reference (or stand-alone) MoT transformer run its blocks on the sm_100a kernels.

The module tree, parameter names and `state_dict()` keys are left untouched — `from_pretrained`, LoRA target regexes,
FSDP wrapping and the trainer's `"_mot_ref" in name` trainable filter (finetrainers/trainer/sft_trainer/trainer.py:154-164)
keep working.  Three seams, from coarse to fine:

  level="block"      rebind `forward` on every WanTransformerBlock / CogVideoXBlock instance to the fused forward
                     (B3; precedent for rebinding: finetrainers/patches/models/wan/patch.py:22-25)
  level="processor"  swap the attention processors for the same-named ones of this package (B1), keep the reference's
                     block arithmetic, and route the block's F.scaled_dot_product_attention call to `joint_sdpa` (B2)
  level="sdpa"       only replace F.scaled_dot_product_attention (B2)
"""
from __future__ import annotations

import types

import torch.nn as nn

from . import cogvideox, rope, sdpa, training, wan

_WAN_PROC = {"WanAttnMOTProcessor2_0": wan.WanAttnMOTProcessor2_0, "WanAttnCrossMOTProcessor2_0": wan.WanAttnCrossMOTProcessor2_0,
             "WanAttnProcessor2_0": wan.WanAttnProcessor2_0}
_COG_PROC = {"CogVideoXAttnMOTProcessor2_0": cogvideox.CogVideoXAttnMOTProcessor2_0, "CogVideoXAttnProcessor2_0": cogvideox.CogVideoXAttnProcessor2_0}


def _blocks(model: nn.Module):
    if hasattr(model, "blocks"):
        return "wan", list(model.blocks)
    if hasattr(model, "transformer_blocks"):
        return "cog", list(model.transformer_blocks)
    raise TypeError(f"{type(model).__name__} has neither `.blocks` (Wan) nor `.transformer_blocks` (CogVideoX)")


def _swap_wan_rope(model: nn.Module) -> None:
    """The reference's Wan shell rebuilds the RoPE tables on the CPU in float64 and copies them to the device at EVERY forward
    (WanRotaryPosEmbed.forward / WanRotaryPosEmbedRef.forward, transformer_wan_mot.py:390-409, 429-464: 2 x S x 64 complex128 = 41.6 MB
    of H2D traffic per forward at 480p).  Our blocks / processors also take compact device tables, so the two rope modules' forwards are
    rebound to the cached device builder (rope.wan_rope_tables: same positions — target t = 0..F-1, reference t = -F..-1 —, float64 angles)."""
    for name, is_ref in (("rope", False), ("rope_mot_ref", True)):
        mod = getattr(model, name, None)
        if mod is None or not all(hasattr(mod, a) for a in ("attention_head_dim", "patch_size", "max_seq_len")):
            continue
        mod.__dict__["_vap_original_forward"] = mod.forward

        def tables(hidden_states, _m=mod, _ref=is_ref):
            return rope.wan_rope_tables(_m.attention_head_dim, tuple(_m.patch_size), tuple(hidden_states.shape[2:]), ref=_ref,
                                        device=hidden_states.device, max_seq_len=_m.max_seq_len)
        mod.forward = tables


def _restore_wan_rope(model: nn.Module) -> None:
    for name in ("rope", "rope_mot_ref"):
        mod = getattr(model, name, None)
        if mod is not None and mod.__dict__.pop("_vap_original_forward", None) is not None:
            del mod.forward


def install(model: nn.Module, level: str = "block", strict: bool = True, trainable: bool = False) -> nn.Module:
    """strict=False (levels "processor" / "sdpa"): SDPA calls outside the kernel's envelope — a VAE's or text encoder's attention under the
    global patch — fall through to the original torch function instead of raising (sdpa.patch_scaled_dot_product_attention).
    trainable=True (level "block", the reference's own block classes): the fused forward becomes the first pass of activation checkpointing —
    the backward pass re-runs the block's original forward under autograd with the attention kernels' backward in the SDPA slot (training.py)."""
    family, blocks = _blocks(model)
    if trainable and level != "block":
        raise ValueError("trainable=True applies to level='block' (levels 'processor' / 'sdpa' keep torch autograd for everything but the attention)")
    if level == "block":
        fwd = wan.wan_block_forward if family == "wan" else cogvideox.cog_block_forward
        for blk in blocks:
            if not hasattr(blk, "with_mot_ref"):
                raise TypeError(f"{type(blk).__name__} is not a MoT block (no `with_mot_ref`)")
            if trainable and getattr(type(blk).forward, "__wrapped__", type(blk).forward) in (wan.wan_block_forward, cogvideox.cog_block_forward):
                raise TypeError("trainable=True needs a block whose own forward is differentiable torch code (the reference's classes); "
                                "this package's stand-alone shell has none — train it with install(level='sdpa') on the reference model")
            blk.__dict__["_vap_original_forward"] = blk.forward
            blk.forward = types.MethodType(training.checkpointed_block_forward(fwd, blk.forward) if trainable else fwd, blk)
        if family == "wan" and not trainable:  # the recomputed reference forward needs the shell's own complex freqs
            _swap_wan_rope(model)
    elif level == "processor":
        table = _WAN_PROC if family == "wan" else _COG_PROC
        for blk in blocks:
            for m in blk.modules():
                proc = getattr(m, "processor", None)
                if proc is not None and type(proc).__name__ in table and hasattr(m, "set_processor"):
                    m.__dict__.setdefault("_vap_original_processor", proc)
                    m.set_processor(table[type(proc).__name__]())
        if family == "wan":
            _swap_wan_rope(model)
        sdpa.patch_scaled_dot_product_attention(strict)
    elif level == "sdpa":
        sdpa.patch_scaled_dot_product_attention(strict)
    else:
        raise ValueError(f"unknown level {level!r}; expected 'block', 'processor' or 'sdpa'")
    return model


def uninstall(model: nn.Module) -> nn.Module:
    _, blocks = _blocks(model)
    for blk in blocks:
        orig = blk.__dict__.pop("_vap_original_forward", None)
        if orig is not None:
            del blk.forward
        for m in blk.modules():
            proc = m.__dict__.pop("_vap_original_processor", None)
            if proc is not None:
                m.set_processor(proc)
    _restore_wan_rope(model)
    sdpa.unpatch_scaled_dot_product_attention()
    return model
