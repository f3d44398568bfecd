"""Trainer seam at block level: the fused forward behind activation checkpointing.

The reference's trainer runs every block through `_gradient_checkpointing_func` when gradient checkpointing is on
(transformer_wan_mot.py:923-935, cogvideox_transformer_3d_mot.py:1016-1029; finetrainers/trainer/sft_trainer/trainer.py:232-282): the forward
pass keeps only the block's INPUTS and the backward pass re-runs the block under autograd.  `install(model, level="block", trainable=True)`
gives that scheme a faster first pass: a block's forward is ONE autograd node whose

  forward   runs the fused sm_100a path (wan_block_forward / cog_block_forward: no autograd graph, no saved activations), and whose
  backward  re-runs the block's ORIGINAL forward — the reference's own code, untouched — on the saved inputs under autograd, with
            `F.scaled_dot_product_attention` routed to `joint_sdpa` (vap_attention_fwd + vap_attention_bwd), and back-propagates the incoming
            gradients through it; parameter gradients accumulate into `.grad` as with `torch.utils.checkpoint(use_reentrant=True)`.

Gradients are therefore exactly those of the reference block evaluated at the saved inputs (the kernels' forward differs from the recomputed
one by bf16 rounding only: per-block parity gate 2e-2).  Without gradient tracking (inference, `torch.no_grad()`) the wrapper is the fused
forward itself.  LoRA / FSDP-sharded layers are refused by the fused path (wan._plain_linear) — use `level="sdpa"` there.
"""
from __future__ import annotations

import contextlib
from typing import Callable

import torch
import torch.nn.functional as F
from torch.utils import _pytree as pytree

from . import sdpa


@contextlib.contextmanager
def _joint_sdpa_slot():
    """F.scaled_dot_product_attention -> joint_sdpa for the duration of a recompute (calls outside its envelope go to torch's own)."""
    if F.scaled_dot_product_attention is sdpa.joint_sdpa:
        yield
        return
    prev_fn, prev_orig, prev_strict = F.scaled_dot_product_attention, sdpa._ORIGINAL_SDPA, sdpa._STRICT
    sdpa._ORIGINAL_SDPA, sdpa._STRICT = prev_fn, False
    F.scaled_dot_product_attention = sdpa.joint_sdpa
    try:
        yield
    finally:
        F.scaled_dot_product_attention = prev_fn
        sdpa._ORIGINAL_SDPA, sdpa._STRICT = prev_orig, prev_strict


class _FusedForwardRecomputeBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, run_fused: Callable, run_original: Callable, _anchor: torch.Tensor, *tensors: torch.Tensor):
        ctx.run_original = run_original
        ctx.set_materialize_grads(False)  # an output nobody reads (the expert stream after the last MoT block) gets None, not zeros: its branch is not back-propagated
        ctx.save_for_backward(*tensors)
        out = run_fused([t.detach() for t in tensors])  # autograd.Function.forward runs with gradient tracking off
        return tuple(o for o in pytree.tree_flatten(out)[0] if torch.is_tensor(o))

    @staticmethod
    def backward(ctx, *grads):
        inputs = [t.detach().requires_grad_(need) for t, need in zip(ctx.saved_tensors, ctx.needs_input_grad[3:])]
        with torch.enable_grad(), _joint_sdpa_slot():
            out = ctx.run_original(inputs)
        outs = [o for o in pytree.tree_flatten(out)[0] if torch.is_tensor(o)]
        pairs = [(o, g) for o, g in zip(outs, grads) if g is not None and o.requires_grad]
        if pairs:
            torch.autograd.backward([o for o, _ in pairs], [g for _, g in pairs])
        return (None, None, None) + tuple(i.grad if i.requires_grad else None for i in inputs)


def checkpointed_block_forward(fused_forward: Callable, original_forward: Callable) -> Callable:
    """-> forward(self, *args, **kwargs) for a block: fused first pass, reference recompute in the backward (module docstring).
    `original_forward` is the block's bound original forward."""

    def forward(self, *args, **kwargs):
        leaves, spec = pytree.tree_flatten((args, kwargs))
        track = torch.is_grad_enabled() and (any(torch.is_tensor(l) and l.requires_grad for l in leaves) or any(p.requires_grad for p in self.parameters()))
        if not track:
            return fused_forward(self, *args, **kwargs)
        idx = [i for i, l in enumerate(leaves) if torch.is_tensor(l) and l.requires_grad]

        def rebuild(tensors):
            cur = list(leaves)
            for i, t in zip(idx, tensors):
                cur[i] = t
            return pytree.tree_unflatten(cur, spec)

        box = {}

        def run_fused(tensors):
            a, kw = rebuild(tensors)
            out = fused_forward(self, *a, **kw)
            box["leaves"], box["spec"] = pytree.tree_flatten(out)
            return out

        def run_original(tensors):
            a, kw = rebuild(tensors)
            return original_forward(*a, **kw)

        # the anchor makes the node part of the graph even when no INPUT requires grad but the block's parameters do (first block of a frozen trunk)
        anchor = torch.empty(0, device=leaves[idx[0]].device if idx else None, requires_grad=True)
        flat = _FusedForwardRecomputeBackward.apply(run_fused, run_original, anchor, *[leaves[i] for i in idx])
        # same structure as the block's own return value (Wan: 2-tuple, CogVideoX: 4-tuple), tensors replaced by the node's outputs
        it = iter(flat)
        return pytree.tree_unflatten([next(it) if torch.is_tensor(l) else l for l in box["leaves"]], box["spec"])

    return forward
