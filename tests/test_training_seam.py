"""CPU: the block-level trainer seam (video-as-prompt_b200/training.py) on a toy two-stream block, independent of the reference tree —
`checkpointed_block_forward(fused, original)` must (a) return what the fused forward returns, in its structure, (b) give the gradients of the
ORIGINAL forward evaluated at the saved inputs — for inputs and parameters, with keyword or positional arguments (the reference calls its
blocks positionally through `_gradient_checkpointing_func`), also when no input requires grad but a parameter does —, (c) not back-propagate an
output nobody reads, (d) be the plain fused forward when gradients are off, (e) leave F.scaled_dot_product_attention as it found it."""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vap = importlib.import_module("video-as-prompt_b200")
training = vap.training


class ToyBlock(nn.Module):
    """Two token streams that meet in one attention, like a MoT block; `tables` is a non-tensor-requiring-grad argument (a tuple)."""

    def __init__(self):
        super().__init__()
        self.lin = nn.Linear(8, 8)
        self.lin_ref = nn.Linear(8, 8)
        self.calls = {"original": 0, "fused": 0}

    def forward(self, x, ctx, scale, x_ref=None, tables=None):
        self.calls["original"] += 1
        return self._math(x, ctx, scale, x_ref, tables)

    def _math(self, x, ctx, scale, x_ref, tables):
        h, hr = self.lin(x) * scale, self.lin_ref(x_ref) + tables[0]
        j = torch.cat([h, hr], dim=1)
        o = F.scaled_dot_product_attention(j[:, None], j[:, None], j[:, None])[:, 0]
        return x + o[:, :x.shape[1]] + ctx.mean(), x_ref + o[:, x.shape[1]:]


def fused(self, x, ctx, scale, x_ref=None, tables=None):
    assert not torch.is_grad_enabled()  # the fused kernels run without a graph
    self.calls["fused"] += 1
    return self._math(x, ctx, scale, x_ref, tables)


def _inputs(requires_grad=True):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 8, generator=g, requires_grad=requires_grad)
    xr = torch.randn(2, 3, 8, generator=g, requires_grad=requires_grad)
    ctx = torch.randn(2, 4, generator=g)
    return x, xr, ctx, (torch.randn(8, generator=g),)


def _grads(blk, call, use_second_output=True):
    blk.zero_grad(set_to_none=True)
    x, xr, ctx, tables = _inputs()
    out = call(blk, x, xr, ctx, tables)
    loss = out[0].square().sum() + (out[1].sum() if use_second_output else 0)
    loss.backward()
    return out, x.grad, xr.grad, {n: (None if p.grad is None else p.grad.clone()) for n, p in blk.named_parameters()}


def test_fused_forward_with_recomputed_backward_gives_the_original_gradients():
    torch.manual_seed(0)
    blk = ToyBlock()
    ref_out, gx, gxr, gp = _grads(blk, lambda b, x, xr, ctx, t: b(x, ctx, 0.5, x_ref=xr, tables=t))
    sdpa_before = F.scaled_dot_product_attention
    blk.forward = types.MethodType(training.checkpointed_block_forward(fused, blk.forward), blk)
    for call in (lambda b, x, xr, ctx, t: b(x, ctx, 0.5, x_ref=xr, tables=t), lambda b, x, xr, ctx, t: b(x, ctx, 0.5, xr, t)):  # keywords / positional
        blk.calls.update(original=0, fused=0)
        out, hx, hxr, hp = _grads(blk, call)
        assert isinstance(out, tuple) and len(out) == 2 and torch.equal(out[0], ref_out[0]) and torch.equal(out[1], ref_out[1])
        assert blk.calls == {"original": 1, "fused": 1}  # one fused pass forward, one recompute in the backward
        assert torch.allclose(hx, gx) and torch.allclose(hxr, gxr)
        assert all(torch.allclose(hp[n], gp[n]) for n in gp)
    assert F.scaled_dot_product_attention is sdpa_before


def test_unused_output_is_not_back_propagated_and_no_grad_is_the_plain_fused_forward():
    torch.manual_seed(0)
    blk = ToyBlock()
    blk.forward = types.MethodType(training.checkpointed_block_forward(fused, blk.forward), blk)
    # the second output (the expert stream after the last MoT block) is dead: gradients still reach lin_ref through the attention, and the
    # incoming gradient of the dead output is None, not a tensor of zeros
    _, gx, gxr, gp = _grads(blk, lambda b, x, xr, ctx, t: b(x, ctx, 0.5, x_ref=xr, tables=t), use_second_output=False)
    assert gx is not None and gxr is not None and gp["lin_ref.weight"] is not None
    blk.calls.update(original=0, fused=0)
    with torch.no_grad():
        x, xr, ctx, tables = _inputs(requires_grad=False)
        blk(x, ctx, 0.5, x_ref=xr, tables=tables)
    assert blk.calls == {"original": 0, "fused": 1}


def test_parameters_train_even_when_no_input_requires_grad():
    torch.manual_seed(0)
    blk = ToyBlock()
    blk.forward = types.MethodType(training.checkpointed_block_forward(fused, blk.forward), blk)
    x, xr, ctx, tables = _inputs(requires_grad=False)  # first block behind a frozen trunk: only the block's own parameters need gradients
    out = blk(x, ctx, 0.5, x_ref=xr, tables=tables)
    assert out[0].requires_grad
    out[0].sum().backward()
    assert blk.lin.weight.grad is not None and blk.lin_ref.weight.grad is not None


def test_install_trainable_refuses_levels_and_blocks_without_a_differentiable_forward():
    import pytest
    cfg = dict(vap.synth.WAN_TINY, num_layers=1, block_idx_with_mot_ref=[0])
    model = vap.WanTransformer3DMOTModel(**cfg)
    with pytest.raises(ValueError):
        vap.install(model, level="sdpa", trainable=True)
    with pytest.raises(TypeError):  # this package's own shell has no torch-autograd block forward to recompute
        vap.install(model, level="block", trainable=True)
