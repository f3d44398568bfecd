"""CPU oracle for the Video-As-Prompt MoT denoise hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker / reported CPU
baseline, never as the thing that is shipped or measured as the GPU path.

It is a from-scratch functional restatement (plain PyTorch on CPU, no nn.Module tree)
of the reference's MoT block and transformer-shell arithmetic, keyed by the reference's
own ``state_dict`` names, with every bf16 rounding point of the reference kept.  Each
function cites the reference file:line it follows (paths relative to
``/root/reference/diffusers/src/diffusers``).

Parity pinning: the reference ships NO test, golden vector or known-answer value for the
MoT path, so the oracle is pinned against the reference module itself,
imported in the authoring container: ``oracle/gen_golden.py`` builds the reference's tiny
``WanTransformer3DMOTModel`` / ``CogVideoXTransformer3DMOTModel``, loads deterministic
synthetic weights, records per-block activations + final outputs into
``tests/golden/*.pt`` and asserts the oracle reproduces them (bit-exact on CPU, same
torch build).  ``tests/test_oracle_golden.py`` re-checks the oracle against those
committed fixtures everywhere (no ``/root/reference`` needed).
"""
