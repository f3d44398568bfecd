"""video-as-prompt_b200 — B200-native (sm_100a) implementation of Video-As-Prompt's MoT denoise hot path.

Import as ``import vap_b200`` (alias module at the repo root) or ``importlib.import_module("video-as-prompt_b200")``.
The CUDA kernels live in ``libvap_b200.so`` (C ABI: ``include/vap_b200.h``), built in-tree by ``csrc/build.py``.
"""
from . import _lib, cogvideox, denoise, graphs, install as _install_mod, modules, ops, rope, sdpa, streams, synth, ulysses, wan  # noqa: F401
from .graphs import GraphedForward  # noqa: F401
from ._lib import VapError  # noqa: F401
from .cogvideox import CogVideoXAttnMOTProcessor2_0, CogVideoXAttnProcessor2_0, CogVideoXTransformer3DMOTModel, cog_block_forward  # noqa: F401
from .install import install, uninstall  # noqa: F401
from .sdpa import joint_sdpa  # noqa: F401
from .streams import dual_streams  # noqa: F401
from .wan import (WanAttnCrossMOTProcessor2_0, WanAttnMOTProcessor2_0, WanAttnProcessor2_0, WanTransformer3DMOTModel,  # noqa: F401
                  wan_block_forward)

__version__ = "0.1.0"
