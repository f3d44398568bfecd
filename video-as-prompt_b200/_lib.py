"""ctypes binding of libvap_b200.so (C ABI declared in include/vap_b200.h).

The library is built in-tree by ``csrc/build.py`` (also run by ``__graft_entry__.build()``).  There is NO CPU
fallback: if the shared library is missing or a call fails, a ``VapError`` is raised with the library's message.
"""
from __future__ import annotations

import ctypes
import os as _os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# VAP_B200_LIB: developer override used by tools/attn_variants.sh to A/B kernel builds; the product path is the in-tree library
LIB_PATH = Path(_os.environ["VAP_B200_LIB"]) if _os.environ.get("VAP_B200_LIB") else _PKG / "libvap_b200.so"


class VapError(RuntimeError):
    """Raised when libvap_b200.so is missing or one of its entry points reports an error."""


_FLOAT_P = c_void_p  # fp32 device pointers are passed as raw addresses

# name -> (restype, argtypes); mirrors include/vap_b200.h one to one (tests check every symbol is exported)
SIGNATURES = {
    "vap_version": (c_int, []),
    "vap_last_error": (c_char_p, []),
    "vap_sm_count": (c_int, []),
    "vap_adaln_layernorm": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int64, _FLOAT_P, _FLOAT_P, _FLOAT_P, _FLOAT_P,
                                    c_int64, c_int64, c_float, c_int, c_void_p]),
    "vap_qk_norm_rope": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int64, _FLOAT_P, _FLOAT_P, _FLOAT_P, _FLOAT_P, _FLOAT_P,
                                 _FLOAT_P, c_int64, c_int64, c_int64, c_float, c_int, c_void_p]),
    "vap_attention_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, _FLOAT_P, c_int, c_int, c_int, c_int, c_int] + [c_int64] * 12
                          + [c_float, c_void_p]),
    "vap_attention_fwd_accumulate": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, _FLOAT_P, c_int, c_int, c_int, c_int, c_int] + [c_int64] * 12
                                     + [c_float, c_void_p]),
    "vap_qkv_scatter": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int64, _FLOAT_P, _FLOAT_P, _FLOAT_P, _FLOAT_P, _FLOAT_P,
                                _FLOAT_P, c_int64, c_int64, c_int64, c_float, c_int, c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p]),
    "vap_attention_fwd_scatter": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, _FLOAT_P, c_int, c_int, c_int, c_int, c_int]
                                  + [c_int64] * 12 + [c_float, c_void_p]),
    "vap_attention_fwd_splitkv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, _FLOAT_P, c_int, c_int, c_int, c_int, c_int, c_int] + [c_int64] * 9
                                  + [c_float, c_void_p]),
    "vap_attention_combine": (c_int, [c_void_p, _FLOAT_P, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, _FLOAT_P, c_int64, c_int64,
                                      c_int64, c_void_p]),
    "vap_gemm_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                              c_int64, _FLOAT_P, c_int64, c_int64, c_void_p]),
    "vap_ulysses_pack": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "vap_ulysses_unpack": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "vap_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, _FLOAT_P, c_void_p, c_void_p, c_void_p, _FLOAT_P, c_int, c_int, c_int,
                                  c_int, c_int, c_void_p, c_float, c_void_p]),
    "vap_cfg_flow_match_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64, c_float, c_float, c_void_p]),
    "vap_wan_modulation": (c_int, [c_void_p, c_int, c_void_p, c_int, _FLOAT_P, c_int64, c_int, c_int, c_int, c_void_p]),
    "vap_debug_set_attention_trace": (c_int, [c_void_p]),
    "vap_probe_umma": (c_int, [c_void_p, c_void_p, _FLOAT_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load (once) and return the library with argtypes/restype set.  Fails loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise VapError(
            f"{LIB_PATH} not found: the sm_100a CUDA library is not built. Run `python video-as-prompt_b200/csrc/build.py` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so is stale w.r.t. the header
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    """rc != 0 -> VapError with the library's message.  (A kernel that dies LATER with "unspecified launch failure" hit the mbarrier watchdog —
    a protocol bug; `python video-as-prompt_b200/csrc/build.py --debug` builds libvap_b200_debug.so, which prints the barrier, block and thread
    before it traps: re-run with VAP_B200_LIB pointing at it.)"""
    if rc != 0:
        msg = load().vap_last_error().decode("utf-8", "replace")
        raise VapError(f"{what} failed (rc={rc}): {msg}")
