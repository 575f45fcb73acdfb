"""ctypes binding of libtokred_sm100a.so (the C ABI declared in include/tokred.h).

There is NO fallback: if the shared library is missing or an op is invoked without a CUDA device the call
raises.  The library is built in-tree by ``python -m tokenreduction_b200.build`` (see __graft_entry__.build).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libtokred_sm100a.so")
ABI_VERSION = 4

F32, BF16 = 0, 1

# name -> argtypes; every entry point returns int except where noted.  Mirrors include/tokred.h one to one.
_P = c_void_p
SIGNATURES = {
    "tokred_topk_gather": [_P, c_int, _P, c_int, c_int64, c_int64, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P],
    "tokred_evit_select_fuse": [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    "tokred_tome_effective_r": [c_int, c_int, c_int],
    "tokred_tome_match": [_P, c_int, c_int, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    "tokred_tome_merge": [_P, c_int, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, c_int, _P],
    "tokred_topk_gather_add": [_P, _P, _P, c_int, c_int64, c_int64, c_int, c_int, c_int, c_int, _P, _P, _P],
    "tokred_evit_select_fuse_add": [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    "tokred_tome_merge_ln": [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, c_float, _P, _P, _P, _P, _P],
    "tokred_pairwise_dist": [_P, c_int, c_int, c_int, c_float, c_int, _P, _P],
    "tokred_dpcknn_cluster": [_P, c_int64, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P],
    "tokred_dpcknn_merge": [_P, c_int64, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    "tokred_attn_colsum": [_P, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "tokred_kmedoids_fit": [_P, c_int64, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    "tokred_kmedoids_fit_init": [_P, c_int64, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    "tokred_sinkhorn_merge": [_P, c_int, c_int64, _P, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_int, _P, c_int, _P, _P, c_size_t, _P],
    "tokred_patchmerger": [_P, c_int, c_int64, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, c_float, c_int, _P, c_int, _P, _P, c_size_t, _P],
    "tokred_sit_merge": [_P, c_int, c_int64, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, c_size_t, _P],
    "tokred_soft_merge_workspace_bytes": [c_int, c_int, c_int, c_int],
    "tokred_ats_sample": [_P, c_int, c_int64, c_int64, c_int64, _P, c_int64, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, _P],
    "tokred_gather_rows": [_P, c_int, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "tokred_dyvit_pool_concat": [_P, c_int, c_int64, _P, c_int, c_int, c_int, c_float, _P, c_int, _P],
    "tokred_add_layernorm": [_P, _P, c_int, _P, _P, c_float, c_int64, c_int, _P, _P, _P],
    "tokred_patchify": [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "tokred_embed_layernorm": [_P, c_int, _P, c_int64, _P, _P, _P, c_float, c_int, c_int, c_int, c_int, _P, _P, _P],
    "tokred_residual_add": [_P, _P, c_int64, _P, _P],
    "tokred_attention": [_P, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, c_int64, c_int, _P, _P, _P, _P],
}
EXPORTS = ["tokred_abi_version", "tokred_last_error", "tokred_launch_count", *SIGNATURES]

_lib = None


class TokredError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load (once) and type the shared library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TokredError(
            f"{LIB_PATH} not found: the CUDA extension has not been built "
            "(run `python -m tokenreduction_b200.build`); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    lib.tokred_abi_version.restype = c_int
    lib.tokred_last_error.restype = c_char_p
    lib.tokred_launch_count.restype = c_uint64
    if lib.tokred_abi_version() != ABI_VERSION:
        raise TokredError(f"ABI mismatch: library {lib.tokred_abi_version()} != binding {ABI_VERSION}")
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_size_t if name.endswith("_workspace_bytes") else c_int
    _lib = lib
    return lib


# Optional launch timeline for bench.py: when a list is installed here every entry-point call is bracketed by two
# CUDA events recorded on the launching (current) stream; bench.py reads the elapsed times after it synchronises.
TIMELINE = None


def call(name: str, *args) -> None:
    """Invoke an entry point and turn a non-zero return into a TokredError carrying tokred_last_error()."""
    lib = load()
    if TIMELINE is not None:
        import torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        rc = getattr(lib, name)(*args)
        ev1.record()
        TIMELINE.append((name, args, ev0, ev1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.tokred_last_error().decode("utf-8", "replace")
        kind = "argument" if rc < 0 else f"cuda error {rc}"
        raise TokredError(f"{name} failed ({kind}): {msg}")


def launch_count() -> int:
    return int(load().tokred_launch_count())
