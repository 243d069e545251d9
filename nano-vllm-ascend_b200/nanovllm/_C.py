"""ctypes binding of ``libb200attn.so`` -- the C-ABI boundary of the hot path.

Every prototype below mirrors ``include/b200attn.h`` one to one. There is deliberately NO
fallback: if the shared library is missing or a call returns non-zero the caller gets an
exception (the reference raises from torch dispatch the same way when ``torch_npu`` is
absent, ``layers/attention.py:2-3`` of the reference).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_LIB_ENV = "B200ATTN_LIB"
_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "lib", "libb200attn.so"))

BF16, F16 = 0, 1

# name -> (restype, argtypes); the single source of truth for the exported symbol list.
PROTOTYPES = {
    "b200_version": (c_int32, []),
    "b200_error_string": (c_char_p, [c_int32]),
    "b200_launch_count": (c_int64, []),
    "b200_reset_launch_count": (None, []),
    "b200_reshape_and_cache": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_int32, c_int32, c_int32, c_int64, c_int64, c_int32,
                                         c_int32, c_int32, c_void_p]),
    "b200_reshape_and_cache_2d": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_int32, c_int32, c_int32, c_int64, c_int64, c_int32,
                                            c_int32, c_int32, c_void_p]),
    "b200_paged_decode_attn_workspace_size": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "b200_paged_decode_attn": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_int32, c_int32, c_int32, c_int32, c_int64,
                                         c_int64, c_int32, c_int32, c_int32, c_float, c_void_p,
                                         c_size_t, c_int32, c_int32, c_void_p]),
    "b200_reshape_and_cache_fp8": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_int32, c_int32, c_int32, c_int32, c_int64, c_int64,
                                             c_int32, c_int32, c_float, c_float, c_int32, c_void_p]),
    "b200_paged_decode_attn_fp8": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_int32, c_int32, c_int32, c_int32, c_int64,
                                             c_int64, c_int32, c_int32, c_int32, c_float, c_float,
                                             c_float, c_void_p, c_size_t, c_int32, c_int32, c_void_p]),
    "b200_prefill_varlen_attn": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                           c_int32, c_int32, c_int32, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                           c_int32, c_int32, c_int32, c_float, c_int32,
                                           c_void_p, c_int64, c_void_p]),
    "b200_prefill_workspace_size": (c_int64, [c_int32]),
    "b200_prefill_set_query_tiles": (c_int32, [c_int32]),
    "b200_gather_dequant_fp8": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                          c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_float,
                                          c_float, c_int32, c_void_p]),
    "b200_qknorm_rope_cache": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_float, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                         c_int32, c_int32, c_int32, c_int64, c_int32, c_int32,
                                         c_int32, c_int32, c_void_p]),
    "b200_rmsnorm": (c_int32, [c_void_p, c_void_p, c_void_p, c_float, c_int32, c_int32, c_int32,
                               c_void_p]),
    "b200_add_rmsnorm": (c_int32, [c_void_p, c_void_p, c_void_p, c_float, c_int32, c_int32,
                                   c_int32, c_void_p]),
    "b200_silu_and_mul": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "b200_sample_shard": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int32,
                                    c_uint64, c_uint64, c_void_p, c_int32, c_void_p]),
    "b200_moe_route": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int64,
                                 c_int32, c_void_p]),
    "b200_moe_align": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                 c_int32, c_void_p]),
    "b200_moe_gemm_gate_up": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                        c_int32, c_int32, c_int32, c_int64, c_int32, c_void_p]),
    "b200_moe_gemm_down": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "b200_moe_gather_rows": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                       c_int32, c_int64, c_int32, c_void_p]),
    "b200_moe_gemm_gate_up_umma": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                             c_int32, c_int32, c_int32, c_void_p]),
    "b200_moe_gemm_down_umma": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "b200_moe_combine": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                   c_int64, c_int32, c_void_p]),
    "b200_allreduce_add_rmsnorm": (c_int32, [c_void_p, c_void_p, c_void_p, c_float, c_int32, c_int32,
                                             c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                             c_void_p]),
    "b200_allreduce_add_rmsnorm_twoshot": (c_int32, [c_void_p, c_void_p, c_void_p, c_float, c_int32,
                                                     c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                                     c_int32, c_int32, c_int32, c_void_p]),
    "b200_allreduce_add_rmsnorm_nvls": (c_int32, [c_void_p, c_void_p, c_void_p, c_float, c_int32,
                                                  c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                                  c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "b200_allreduce_add_rmsnorm_nvls_fused": (c_int32, [c_void_p, c_void_p, c_float, c_int32, c_int32,
                                                        c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                                        c_int32, c_int32, c_void_p]),
    "b200_allreduce_add_rmsnorm_nvls_oneshot": (c_int32, [c_void_p, c_void_p, c_void_p, c_float, c_int32,
                                                          c_int32, c_void_p, c_void_p, c_void_p, c_int32,
                                                          c_int32, c_int32, c_void_p]),
    "b200_comm_alloc": (c_int32, [c_size_t, c_void_p]),
    "b200_comm_free": (c_int32, [c_void_p]),
    "b200_comm_get_handle": (c_int32, [c_void_p, c_void_p]),
    "b200_comm_open_handle": (c_int32, [c_void_p, c_void_p]),
    "b200_comm_close_handle": (c_int32, [c_void_p]),
    "b200_sample": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_uint64,
                              c_uint64, c_void_p, c_int32, c_void_p]),
}

_lib = None


class B200Error(RuntimeError):
    """A libb200attn entry point returned a non-zero status."""


def lib_path() -> str:
    return os.environ.get(_LIB_ENV, DEFAULT_LIB_PATH)


def load() -> ctypes.CDLL:
    """Load the library once; raises (never falls back) if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise B200Error(
            f"libb200attn.so not found at {path}: build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` (or make -C "
            f"nano-vllm-ascend_b200/csrc). There is no CPU/eager fallback for the hot path.")
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    built = lib.b200_version()
    if built != 100:
        raise B200Error(f"libb200attn.so version {built} does not match the binding (100)")
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().b200_error_string(code).decode()
        raise B200Error(f"{what} failed with status {code}: {msg}")
