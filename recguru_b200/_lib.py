"""ctypes binding of librecguru_b200.so (the C ABI declared in include/recguru_b200.h).

There is no CPU fallback anywhere in this package: if the shared library is missing the
import of any op fails loudly with instructions to build it.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "librecguru_b200.so"

ABI_VERSION = 6
MAX_SITES = 8
LOSS_KINDS = {"bpr": 0, "bpr_sas": 1, "sce": 2}
LOSS_SKIP_MASKED = 0x100
TIE_MODES = {"id": 0, "target_first": 1, "target_last": 2}

c_i64 = C.c_int64
c_ptr = C.c_void_p


class LookupSite(C.Structure):
    """recguru_lookup_site_t"""
    _fields_ = [
        ("ids", c_ptr), ("count", c_i64), ("X", c_ptr), ("coef", c_ptr),
        ("scale", c_ptr), ("scale2", c_ptr), ("group", c_i64), ("live", c_ptr),
    ]


# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "recguru_abi_version": (C.c_int, []),
    "recguru_last_error": (C.c_char_p, []),
    "recguru_launch_count": (c_i64, []),
    "recguru_gather_pe_mask": (C.c_int, [c_ptr, c_i64, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr,
                                         c_i64, c_ptr, c_ptr, c_ptr]),
    "recguru_pad_mask": (C.c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "recguru_pe_mask": (C.c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr]),
    "recguru_gather_pe_mask_dropout": (C.c_int, [c_ptr, c_i64, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_i64,
                                                 C.c_double, C.c_uint64, C.c_uint64, c_i64, c_ptr, c_ptr, c_ptr]),
    "recguru_dropout": (C.c_int, [c_ptr, c_ptr, c_i64, c_i64, C.c_double, C.c_uint64, C.c_uint64, c_i64, c_ptr,
                                  c_ptr]),
    "recguru_pe_grad": (C.c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, C.c_int, c_ptr, c_ptr]),
    "recguru_extract_last": (C.c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr]),
    "recguru_embedding_bwd_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64]),
    "recguru_embedding_bwd": (C.c_int, [C.POINTER(LookupSite), C.c_int, c_i64, c_i64, c_i64, c_ptr,
                                        C.c_int, c_ptr, c_i64, c_ptr]),
    "recguru_embedding_bwd_plan": (C.c_int, [C.POINTER(LookupSite), C.c_int, c_i64, c_i64, c_ptr, c_i64,
                                             c_ptr]),
    "recguru_embedding_bwd_plan_padmask": (C.c_int, [C.POINTER(LookupSite), C.c_int, c_i64, c_i64, c_ptr, c_i64,
                                                     c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "recguru_embedding_bwd_apply": (C.c_int, [C.POINTER(LookupSite), C.c_int, c_i64, c_i64, c_ptr,
                                              C.c_int, c_ptr, c_i64, c_ptr]),
    "recguru_fill_zero": (C.c_int, [c_ptr, c_i64, c_ptr]),
    "recguru_sort_ids_workspace_bytes": (c_i64, [c_i64, c_i64]),
    "recguru_sort_ids": (C.c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "recguru_sampled_workspace_bytes": (c_i64, [c_i64, c_i64]),
    "recguru_sampled_score_loss_fwd": (C.c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_i64,
                                                 c_i64, c_ptr, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                                                 c_ptr, c_ptr, c_i64, c_ptr]),
    "recguru_sampled_loss_fwd": (C.c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, C.c_int, c_ptr, c_ptr,
                                           c_ptr, c_ptr, c_i64, c_ptr]),
    "recguru_sampled_score_bwd_dh": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64,
                                               c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "recguru_candidate_scores_rank": (C.c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_i64,
                                              c_ptr, c_ptr, c_ptr]),
    "recguru_cast_bf16": (C.c_int, [c_ptr, c_i64, c_ptr, c_ptr]),
    "recguru_catalog_target_scores": (C.c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_i64, c_ptr,
                                                c_ptr, c_ptr]),
    "recguru_catalog_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64, c_i64]),
    "recguru_catalog_topk_rank": (C.c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr,
                                            c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "recguru_catalog_topk_rank_ties": (C.c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr,
                                                 c_i64, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "recguru_catalog_scores_dense": (C.c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "recguru_topk_merge": (C.c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr,
                                     c_ptr]),
    "recguru_rank_histogram": (C.c_int, [c_ptr, c_i64, c_i64, c_ptr, C.c_int, c_ptr]),
    "recguru_sample_negatives": (C.c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_i64,
                                           C.c_uint64, C.c_uint64, c_ptr, c_ptr]),
    "recguru_pad_mask_count": (C.c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr]),
    "recguru_sampled_score_loss_fwd_dh": (C.c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_i64, c_ptr,
                                                    C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64,
                                                    c_ptr, c_ptr, c_ptr, c_ptr]),
    "recguru_sampled_score_loss_fwd_dh_rows": (C.c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_i64, c_ptr,
                                                         C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64,
                                                         c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "recguru_embedding_bwd_plan_stage": (C.c_int, [C.POINTER(LookupSite), C.c_int, c_i64, c_i64, c_ptr, c_i64,
                                                   C.c_int, c_ptr]),
    "recguru_embedding_bwd_plan_live_rows": (C.c_int, [C.POINTER(LookupSite), C.c_int, c_i64, c_ptr, c_i64,
                                                       C.POINTER(c_ptr), C.POINTER(c_ptr)]),
    "recguru_live_rows_workspace_bytes": (c_i64, [c_i64]),
    "recguru_live_rows": (C.c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "recguru_sample_negatives_devstep": (C.c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_ptr,
                                                   c_i64, C.c_uint64, c_ptr, c_ptr, c_ptr]),
    "recguru_sample_negatives_csr": (C.c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_ptr,
                                               c_i64, C.c_uint64, C.c_uint64, c_ptr, c_ptr, c_ptr]),
    "recguru_counter_add": (C.c_int, [c_ptr, C.c_uint64, c_ptr]),
    "recguru_widen_ids": (C.c_int, [c_ptr, c_i64, c_ptr, c_ptr]),
    "recguru_debug_timestamp": (C.c_int, [c_ptr, c_ptr]),
    "recguru_embedding_bwd_mark_rows": (C.c_int, [C.POINTER(LookupSite), C.c_int, c_i64, c_ptr, c_i64, c_ptr, C.c_uint32, c_ptr]),
    "recguru_table_adam_rows": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_uint32, c_i64, c_i64, C.c_double,
                                          C.c_double, C.c_double, C.c_double, c_i64, c_ptr]),
    "recguru_build_batches": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, C.c_int, c_ptr, c_ptr,
                                        c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "recguru_table_adam": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, C.c_double, C.c_double, C.c_double,
                                     C.c_double, c_i64, c_ptr]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -m recguru_b200.build` (needs nvcc; cross-compiles sm_100a without a GPU). "
            "recguru_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    got = lib.recguru_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {got}, expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


class RecGuruError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().recguru_last_error().decode("utf-8", "replace")
        raise RecGuruError(f"{what} failed with code {rc}: {msg}")


def launch_count() -> int:
    return int(load().recguru_launch_count())
