"""The C-ABI library builds, loads and exports every symbol include/recguru_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "recguru_b200.h").read_text()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    decls = re.findall(r"\b(recguru_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", body, flags=re.S)
    out = {}
    for name, args in decls:
        args = " ".join(args.split())
        n = 0 if args in ("", "void") else args.count(",") + 1
        out[name] = n
    return out


def test_header_declares_expected_entry_points():
    fns = declared_functions()
    for must in ("recguru_gather_pe_mask", "recguru_embedding_bwd", "recguru_sampled_score_loss_fwd",
                 "recguru_catalog_topk_rank", "recguru_topk_merge", "recguru_rank_histogram"):
        assert must in fns
    assert len(fns) >= 19


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_ctypes_table_matches_header(built_lib):
    from recguru_b200 import _lib
    fns = declared_functions()
    assert set(_lib.SIGNATURES) == set(fns)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert len(argtypes) == fns[name], f"{name}: ctypes has {len(argtypes)} args, header {fns[name]}"
    lib = _lib.load()
    assert lib.recguru_abi_version() == _lib.ABI_VERSION
    assert lib.recguru_last_error() == b""
    assert _lib.launch_count() >= 0          # (0 in a fresh process; other tests may have launched kernels)
    assert ctypes.sizeof(_lib.LookupSite) == 64      # 8 fields of 8 bytes (recguru_lookup_site_t)


def test_argument_validation_without_gpu(built_lib):
    """Shape/argument checks return the documented negative codes before any CUDA call."""
    from recguru_b200 import _lib
    lib = _lib.load()
    assert lib.recguru_gather_pe_mask(None, 10, 32, None, 4, 2, None, None, 0, None, None, None) == -1
    assert b"null" in lib.recguru_last_error()
    assert lib.recguru_embedding_bwd_workspace_bytes(-1, 10, 32) == -1
    assert lib.recguru_embedding_bwd_workspace_bytes(1000, 100, 64) > 0
    assert lib.recguru_catalog_workspace_bytes(100, 100, 128, 64) == -1      # k > 32
    assert lib.recguru_sort_ids_workspace_bytes(4096, 1 << 20) > 4096 * 16
    with pytest.raises(_lib.RecGuruError):
        _lib.check(-2, "demo")


def test_round2_entry_points_validate_without_gpu(built_lib):
    """The entry points added in round 2 check their arguments before any CUDA call as well."""
    import ctypes as C
    from recguru_b200 import _lib
    lib = _lib.load()
    # live-row list without a plan
    assert lib.recguru_live_rows_workspace_bytes(-1) == -1
    assert lib.recguru_live_rows_workspace_bytes(204800) >= 4 * 200
    assert lib.recguru_live_rows(None, 8, None, None, None, 0, None) == -1          # n_live_rows is null
    assert b"null" in lib.recguru_last_error()
    # two-stage plan: a stage outside {ALL, COMPACT, SORT}; live-row pointers need row mode
    site = (_lib.LookupSite * 1)()
    site[0].group = 1                                    # an empty site (count 0): valid, nothing to plan
    assert lib.recguru_embedding_bwd_plan_stage(site, 1, 100, 0, None, 0, 7, None) == -1
    assert b"stage" in lib.recguru_last_error()
    rows, n_rows = C.c_void_p(), C.c_void_p()
    rc = lib.recguru_embedding_bwd_plan_live_rows(site, 1, 100, None, 0, C.byref(rows), C.byref(n_rows))
    assert rc == -2 and b"mask" in lib.recguru_last_error()                           # no sites with a shared mask
    # scoring over a live-row list: the list pointers are mandatory, the kind must skip masked positions
    assert lib.recguru_sampled_score_loss_fwd_dh_rows(None, None, 10, 32, None, None, 4, 2, None, 0, None, None, None,
                                                      None, None, None, 0, None, None, None, None, None, None) == -1
    assert b"live_rows" in lib.recguru_last_error()
    # dropout parameters
    assert lib.recguru_dropout(None, None, 16, 32, 1.5, 0, 0, 1024, None, None) == -1
    assert b"dropout" in lib.recguru_last_error()


def test_sources_are_sm100a_only():
    from recguru_b200 import build
    assert "arch=compute_100a,code=sm_100a" in build.NVCC_FLAGS
    assert "-lineinfo" in build.NVCC_FLAGS
    assert len(build.sources()) >= 5
