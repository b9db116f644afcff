"""ctypes binding of libqdx.so (the C ABI declared in include/qdx.h).

There is no CPU fallback: if the shared library is missing or a call fails, we raise.  The library is built
in-tree by `__graft_entry__.build()` / `make -C qdax_b200/csrc`.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QDX_LIB_PATH", os.path.join(_HERE, "libqdx.so"))   # override: A/B builds of the same ABI

ERRORS = {
    -1: "QDX_ERR_ARG: bad argument",
    -2: "QDX_ERR_UNSUPPORTED: shape outside the fused path",
    -3: "QDX_ERR_EMPTY_REPERTOIRE: selection from an all-empty repertoire",
    -4: "QDX_ERR_BAD_CELL: cell index out of range",
    -5: "QDX_ERR_BAD_INDEX: winner index outside the offspring buffer",
    -6: "QDX_ERR_PEER_TIMEOUT: a peer's keys did not arrive (peer-memory exchange)",
    -7: "QDX_ERR_INTERNAL: an in-kernel wait between CTAs timed out",
}


class QdxError(RuntimeError):
    def __init__(self, fn: str, rc: int):
        self.rc = rc
        msg = ERRORS.get(rc, f"cudaError_t {rc}" if rc > 0 else f"error {rc}")
        super().__init__(f"{fn} failed: {msg}")


class GridDesc(C.Structure):
    _fields_ = [
        ("dd", C.c_int32),
        ("n", C.c_int32 * 4),
        ("stride", C.c_int32 * 4),
        ("lo", C.c_float * 4),
        ("hi", C.c_float * 4),
        ("axes", C.c_void_p),
    ]


class CvtIndexDesc(C.Structure):
    _fields_ = [
        ("dd", C.c_int32),
        ("g", C.c_int32 * 3),
        ("lo", C.c_float * 3),
        ("h", C.c_float * 3),
        ("start", C.c_void_p),
        ("ids", C.c_void_p),
        ("pts", C.c_void_p),
    ]


MAX_LEAVES = 32


class LeafTable(C.Structure):
    """qdx_leaf_table (include/qdx.h): leaf offsets inside the packed genotype row + per-leaf noise keys."""

    _fields_ = [
        ("n", C.c_int32),
        ("off", C.c_int32 * (MAX_LEAVES + 1)),
        ("key", C.c_uint32 * (2 * MAX_LEAVES)),
    ]


_i32, _i64, _u32, _f32, _vp = C.c_int32, C.c_int64, C.c_uint32, C.c_float, C.c_void_p


class StepDesc(C.Structure):
    """qdx_step_desc (include/qdx.h): everything one generation needs, filled once per (repertoire buffers, configuration)."""

    _fields_ = [
        ("rep_genotypes", _vp), ("rep_fitness", _vp), ("rep_desc", _vp), ("centroids", _vp), ("ws", _vp),
        ("K", _i64), ("D", _i64), ("B", _i64),
        ("desc_dim", _i32), ("task", _i32),
        ("iso_sigma", _f32), ("line_sigma", _f32),
        ("has_min", _i32), ("minval", _f32), ("has_max", _i32), ("maxval", _f32),
        ("grid", C.POINTER(GridDesc)), ("cvt", C.POINTER(CvtIndexDesc)),
        ("tc_prep", _vp), ("tc_scratch", _vp),
        ("first_wins", _i32), ("qd_offset", _f32),
        ("off_genotypes", _vp), ("off_fitness", _vp), ("off_desc", _vp), ("off_cells", _vp),
        ("rank", _i32), ("nranks", _i32), ("exchange", _i32),
    ]


# name -> argtypes, exactly the prototypes of include/qdx.h
PROTOTYPES = {
    "qdx_version": [],
    "qdx_workspace_bytes": [_i64, C.POINTER(_i64)],
    "qdx_workspace_keytab_offset": [_i64, C.POINTER(_i64)],
    "qdx_workspace_init": [_vp, _i64, _vp],
    "qdx_workspace_set_carry_key": [_vp, _u32, _u32, _vp],
    "qdx_workspace_copy_carry_key": [_vp, _vp, _i32, _vp],
    "qdx_workspace_set_error_mirror": [_vp, _vp, _vp],
    "qdx_workspace_read": [_vp, C.POINTER(_u32), C.POINTER(_f32), C.POINTER(_i32), _vp],
    "qdx_select_prepare": [_vp, _i64, _vp, _i32, _u32, _u32, _i32, _vp],
    "qdx_regenerate_winners": [_vp, _i64, _i64, _i64, _i32, _vp, _f32, _f32, _i32, _f32, _i32, _f32, _i32, _vp, _vp],
    "qdx_generate": [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _f32, _f32, _i32, _f32, _i32, _f32, _i32, _i32,
                     C.POINTER(GridDesc), _i32, _u32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_u32), C.POINTER(CvtIndexDesc), _i32, _vp],
    "qdx_elect_winners": [_vp, _i64, _i64, _i32, _i32, _i64, _i32, _vp, _f32, _f32, _i32, _f32, _i32, _f32, _i32, _vp, _vp, _vp,
                          _i32, _vp],
    "qdx_xchg_bytes": [_i64, _i64, _i64, _i32, C.POINTER(_i64)],
    "qdx_xchg_create": [_i64, _i64, _i64, _i32, C.POINTER(_vp), _vp],
    "qdx_xchg_open": [_vp, C.POINTER(_vp)],
    "qdx_xchg_close": [_vp],
    "qdx_xchg_destroy": [_vp],
    "qdx_xchg_attach": [_vp, _i32, _i32, C.POINTER(_vp), _i64, _i64, _i32, _i32, _vp],
    "qdx_xchg_push": [_vp, _i64, C.POINTER(_u32), _vp],
    "qdx_map_elites_step": [C.POINTER(StepDesc), _i32, _u32, _u32, C.POINTER(_u32), _vp, _vp],
    "qdx_host_split": [_u32, _u32, _i32, _vp],
    "qdx_host_generation_keys": [_i32, _u32, _u32, C.POINTER(_u32), C.POINTER(_u32)],
    "qdx_score": [_i32, _vp, _i64, _i64, _i32, _vp, _vp, _vp],
    "qdx_score_noisy_arm": [_vp, _i64, _i64, _u32, _u32, _f32, _f32, _f32, _vp, _vp, _vp],
    "qdx_cells": [_vp, _i64, _i32, _vp, _i64, C.POINTER(GridDesc), _vp, _vp, _vp, _vp, _i32, _u32, _i32, _vp],
    "qdx_cvt_index_plan": [_vp, _i64, _i32, C.POINTER(CvtIndexDesc), C.POINTER(_i64)],
    "qdx_cvt_index_build": [_vp, _i64, C.POINTER(CvtIndexDesc), _vp, _vp, _vp],
    "qdx_cells_indexed": [_vp, _i64, C.POINTER(CvtIndexDesc), _i64, _vp, _vp, _vp, _vp, _i32, _u32, _i32, _vp],
    "qdx_cells_tc_workspace": [_i64, _i64, C.POINTER(_i64), C.POINTER(_i64)],
    "qdx_cells_tc_prepare": [_vp, _i64, _i32, _vp, _vp],
    "qdx_cells_tc": [_vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _u32, _i32, _vp],
    "qdx_offer_cells": [_vp, _vp, _i64, _i64, _vp, _vp, _u32, _i32, _vp],
    "qdx_commit": [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _u32, _i64, _i32, _vp, _vp, _vp, _f32, _vp, _vp, _i32, _vp],
    "qdx_select_indices": [_vp, _u32, _u32, _i64, _vp, _vp],
    "qdx_select_indices_without_replacement": [_vp, _i64, _vp, _u32, _u32, _i64, _vp, _vp, _vp],
    "qdx_gather_rows": [_vp, _vp, _i64, _i64, _vp, _vp],
    "qdx_isoline_variation": [_vp, _vp, _i64, _i64, _u32, _u32, _f32, _f32, _i32, _f32, _i32, _f32, _vp, _vp],
    "qdx_generate_leaves": [_vp, _vp, _vp, _i64, _i64, _i64, _f32, _f32, _i32, _f32, _i32, _f32, _vp, _vp, _vp, C.POINTER(_u32),
                            C.POINTER(LeafTable), _vp],
    "qdx_isoline_variation_leaves": [_vp, _vp, _i64, _i64, _u32, _u32, C.POINTER(LeafTable), _f32, _f32, _i32, _f32, _i32, _f32, _vp, _vp],
    "qdx_copy_2d": [_vp, _i64, _vp, _i64, _i64, _i64, _vp],
    "qdx_mels_offer": [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp],
    "qdx_mome_add": [_vp, _vp, _vp, _i64, _i32, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _i64, _i32, _vp],
    "qdx_unstructured_scratch": [_i64, _i64, C.POINTER(_i64)],
    "qdx_unstructured_plan": [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _f32, _vp, _vp, _vp],
    "qdx_unstructured_offer": [_vp, _vp, _i64, _i32, _i64, _f32, _vp, _vp, _vp, _vp, _i32, _vp],
    "qdx_scatter_rows_by_source": [_vp, _vp, _i64, _i64, _vp, _vp],
    "qdx_kmeans_accumulate": [_vp, _vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp],
    "qdx_kmeans_update": [_vp, _vp, _vp, _i64, _i32, _vp, _vp],
    "qdx_polynomial_mutation": [_vp, _i64, _i64, _u32, _u32, _i32, _f32, _f32, _f32, _f32, _vp, _vp],
    "qdx_polynomial_crossover": [_vp, _vp, _i64, _i64, _u32, _u32, _i32, _vp, _vp],
    "qdx_random": [_u32, _u32, _i64, _i32, _f32, _f32, _vp, _vp],
    "qdx_metrics": [_vp, _i64, _f32, _vp, _vp],
    "qdx_dns_add": [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp],
    "qdx_host_select_table": [_i32, _vp, C.POINTER(_i32)],
    "qdx_host_select_rank": [_i32, _vp, _i64, _vp],
}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load libqdx.so (once).  Raises if the CUDA extension has not been built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C qdax_b200/csrc`. qdax_b200 has no CPU fallback."
            )
        h = C.CDLL(LIB_PATH)
        for name, argtypes in PROTOTYPES.items():
            fn = getattr(h, name)  # AttributeError if the symbol is missing
            fn.argtypes = argtypes
            fn.restype = C.c_int
        _lib = h
    return _lib


# every C-ABI call that launches at least one of OUR kernels: name -> launches per call (bench.py `gpu_launches`)
KERNEL_LAUNCHES = {"qdx_select_prepare": 1, "qdx_regenerate_winners": 1, "qdx_elect_winners": 1, "qdx_xchg_push": 1, "qdx_generate": 1, "qdx_generate_leaves": 1, "qdx_isoline_variation_leaves": 1, "qdx_copy_2d": 1, "qdx_kmeans_accumulate": 1, "qdx_kmeans_update": 1, "qdx_mels_offer": 1, "qdx_mome_add": 1, "qdx_unstructured_plan": 3, "qdx_unstructured_offer": 3, "qdx_scatter_rows_by_source": 1, "qdx_score": 1, "qdx_score_noisy_arm": 1, "qdx_cells": 1, "qdx_cells_indexed": 1, "qdx_cells_tc": 2, "qdx_cells_tc_prepare": 1, "qdx_offer_cells": 1,
                   "qdx_commit": 1, "qdx_select_indices": 1, "qdx_select_indices_without_replacement": 2, "qdx_gather_rows": 1, "qdx_isoline_variation": 1, "qdx_polynomial_mutation": 1, "qdx_polynomial_crossover": 1,
                   "qdx_random": 1, "qdx_metrics": 1, "qdx_dns_add": 3}
launch_count = 0


def call(name: str, *args) -> None:
    global launch_count
    launch_count += KERNEL_LAUNCHES.get(name, 0)
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise QdxError(name, rc)
