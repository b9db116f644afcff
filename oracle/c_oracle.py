"""ORACLE (test infrastructure only): ctypes binding of oracle/liboracle.so (qdx_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  The product package never does.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

F32 = np.float32
TASKS = {"arm": 0, "rastrigin": 1, "sphere": 2}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "qdx_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=F32)


def _key(k) -> np.ndarray:
    return np.ascontiguousarray(k, dtype=np.uint32).reshape(2)


def _chk(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed rc={rc}")


def set_threads(n: int) -> None:
    lib().qo_set_threads(C.c_int(n))


def get_threads() -> int:
    return int(lib().qo_get_threads())


def threefry2x32(k0, k1, c0, c1) -> Tuple[int, int]:
    out = np.zeros(2, dtype=np.uint32)
    lib().qo_threefry2x32(C.c_uint32(k0), C.c_uint32(k1), C.c_uint32(c0), C.c_uint32(c1), _p(out))
    return int(out[0]), int(out[1])


def split(key, num: int = 2) -> np.ndarray:
    out = np.zeros((num, 2), dtype=np.uint32)
    lib().qo_split(_p(_key(key)), C.c_int64(num), _p(out))
    return out


def random_bits(key, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint32)
    lib().qo_random_bits(_p(_key(key)), C.c_int64(n), _p(out))
    return out


def uniform(key, n: int, minval=0.0, maxval=1.0) -> np.ndarray:
    out = np.zeros(n, dtype=F32)
    lib().qo_uniform(_p(_key(key)), C.c_int64(n), C.c_float(minval), C.c_float(maxval), _p(out))
    return out


def normal(key, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=F32)
    lib().qo_normal(_p(_key(key)), C.c_int64(n), _p(out))
    return out


def math_probe(which: int, x) -> np.ndarray:
    x = _f(x)
    out = np.zeros_like(x)
    _chk(lib().qo_math_probe(C.c_int(which), C.c_int64(x.size), _p(x), _p(out)), "math_probe")
    return out


def math_probe2(which: int, x, aux: float = 0.0) -> np.ndarray:
    x = _f(x)
    out = np.zeros_like(x)
    _chk(lib().qo_math_probe2(C.c_int(which), C.c_int64(x.size), _p(x), C.c_float(aux), _p(out)), "math_probe2")
    return out


def polynomial_mutation(x, key, proportion_to_mutate, eta, minval, maxval) -> np.ndarray:
    x = _f(x)
    B, D = x.shape
    out = np.zeros_like(x)
    _chk(lib().qo_polynomial_mutation(_p(x), C.c_int64(B), C.c_int64(D), _p(_key(key)), C.c_float(proportion_to_mutate), C.c_float(eta),
                                      C.c_float(minval), C.c_float(maxval), _p(out)), "polynomial_mutation")
    return out


def polynomial_crossover(x1, x2, key, proportion_var_to_change) -> np.ndarray:
    x1, x2 = _f(x1), _f(x2)
    B, D = x1.shape
    out = np.zeros_like(x1)
    _chk(lib().qo_polynomial_crossover(_p(x1), _p(x2), C.c_int64(B), C.c_int64(D), _p(_key(key)), C.c_float(proportion_var_to_change), _p(out)),
         "polynomial_crossover")
    return out


def select_indices(fitnesses, key, num: int) -> np.ndarray:
    f = _f(fitnesses).reshape(-1)
    out = np.zeros(num, dtype=np.int32)
    _chk(lib().qo_select_indices(_p(f), C.c_int64(f.size), _p(_key(key)), C.c_int64(num), _p(out)), "select")
    return out


def select_indices_without_replacement(fitnesses, key, num: int) -> np.ndarray:
    f = _f(fitnesses).reshape(-1)
    out = np.zeros(num, dtype=np.int32)
    _chk(lib().qo_select_indices_noreplace(_p(f), C.c_int64(f.size), _p(_key(key)), C.c_int64(num), _p(out)), "select_noreplace")
    return out


def _clip_args(minval, maxval):
    return (
        C.c_int(minval is not None), C.c_float(0.0 if minval is None else minval),
        C.c_int(maxval is not None), C.c_float(0.0 if maxval is None else maxval),
    )


def isoline_variation(x1, x2, key, iso_sigma, line_sigma, minval=None, maxval=None) -> np.ndarray:
    x1, x2 = _f(x1), _f(x2)
    B, D = x1.shape
    out = np.zeros((B, D), dtype=F32)
    _chk(lib().qo_isoline_variation(_p(x1), _p(x2), C.c_int64(B), C.c_int64(D), _p(_key(key)), C.c_float(iso_sigma),
                                    C.c_float(line_sigma), *_clip_args(minval, maxval), _p(out)), "isoline")
    return out


def _offsets(leaf_sizes) -> np.ndarray:
    return np.ascontiguousarray(np.concatenate([[0], np.cumsum(leaf_sizes)]), dtype=np.int32)


def isoline_variation_leaves(x1, x2, key, leaf_sizes, iso_sigma, line_sigma, minval=None, maxval=None) -> np.ndarray:
    """isoline_variation on a pytree genotype given as packed rows (B, sum(leaf_sizes)), leaves in jax.tree.leaves order."""
    x1, x2 = _f(x1), _f(x2)
    B, D = x1.shape
    off = _offsets(leaf_sizes)
    out = np.zeros((B, D), dtype=F32)
    _chk(lib().qo_isoline_variation_leaves(_p(x1), _p(x2), C.c_int64(B), C.c_int64(D), _p(_key(key)), C.c_int(len(leaf_sizes)), _p(off),
                                           C.c_float(iso_sigma), C.c_float(line_sigma), *_clip_args(minval, maxval), _p(out)), "isoline_leaves")
    return out


def emit_isoline_leaves(rep_g, rep_f, key, B, leaf_sizes, iso_sigma, line_sigma, minval=None, maxval=None):
    rep_g, rep_f = _f(rep_g), _f(rep_f).reshape(-1)
    K, D = rep_g.shape
    off = _offsets(leaf_sizes)
    out = np.zeros((B, D), dtype=F32)
    p1 = np.zeros(B, dtype=np.int32)
    p2 = np.zeros(B, dtype=np.int32)
    _chk(lib().qo_emit_isoline_leaves(_p(rep_g), _p(rep_f), C.c_int64(K), C.c_int64(D), _p(_key(key)), C.c_int64(B), C.c_int(len(leaf_sizes)),
                                      _p(off), C.c_float(iso_sigma), C.c_float(line_sigma), *_clip_args(minval, maxval),
                                      _p(out), _p(p1), _p(p2)), "emit_leaves")
    return out, p1, p2


def emit_isoline(rep_g, rep_f, key, B, iso_sigma, line_sigma, minval=None, maxval=None):
    rep_g, rep_f = _f(rep_g), _f(rep_f).reshape(-1)
    K, D = rep_g.shape
    out = np.zeros((B, D), dtype=F32)
    p1 = np.zeros(B, dtype=np.int32)
    p2 = np.zeros(B, dtype=np.int32)
    _chk(lib().qo_emit_isoline(_p(rep_g), _p(rep_f), C.c_int64(K), C.c_int64(D), _p(_key(key)), C.c_int64(B),
                               C.c_float(iso_sigma), C.c_float(line_sigma), *_clip_args(minval, maxval),
                               _p(out), _p(p1), _p(p2)), "emit")
    return out, p1, p2


def score(task: str, g, desc_dim: int = 2):
    g = _f(g)
    B, D = g.shape
    f = np.zeros(B, dtype=F32)
    d = np.zeros((B, desc_dim), dtype=F32)
    _chk(lib().qo_score(C.c_int(TASKS[task]), _p(g), C.c_int64(B), C.c_int64(D), C.c_int64(desc_dim), _p(f), _p(d)), "score")
    return f, d


def noisy_arm(g, key, fit_variance: float, desc_variance: float, params_variance: float):
    """noisy_arm_scoring_function (qdax/tasks/arm.py:53-81) with the exact arithmetic."""
    g = _f(g)
    B, D = g.shape
    f = np.zeros(B, dtype=F32)
    d = np.zeros((B, 2), dtype=F32)
    _chk(lib().qo_noisy_arm(_p(g), C.c_int64(B), C.c_int64(D), _p(_key(key)), C.c_float(fit_variance), C.c_float(desc_variance),
                            C.c_float(params_variance), _p(f), _p(d)), "noisy_arm")
    return f, d


def cells(desc, centroids) -> np.ndarray:
    desc, centroids = _f(desc), _f(centroids)
    B, Dd = desc.shape
    out = np.zeros(B, dtype=np.int32)
    _chk(lib().qo_cells(_p(desc), C.c_int64(B), C.c_int64(Dd), _p(centroids), C.c_int64(centroids.shape[0]), _p(out)), "cells")
    return out


def add(rep_g, rep_f, rep_d, g, f, desc, cell_idx, tie_break: str = "first"):
    """In-place on copies; returns (genotypes, fitnesses (K,), descriptors, scatter_idx)."""
    rep_g, rep_f, rep_d = _f(rep_g).copy(), _f(rep_f).reshape(-1).copy(), _f(rep_d).copy()
    g, f, desc = _f(g), _f(f).reshape(-1), _f(desc)
    cell_idx = np.ascontiguousarray(cell_idx, dtype=np.int32)
    K, D = rep_g.shape
    Dd = rep_d.shape[1]
    B = g.shape[0]
    sidx = np.zeros(B, dtype=np.int32)
    _chk(lib().qo_add(_p(rep_g), _p(rep_f), _p(rep_d), C.c_int64(K), C.c_int64(D), C.c_int64(Dd), _p(g), _p(f), _p(desc),
                      _p(cell_idx), C.c_int64(B), C.c_int(tie_break == "last"), _p(sidx)), "add")
    return rep_g, rep_f, rep_d, sidx


def mels_reduce(cells_all, desc_all, fit_all):
    """(mode cell, spread, mean fitness) per individual from its S evaluations (mels_repertoire.py:26-57, :148-169)."""
    fit_all = _f(fit_all)
    B, S = fit_all.shape
    desc_all = _f(desc_all).reshape(B, S, -1)
    cells_all = np.ascontiguousarray(cells_all, dtype=np.int32).reshape(B, S)
    cell = np.zeros(B, dtype=np.int32)
    spread = np.zeros(B, dtype=F32)
    fmean = np.zeros(B, dtype=F32)
    _chk(lib().qo_mels_reduce(_p(cells_all), _p(desc_all), _p(fit_all), C.c_int64(B), C.c_int64(S), C.c_int64(desc_all.shape[2]),
                              _p(cell), _p(spread), _p(fmean)), "mels_reduce")
    return cell, spread, fmean


def mels_add(rep_g, rep_f, rep_d, rep_spread, centroids, g, desc_all, fit_all, tie_break: str = "first"):
    """MELSRepertoire.add (mels_repertoire.py:89-230) with the exact-arithmetic reduction; returns the new
    (genotypes, fitnesses (K,), descriptors, spreads)."""
    rep_g, rep_f, rep_d, rep_spread = _f(rep_g).copy(), _f(rep_f).reshape(-1).copy(), _f(rep_d).copy(), _f(rep_spread).copy()
    centroids, g, fit_all = _f(centroids), _f(g), _f(fit_all)
    B, S = fit_all.shape
    cell, spread, fmean = mels_reduce(cells(_f(desc_all).reshape(B * S, -1), centroids), desc_all, fit_all)
    cond = (fmean > rep_f[cell]) & (spread <= rep_spread[cell])
    order = range(B - 1, -1, -1) if tie_break == "first" else range(B)
    for b in order:
        if cond[b]:
            c = cell[b]
            rep_g[c], rep_f[c], rep_d[c], rep_spread[c] = g[b], fmean[b], centroids[c], spread[b]
    return rep_g, rep_f, rep_d, rep_spread


def metrics(rep_f, qd_offset: float = 0.0) -> np.ndarray:
    rep_f = _f(rep_f).reshape(-1)
    out = np.zeros(3, dtype=F32)
    lib().qo_metrics(_p(rep_f), C.c_int64(rep_f.size), C.c_float(qd_offset), _p(out))
    return out  # qd_score, max_fitness, coverage


def map_elites_scan(rep_g, rep_f, rep_d, centroids, key, n_iter, B, task="arm", iso_sigma=0.05, line_sigma=0.1,
                    minval=0.0, maxval=1.0, tie_break="first", qd_offset=0.0):
    """Returns (genotypes, fitnesses (K,), descriptors, new key, metrics (n_iter, 3), stage seconds (4,))."""
    rep_g, rep_f, rep_d = _f(rep_g).copy(), _f(rep_f).reshape(-1).copy(), _f(rep_d).copy()
    centroids = _f(centroids)
    K, D = rep_g.shape
    Dd = rep_d.shape[1]
    k = _key(key).copy()
    m = np.zeros((n_iter, 3), dtype=F32)
    secs = np.zeros(4, dtype=np.float64)
    _chk(lib().qo_map_elites_scan(_p(rep_g), _p(rep_f), _p(rep_d), _p(centroids), C.c_int64(K), C.c_int64(D), C.c_int64(Dd),
                                  _p(k), C.c_int64(n_iter), C.c_int64(B), C.c_int(TASKS[task]), C.c_float(iso_sigma),
                                  C.c_float(line_sigma), *_clip_args(minval, maxval), C.c_int(tie_break == "last"),
                                  C.c_float(qd_offset), _p(m), _p(secs)), "scan")
    return rep_g, rep_f, rep_d, k, m, secs


def distributed_update(rep_g, rep_f, rep_d, centroids, keys, B_dev, task="arm", iso_sigma=0.05, line_sigma=0.1,
                       minval=0.0, maxval=1.0, tie_break="first"):
    rep_g, rep_f, rep_d = _f(rep_g).copy(), _f(rep_f).reshape(-1).copy(), _f(rep_d).copy()
    centroids = _f(centroids)
    keys = np.ascontiguousarray(keys, dtype=np.uint32).reshape(-1, 2)
    R = keys.shape[0]
    K, D = rep_g.shape
    Dd = rep_d.shape[1]
    B = R * B_dev
    g = np.zeros((B, D), dtype=F32)
    f = np.zeros(B, dtype=F32)
    d = np.zeros((B, Dd), dtype=F32)
    c = np.zeros(B, dtype=np.int32)
    _chk(lib().qo_distributed_update(_p(rep_g), _p(rep_f), _p(rep_d), _p(centroids), C.c_int64(K), C.c_int64(D), C.c_int64(Dd),
                                     _p(keys), C.c_int64(R), C.c_int64(B_dev), C.c_int(TASKS[task]), C.c_float(iso_sigma),
                                     C.c_float(line_sigma), *_clip_args(minval, maxval), C.c_int(tie_break == "last"),
                                     _p(g), _p(f), _p(d), _p(c)), "distributed_update")
    return rep_g, rep_f, rep_d, g, f, d, c


def dns_dominated_novelty(f, desc, k: int) -> np.ndarray:
    f, desc = _f(f).reshape(-1), _f(desc)
    out = np.zeros(f.size, dtype=F32)
    _chk(lib().qo_dns_dominated_novelty(_p(f), _p(desc), C.c_int64(f.size), C.c_int64(desc.shape[1]), C.c_int(k), _p(out)), "dns")
    return out


def dns_survivors(meta, P: int) -> np.ndarray:
    meta = _f(meta).reshape(-1)
    out = np.zeros(min(P, meta.size), dtype=np.int32)
    _chk(lib().qo_dns_survivors(_p(meta), C.c_int64(meta.size), C.c_int64(P), _p(out)), "dns_survivors")
    return out


def dns_add(pop_g, pop_f, pop_d, g, f, desc, k: int):
    """DominatedNoveltyRepertoire.add (dns_repertoire.py:94-165): returns (genotypes, fitnesses (P,), descriptors,
    meta (N,), survivors (P,))."""
    pop_g, pop_f, pop_d = _f(pop_g), _f(pop_f).reshape(-1), _f(pop_d)
    cg = np.concatenate([pop_g, _f(g)], axis=0)
    cf = np.concatenate([pop_f, _f(f).reshape(-1)], axis=0)
    cd = np.concatenate([pop_d, _f(desc)], axis=0)
    dn = dns_dominated_novelty(cf, cd, k)
    meta = np.where(cf != -np.inf, dn, F32(-np.inf)).astype(F32)
    surv = dns_survivors(meta, pop_g.shape[0])
    return cg[surv], cf[surv], cd[surv], meta, surv
