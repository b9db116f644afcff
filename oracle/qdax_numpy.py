"""ORACLE (test infrastructure only -- never imported by the product path).

NumPy restatement of the MAP-Elites generation step of QDax 0.5.1 (reference at
/root/reference, read-only).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; qdax_b200/ never does.

Every function cites the reference file:line it follows.  Library transcendental
functions (np.log1p, np.sin, np.cos) are used here on purpose: this file is the
*independent, literal* restatement.  Its exact-arithmetic twin is oracle/qdx_oracle.c
(same algorithm with every rounding step spelled out, multi-threaded), which is what
the bit-exact GPU parity tests compare against; tests/test_oracle_cross.py checks the
two against each other (integers bit-exact, floats <= 1e-6).

Parity status:
  PINNED by the reference's own tests and reproduced in tests/test_oracle_golden.py:
    compute_euclidean_centroids((2,2))           tests/core_test/containers_test/mapelites_repertoire_test.py:21-36
    MapElitesRepertoire.init + add (2 offspring) same file :39-81
    arm descriptors of 7 genotypes               tests/tasks_test/arm_test.py:123-163
  PARITY UNPINNED (no reference test fixes a value and no jaxlib is installable here):
    the PRNG stream, UniformSelector indices, isoline/polynomial outputs, rastrigin and
    sphere values, the tie-break inside add, default_qd_metrics values, all of DNS.

Canonical choices where the reference leaves behaviour to XLA (all documented in DESIGN.md):
  * reductions over the genotype/descriptor axis are sequential, left to right, float32;
  * no FMA contraction in elementwise expressions;
  * duplicate-index scatter in `add` resolves to the FIRST offspring index (tie_break="first"),
    "last" is available;
  * `choice(p)` uses the sequential float32 cumsum.
"""

from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import jax_prng as jr

F32 = np.float32
NEG_INF = F32(-np.inf)
TWO_PI = F32(2 * np.pi)  # python float 6.283185307179586 weakly typed -> f32
PI = F32(np.pi)


# --------------------------------------------------------------------------------------
# helpers: canonical sequential float32 reductions over the last axis
# --------------------------------------------------------------------------------------
def seq_sum(a: np.ndarray) -> np.ndarray:
    """Left-to-right float32 sum over the last axis (vectorised over leading axes)."""
    a = np.asarray(a, dtype=F32)
    acc = a[..., 0].copy()
    for d in range(1, a.shape[-1]):
        acc = (acc + a[..., d]).astype(F32)
    return acc


def seq_cumsum(a: np.ndarray) -> np.ndarray:
    return np.cumsum(np.asarray(a, dtype=F32), axis=-1, dtype=F32)


# --------------------------------------------------------------------------------------
# centroids -- qdax/core/containers/mapelites_repertoire.py:75-108
# --------------------------------------------------------------------------------------
def _linspace_f32(start: float, stop: float, num: int) -> np.ndarray:
    """jnp.linspace(start, stop, num) in float32: start*(1-step) + stop*step with
    step = iota(div)/div, endpoint appended exactly."""
    start, stop = F32(start), F32(stop)
    if num == 1:
        return np.array([start], dtype=F32)
    div = num - 1
    step = (np.arange(div, dtype=F32) / F32(div)).astype(F32)
    out = ((start * (F32(1) - step).astype(F32)).astype(F32) + (stop * step).astype(F32)).astype(F32)
    return np.concatenate([out, np.array([stop], dtype=F32)])


def compute_euclidean_centroids(grid_shape: Sequence[int], minval, maxval) -> np.ndarray:
    """mapelites_repertoire.py:75-108.  meshgrid default indexing='xy'."""
    lin = []
    for n in grid_shape:
        offset = 1 / (2 * n)
        lin.append(_linspace_f32(offset, 1.0 - offset, n))
    meshes = np.meshgrid(*lin, sparse=False)  # indexing='xy' like jnp.meshgrid
    cent = np.stack([np.ravel(m) for m in meshes], axis=-1).astype(F32)
    minval = np.asarray(minval, dtype=F32)
    maxval = np.asarray(maxval, dtype=F32)
    return ((cent * (maxval - minval).astype(F32)).astype(F32) + minval).astype(F32)


# --------------------------------------------------------------------------------------
# cell assignment -- mapelites_repertoire.py:111-137
# --------------------------------------------------------------------------------------
def get_cells_indices(descriptors: np.ndarray, centroids: np.ndarray, chunk: int = 4096) -> np.ndarray:
    """argmin_k sum_d (desc_d - c_kd)^2 ; first minimum; NaN counts as minimal (first NaN)."""
    descriptors = np.asarray(descriptors, dtype=F32)
    centroids = np.asarray(centroids, dtype=F32)
    B, Dd = descriptors.shape
    out = np.empty((B,), dtype=np.int32)
    with np.errstate(invalid="ignore", over="ignore"):
        for s in range(0, B, chunk):
            d = descriptors[s : s + chunk]
            acc = None
            for j in range(Dd):
                diff = (d[:, None, j] - centroids[None, :, j]).astype(F32)
                sq = (diff * diff).astype(F32)
                acc = sq if acc is None else (acc + sq).astype(F32)
            out[s : s + chunk] = np.argmin(acc, axis=1)
    return out


# --------------------------------------------------------------------------------------
# repertoire -- mapelites_repertoire.py:140-388, ga_repertoire.py:34-39
# --------------------------------------------------------------------------------------
@dataclass
class Repertoire:
    genotypes: np.ndarray  # (K, D)
    fitnesses: np.ndarray  # (K, 1)
    descriptors: np.ndarray  # (K, Dd)
    centroids: np.ndarray  # (K, Dd)

    def copy(self) -> "Repertoire":
        return Repertoire(self.genotypes.copy(), self.fitnesses.copy(), self.descriptors.copy(), self.centroids)


def repertoire_init_default(genotype_dim: int, centroids: np.ndarray) -> Repertoire:
    """mapelites_repertoire.py:328-388: fitness -inf (K,1), genotypes 0, descriptors 0."""
    K = centroids.shape[0]
    return Repertoire(
        genotypes=np.zeros((K, genotype_dim), dtype=F32),
        fitnesses=np.full((K, 1), NEG_INF, dtype=F32),
        descriptors=np.zeros_like(np.asarray(centroids, dtype=F32)),
        centroids=np.asarray(centroids, dtype=F32),
    )


def segment_max(values: np.ndarray, seg: np.ndarray, num_segments: int) -> np.ndarray:
    """jax.ops.segment_max: identity -inf, NaN-propagating (XLA max)."""
    out = np.full((num_segments,), NEG_INF, dtype=F32)
    nan = np.isnan(values)
    np.maximum.at(out, seg[~nan], values[~nan])
    out[np.unique(seg[nan])] = F32(np.nan)
    return out


def repertoire_add(
    rep: Repertoire,
    genotypes: np.ndarray,
    descriptors: np.ndarray,
    fitnesses: np.ndarray,
    tie_break: str = "first",
    cells: Optional[np.ndarray] = None,
) -> Tuple[Repertoire, np.ndarray, np.ndarray]:
    """mapelites_repertoire.py:173-266.  Returns (new repertoire, cells (B,), scatter index (B,)
    where K means "dropped").  Duplicate-index scatter order is unspecified in the reference;
    `tie_break` makes it explicit."""
    genotypes = np.asarray(genotypes, dtype=F32)
    descriptors = np.asarray(descriptors, dtype=F32)
    f = np.asarray(fitnesses, dtype=F32).reshape(-1)
    K = rep.centroids.shape[0]
    if cells is None:
        cells = get_cells_indices(descriptors, rep.centroids)  # :202
    cells = cells.astype(np.int64)
    best = segment_max(f, cells, K)  # :211-215
    with np.errstate(invalid="ignore"):
        fm = np.where(f == best[cells], f, NEG_INF).astype(F32)  # :217-222
        cond = fm > rep.fitnesses[cells, 0]  # :225-226 (strict)
    idx = np.where(cond, cells, K)  # :229-231
    new = rep.copy()
    order = np.nonzero(cond)[0]
    if tie_break == "first":
        order = order[::-1]  # last write wins => write the first index last
    elif tie_break != "last":
        raise ValueError(tie_break)
    for i in order:  # .at[idx].set, out-of-bounds rows dropped  :234-257
        c = idx[i]
        new.genotypes[c] = genotypes[i]
        new.fitnesses[c, 0] = fm[i]
        new.descriptors[c] = descriptors[i]
    return new, cells.astype(np.int32), idx.astype(np.int32)


def repertoire_init(genotypes, fitnesses, descriptors, centroids, tie_break="first") -> Repertoire:
    """mapelites_repertoire.py:268-326."""
    rep = repertoire_init_default(np.asarray(genotypes).shape[1], centroids)
    return repertoire_add(rep, genotypes, descriptors, fitnesses, tie_break)[0]


# --------------------------------------------------------------------------------------
# metrics -- qdax/utils/metrics.py:74-98
# --------------------------------------------------------------------------------------
# ------------------------------------------------------------------ CVT centroids by Lloyd iterations
def lloyd_cvt_centroids(x: np.ndarray, num_centroids: int, num_iterations: int = 100):
    """The GPU backend of compute_cvt_centroids (mapelites_repertoire.py:30-72 runs scikit-learn KMeans instead; this rule is
    a declared replacement, see qdax_b200 lloyd_cvt_centroids): centroids start as the first K samples; assignment =
    get_cells_indices (first minimum); update = per-cluster mean of the coordinates quantised to 32 fractional bits, summed
    exactly in integers, (double)sum / count * 2^-32 -> float32; empty clusters keep their centroid; stop at a fixed point."""
    x = np.asarray(x, dtype=F32)
    N, Dd = x.shape
    K = int(num_centroids)
    cent = x[:K].copy()
    q = np.minimum(np.floor(np.maximum(x, 0).astype(np.float64) * 4294967296.0), 4294967295.0).astype(np.uint64)
    prev = None
    it = 0
    for it in range(1, num_iterations + 1):
        cells = get_cells_indices(x, cent)
        if prev is not None and np.array_equal(prev, cells):
            break
        acc = np.zeros((K, Dd), dtype=np.uint64)
        np.add.at(acc, cells, q)
        count = np.bincount(cells, minlength=K)
        mean = ((acc.astype(np.float64) / np.maximum(count, 1)[:, None]) * (1.0 / 4294967296.0)).astype(F32)
        cent = np.where((count > 0)[:, None], mean, cent).astype(F32)
        prev = cells
    return cent, it


# ------------------------------------------------------------------ MELS (qdax/core/containers/mels_repertoire.py)
def mels_dispersion(descriptors: np.ndarray) -> np.float32:
    """_dispersion :26-48: mean of the unique pairwise distances (float32; sums sequential, row-major over i < j)."""
    d = np.asarray(descriptors, dtype=F32)
    S = d.shape[0]
    total = F32(0.0)
    for i in range(S):
        for j in range(i + 1, S):
            diff = (d[i] - d[j]).astype(F32)
            total = F32(total + F32(np.sqrt(seq_sum((diff * diff).astype(F32)[None, :])[0])))
    return F32(total / F32(S * (S - 1) / 2.0))


def mels_mode(x: np.ndarray) -> int:
    """_mode :51-57: jnp.unique (sorted) + argmax of the counts (first maximum) = smallest most frequent value."""
    vals, counts = np.unique(np.asarray(x), return_counts=True)
    return int(vals[np.argmax(counts)])


def mels_add(rep: "Repertoire", spreads: np.ndarray, genotypes, descriptors, fitnesses, tie_break: str = "first"):
    """MELSRepertoire.add :89-230 on (B, S, Dd) descriptors and (B, S) fitnesses.  Returns (new repertoire, new spreads).
    Every candidate with mean fitness > occupant and spread <= occupant's spread scatters to its cell; a collision is
    resolved to the first / last offspring index (the reference leaves it to the scatter, :103-110)."""
    g = np.asarray(genotypes, dtype=F32)
    d = np.asarray(descriptors, dtype=F32)
    f = np.asarray(fitnesses, dtype=F32)
    B, S = f.shape
    cells_all = get_cells_indices(d.reshape(B * S, -1), rep.centroids).reshape(B, S)  # :143-145
    cell = np.array([mels_mode(cells_all[b]) for b in range(B)], dtype=np.int64)  # :148
    spread = np.zeros(B, dtype=F32) if S == 1 else np.array([mels_dispersion(d[b]) for b in range(B)], dtype=F32)  # :152-158
    fmean = (seq_sum(f) / F32(S)).astype(F32)  # :169
    out = rep.copy()
    new_spreads = np.array(spreads, dtype=F32, copy=True)
    cond = (fmean > rep.fitnesses[cell, 0]) & (spread <= new_spreads[cell])  # :181-187
    order = range(B - 1, -1, -1) if tie_break == "first" else range(B)  # the last write stays
    for b in order:
        if cond[b]:
            c = cell[b]
            out.genotypes[c] = g[b]
            out.fitnesses[c, 0] = fmean[b]
            out.descriptors[c] = rep.centroids[c]  # :162-164
            new_spreads[c] = spread[b]
    return out, new_spreads


def default_qd_metrics(rep: Repertoire, qd_offset: float = 0.0) -> Dict[str, np.float32]:
    empty = rep.fitnesses == NEG_INF
    with np.errstate(invalid="ignore"):
        qd = F32(np.sum(np.where(empty, F32(0), rep.fitnesses), dtype=np.float64))
    filled = F32(np.sum(~empty))
    qd = F32(qd + F32(qd_offset) * filled)
    coverage = F32(F32(100) * F32(filled / F32(empty.size)))
    return {"qd_score": qd, "max_fitness": F32(np.max(rep.fitnesses)), "coverage": coverage}


# --------------------------------------------------------------------------------------
# selection -- repertoire_selectors/uniform_selector.py:22-62
# --------------------------------------------------------------------------------------
def uniform_select_indices(fitnesses: np.ndarray, key: np.ndarray, num_samples: int) -> np.ndarray:
    empty = np.any(np.asarray(fitnesses).reshape(len(fitnesses), -1) == NEG_INF, axis=-1)  # :44
    occ = (F32(1.0) - empty.astype(F32)).astype(F32)
    p = (occ / F32(np.sum(occ, dtype=np.float64))).astype(F32)  # :45  (sum of M ones is exact)
    sub = jr.split(key)[1]  # :48
    return jr.choice_p_replace(sub, p, num_samples)  # :49-55


def uniform_select_indices_without_replacement(fitnesses: np.ndarray, key: np.ndarray, num_samples: int) -> np.ndarray:
    """UniformSelector(select_with_replacement=False) (uniform_selector.py:19-20, :49-55): jax.random.choice(subkey, arange(K),
    (n,), p=p, replace=False) is the Gumbel top-k trick [recalled from jax/_src/random.py]: g = gumbel(key, (K,)) + log(p),
    gumbel = -log(-log(uniform(key, (K,), minval=tiny, maxval=1))); indices = top_k(g, n) (equal values: lower index first).
    Empty cells have log(0) = -inf and come last.  PARITY UNPINNED (XLA's log, and jax's gumbel sampling mode)."""
    f = np.asarray(fitnesses)
    empty = np.any(f.reshape(len(f), -1) == NEG_INF, axis=-1)  # :44
    occ = (F32(1.0) - empty.astype(F32)).astype(F32)
    p = (occ / F32(np.sum(occ, dtype=np.float64))).astype(F32)  # :45
    sub = jr.split(key)[1]  # :48
    K = p.shape[0]
    if num_samples > K:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    u = jr.uniform(sub, (K,), np.finfo(F32).tiny, 1.0)
    with np.errstate(divide="ignore"):
        g = (-np.log((-np.log(u)).astype(F32))).astype(F32)
        g = (g + np.log(p).astype(F32)).astype(F32)
    return np.argsort(-g, kind="stable")[:num_samples].astype(np.int32)


# --------------------------------------------------------------------------------------
# variation -- emitters/mutation_operators.py
# --------------------------------------------------------------------------------------
def isoline_variation(x1, x2, key, iso_sigma, line_sigma, minval=None, maxval=None) -> np.ndarray:
    """mutation_operators.py:175-226 (single-leaf genotype)."""
    x1 = np.asarray(x1, dtype=F32)
    x2 = np.asarray(x2, dtype=F32)
    B = x1.shape[0]
    ks = jr.split(key)  # :205
    key, k_line = ks[0], ks[1]
    line = (jr.normal(k_line, (B,)) * F32(line_sigma)).astype(F32)  # :207
    k_leaf = jr.split(key, 1)[0]  # :220 (one leaf)
    iso = (jr.normal(k_leaf, x1.shape) * F32(iso_sigma)).astype(F32)  # :210
    x = ((x1 + iso).astype(F32) + ((x2 - x1).astype(F32) * line[:, None]).astype(F32)).astype(F32)  # :211
    if minval is not None or maxval is not None:  # :214-215
        if minval is not None:
            x = np.maximum(x, F32(minval))
        if maxval is not None:
            x = np.minimum(x, F32(maxval))
    return x.astype(F32)


def isoline_variation_tree(x1_leaves, x2_leaves, key, iso_sigma, line_sigma, minval=None, maxval=None):
    """mutation_operators.py:175-226 on a pytree genotype given as its list of leaves (jax.tree.leaves order), each of
    shape (B, ...): shared line noise (:205-207), keys = split(key, nb_leaves) (:220), one normal(key_l, leaf.shape) per
    leaf (:210)."""
    B = np.asarray(x1_leaves[0]).shape[0]
    ks = jr.split(key)  # :205
    key, k_line = ks[0], ks[1]
    line = (jr.normal(k_line, (B,)) * F32(line_sigma)).astype(F32)  # :207
    keys = jr.split(key, len(x1_leaves))  # :220
    out = []
    for a, b, k in zip(x1_leaves, x2_leaves, keys):
        a = np.asarray(a, dtype=F32)
        b = np.asarray(b, dtype=F32)
        iso = (jr.normal(k, a.shape) * F32(iso_sigma)).astype(F32)  # :210
        ln = line.reshape((B,) + (1,) * (a.ndim - 1))  # jax.vmap(jnp.multiply)((x2 - x1), line_noise)
        x = ((a + iso).astype(F32) + ((b - a).astype(F32) * ln).astype(F32)).astype(F32)  # :211
        if minval is not None:
            x = np.maximum(x, F32(minval))
        if maxval is not None:
            x = np.minimum(x, F32(maxval))
        out.append(x.astype(F32))
    return out


def _polynomial_mutation_row(x, key, proportion_to_mutate, eta, minval, maxval) -> np.ndarray:
    """mutation_operators.py:12-78 for one genotype."""
    x = np.asarray(x, dtype=F32).copy()
    n = x.shape[0]
    m = int(proportion_to_mutate * n)
    ks = jr.split(key)
    key, sub = ks[0], ks[1]
    pos = jr.choice_no_replace(sub, n, m)  # :42-45
    lo, hi = F32(minval), F32(maxval)
    rng = F32(hi - lo)
    mx = x[pos]
    d1 = ((mx - lo) / rng).astype(F32)
    d2 = ((hi - mx) / rng).astype(F32)
    mutpow = F32(1.0 / (1.0 + eta))
    ep1 = F32(1.0 + eta)
    ks = jr.split(key)
    key, sub = ks[0], ks[1]
    r = jr.uniform(sub, (m,))
    with np.errstate(invalid="ignore"):
        v1 = (F32(2) * r + (np.power(d1, ep1).astype(F32) * (F32(1) - F32(2) * r).astype(F32)).astype(F32)).astype(F32)
        v2 = (
            F32(2) * (F32(1) - r).astype(F32)
            + (F32(2) * (np.power(d2, ep1).astype(F32) * (r - F32(0.5)).astype(F32)).astype(F32)).astype(F32)
        ).astype(F32)
        v1 = (np.power(v1, mutpow).astype(F32) - F32(1)).astype(F32)
        v2 = (F32(1) - np.power(v2, mutpow).astype(F32)).astype(F32)
    dq = np.where(r < F32(0.5), v1, v2).astype(F32)
    x[pos] = (mx + (dq * rng).astype(F32)).astype(F32)
    return np.minimum(np.maximum(x, lo), hi).astype(F32)


def polynomial_mutation(x, key, proportion_to_mutate, eta, minval, maxval) -> np.ndarray:
    """mutation_operators.py:81-117."""
    x = np.asarray(x, dtype=F32)
    keys = jr.split(key, x.shape[0])  # :107
    return np.stack(
        [_polynomial_mutation_row(x[i], keys[i], proportion_to_mutate, eta, minval, maxval) for i in range(x.shape[0])]
    )


def polynomial_crossover(x1, x2, key, proportion_var_to_change) -> np.ndarray:
    """mutation_operators.py:120-172."""
    x1 = np.asarray(x1, dtype=F32)
    x2 = np.asarray(x2, dtype=F32)
    B, D = x1.shape
    n = int(proportion_var_to_change * D)
    keys = jr.split(key, B)
    out = x1.copy()
    for i in range(B):
        sel = jr.choice_uniform_replace(keys[i], D, n)
        out[i, sel] = x2[i, sel]
    return out


# --------------------------------------------------------------------------------------
# emitter -- emitters/standard_emitters.py:27-82 (variation_percentage == 1.0 path + mixed)
# --------------------------------------------------------------------------------------
def mixing_emit_isoline(
    rep: Repertoire, key, batch_size, iso_sigma, line_sigma, minval=None, maxval=None
) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """variation_percentage = 1.0: returns (offspring, parent1 idx, parent2 idx)."""
    ks = jr.split(key, 3)  # :55
    i1 = uniform_select_indices(rep.fitnesses, ks[0], batch_size)  # :56
    i2 = uniform_select_indices(rep.fitnesses, ks[1], batch_size)  # :59
    x = isoline_variation(rep.genotypes[i1], rep.genotypes[i2], ks[2], iso_sigma, line_sigma, minval, maxval)
    return x, i1, i2


# --------------------------------------------------------------------------------------
# tasks -- qdax/tasks/arm.py:9-50, qdax/tasks/standard_functions.py:9-48
# --------------------------------------------------------------------------------------
def arm_scoring_function(params: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    p = np.asarray(params, dtype=F32)
    D = p.shape[-1]
    x = np.minimum(np.maximum(p, F32(0)), F32(1))  # :27
    mean = (seq_sum(x) / F32(D)).astype(F32)
    dev = (x - mean[..., None]).astype(F32)
    var = (seq_sum((dev * dev).astype(F32)) / F32(D)).astype(F32)
    f = np.sqrt(var).astype(F32)  # :30
    ang = ((TWO_PI * x).astype(F32) - PI).astype(F32)
    cum = seq_cumsum(ang)  # :33
    xs = ((seq_sum(np.cos(cum).astype(F32)) / F32(2 * D)).astype(F32) + F32(0.5)).astype(F32)  # :34
    ys = ((seq_sum(np.sin(cum).astype(F32)) / F32(2 * D)).astype(F32) + F32(0.5)).astype(F32)  # :35
    return (-f).astype(F32), np.stack([xs, ys], axis=-1).astype(F32)


def noisy_arm_scoring_function(params: np.ndarray, key, fit_variance: float, desc_variance: float,
                               params_variance: float) -> Tuple[np.ndarray, np.ndarray]:
    """qdax/tasks/arm.py:53-81: key, f_sub, d_sub, p_sub = split(key, 4); params += normal(p_sub) * params_variance; arm;
    fitnesses += normal(f_sub) * fit_variance; descriptors += normal(d_sub) * desc_variance."""
    from . import jax_prng as jr

    p = np.asarray(params, dtype=F32)
    ks = jr.split(np.asarray(key, dtype=np.uint32), 4)                                   # :65
    f_sub, d_sub, p_sub = ks[1], ks[2], ks[3]
    noisy = (p + (jr.normal(p_sub, p.shape) * F32(params_variance)).astype(F32)).astype(F32)   # :68
    f, d = arm_scoring_function(noisy)                                                   # :71
    f = (f + (jr.normal(f_sub, f.shape) * F32(fit_variance)).astype(F32)).astype(F32)    # :74-76
    d = (d + (jr.normal(d_sub, d.shape) * F32(desc_variance)).astype(F32)).astype(F32)   # :77-80
    return f, d


def rastrigin_scoring_function(params: np.ndarray, desc_dim: int = 2) -> Tuple[np.ndarray, np.ndarray]:
    p = np.asarray(params, dtype=F32)
    D = p.shape[-1]
    x = ((p * F32(10)).astype(F32) - F32(5)).astype(F32)  # :13
    term = ((x * x).astype(F32) - (F32(10) * np.cos((TWO_PI * x).astype(F32)).astype(F32)).astype(F32)).astype(F32)
    f = (F32(10.0 * D) + seq_sum(term)).astype(F32)  # :14
    return (-f).astype(F32), p[..., :desc_dim].copy()


def sphere_scoring_function(params: np.ndarray, desc_dim: int = 2) -> Tuple[np.ndarray, np.ndarray]:
    """standard_functions.py:18-24.  desc_dim != 2 is the declared C4 extension (desc = p[:desc_dim])."""
    p = np.asarray(params, dtype=F32)
    x = ((p * F32(10)).astype(F32) - F32(5)).astype(F32)
    f = seq_sum((x * x).astype(F32))
    return (-f).astype(F32), p[..., :desc_dim].copy()


SCORING = {"arm": arm_scoring_function, "rastrigin": rastrigin_scoring_function, "sphere": sphere_scoring_function}


# --------------------------------------------------------------------------------------
# driver -- qdax/core/map_elites.py:57-286 (key chain is observable behaviour)
# --------------------------------------------------------------------------------------
@dataclass
class EmitterConfig:
    batch_size: int
    iso_sigma: float = 0.05
    line_sigma: float = 0.1
    minval: Optional[float] = 0.0
    maxval: Optional[float] = 1.0


def map_elites_init(genotypes, centroids, key, task="arm", tie_break="first"):
    """map_elites.py:57-146.  Returns (repertoire, metrics)."""
    ks = jr.split(key)  # :81
    f, d = SCORING[task](genotypes)
    rep = repertoire_init(genotypes, f, d, centroids, tie_break)
    # :133 key, subkey = split(key); emitter.init -> None for MixingEmitter
    return rep, default_qd_metrics(rep)


def map_elites_update(rep: Repertoire, key, cfg: EmitterConfig, task="arm", tie_break="first", debug=None):
    """map_elites.py:148-195: update -> ask -> emit ; scoring ; tell."""
    ks = jr.split(key)  # :177
    key, ask_key = ks[0], ks[1]
    emit_key = jr.split(ask_key)[1]  # :241
    x, i1, i2 = mixing_emit_isoline(rep, emit_key, cfg.batch_size, cfg.iso_sigma, cfg.line_sigma, cfg.minval, cfg.maxval)
    # :181 key, subkey = split(key): scoring key ignored by arm / rastrigin / sphere
    f, d = SCORING[task](x)
    new, cells, idx = repertoire_add(rep, x, d, f, tie_break)
    if debug is not None:
        debug.update(genotypes=x, fitnesses=f, descriptors=d, cells=cells, scatter_idx=idx, parents1=i1, parents2=i2)
    return new, default_qd_metrics(new)


def map_elites_scan(rep: Repertoire, key, num_iterations, cfg: EmitterConfig, task="arm", tie_break="first"):
    """lax.scan(map_elites.scan_update, ...) -- map_elites.py:197-225."""
    metrics = []
    for _ in range(num_iterations):
        ks = jr.split(key)  # :214
        key, sub = ks[0], ks[1]
        rep, m = map_elites_update(rep, sub, cfg, task, tie_break)
        metrics.append(m)
    return rep, key, metrics


def distributed_update(rep: Repertoire, keys: Sequence[np.ndarray], cfg: EmitterConfig, task="arm", tie_break="first"):
    """distributed_map_elites.py:92-161 simulated for R = len(keys) devices; cfg.batch_size is per device.
    Global offspring index = rank * B_dev + i (all_gather + concatenate(axis=0), :134-141)."""
    gs, fs, ds = [], [], []
    for key in keys:
        ks = jr.split(key)  # :124
        emit_key = ks[1]
        x, _, _ = mixing_emit_isoline(rep, emit_key, cfg.batch_size, cfg.iso_sigma, cfg.line_sigma, cfg.minval, cfg.maxval)
        f, d = SCORING[task](x)
        gs.append(x), fs.append(f), ds.append(d)
    new, cells, idx = repertoire_add(rep, np.concatenate(gs), np.concatenate(ds), np.concatenate(fs), tie_break)
    return new, default_qd_metrics(new)


# --------------------------------------------------------------------------------------
# Dominated Novelty Search -- qdax/core/containers/dns_repertoire.py:22-165
# --------------------------------------------------------------------------------------
def dominated_novelty(fitness: np.ndarray, descriptor: np.ndarray, k: int, block: int = 1024) -> np.ndarray:
    """dns_repertoire.py:22-76 (dominated novelty only; plain novelty is discarded by add, :136).
    Row-blocked so that it runs at large N; each row is the dense formulation."""
    f = np.asarray(fitness, dtype=F32)
    desc = np.asarray(descriptor, dtype=F32)
    N, Dd = desc.shape
    valid = f != NEG_INF
    out = np.empty((N,), dtype=F32)
    with np.errstate(invalid="ignore", divide="ignore"):
        for s in range(0, N, block):
            e = min(N, s + block)
            nb = valid[s:e, None] & valid[None, :]
            nb[np.arange(e - s), np.arange(s, e)] = False  # :45
            fitter = (f[s:e, None] <= f[None, :]) & nb  # :48-49
            acc = None
            for j in range(Dd):
                diff = (desc[s:e, None, j] - desc[None, :, j]).astype(F32)
                sq = (diff * diff).astype(F32)
                acc = sq if acc is None else (acc + sq).astype(F32)
            dist = np.sqrt(acc).astype(F32)  # :52
            dist_fit = np.where(fitter, dist, F32(np.inf)).astype(F32)  # :53,:56
            kk = min(k, N)
            # top_k(-x, k): k smallest distances, ties -> lower index first (stable sort)
            idx = np.argsort(dist_fit, axis=1, kind="stable")[:, :kk]
            vals = np.take_along_axis(dist_fit, idx, axis=1)
            mask = np.take_along_axis(fitter, idx, axis=1)
            tot = np.zeros((e - s,), dtype=F32)
            for j in range(kk):  # mean(-values, where=mask): sequential sum in top-k order
                tot = (tot + np.where(mask[:, j], vals[:, j], F32(0))).astype(F32)
            cnt = mask.sum(axis=1).astype(F32)
            out[s:e] = (tot / cnt).astype(F32)  # 0/0 -> NaN  (:70-74)
    return out


@dataclass
class DNSRepertoire:
    genotypes: np.ndarray  # (P, D)
    fitnesses: np.ndarray  # (P, 1)
    descriptors: np.ndarray  # (P, Dd)
    k: int


def dns_init_default(genotype_dim: int, descriptor_dim: int, population_size: int, k: int) -> DNSRepertoire:
    """dns_repertoire.py:214-273: fitness -inf, genotypes 0, descriptors NaN."""
    return DNSRepertoire(
        np.zeros((population_size, genotype_dim), dtype=F32),
        np.full((population_size, 1), NEG_INF, dtype=F32),
        np.full((population_size, descriptor_dim), np.nan, dtype=F32),
        k,
    )


def dns_survivor_order(meta: np.ndarray) -> np.ndarray:
    """argsort(meta)[::-1]: stable ascending with NaN last, reversed (dns_repertoire.py:148)."""
    return np.argsort(meta, kind="stable")[::-1]


def dns_add(rep: DNSRepertoire, genotypes, descriptors, fitnesses) -> Tuple[DNSRepertoire, np.ndarray, np.ndarray]:
    """dns_repertoire.py:94-165.  Returns (repertoire, meta fitness (N,), survivor indices (P,))."""
    g = np.concatenate([rep.genotypes, np.asarray(genotypes, dtype=F32)], axis=0)
    f = np.concatenate([rep.fitnesses, np.asarray(fitnesses, dtype=F32).reshape(-1, 1)], axis=0)
    d = np.concatenate([rep.descriptors, np.asarray(descriptors, dtype=F32)], axis=0)
    dn = dominated_novelty(f[:, 0], d, rep.k)
    valid = f[:, 0] != NEG_INF
    meta = np.where(valid, dn, NEG_INF).astype(F32)  # :144-145
    surv = dns_survivor_order(meta)[: rep.genotypes.shape[0]]  # :148-149
    return DNSRepertoire(g[surv], f[surv], d[surv], rep.k), meta, surv.astype(np.int32)
