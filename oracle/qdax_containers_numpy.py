"""ORACLE (test infrastructure only) -- literal NumPy restatements of the two remaining sibling insertion rules of SURVEY.md
8(f) rank 3, float32 throughout:

  * MOMERepertoire.add          /root/reference/qdax/core/containers/mome_repertoire.py:211-322 (+ _update_masked_pareto_front
                                :72-209, qdax/utils/pareto_front.py:48-95)
  * UnstructuredRepertoire.add  /root/reference/qdax/core/containers/unstructured_repertoire.py:162-337 (+ get_cells_indices
                                :21-66, intra_batch_comp :69-129)

PARITY UNPINNED: no reference test fixes a value for either (mome_test.py:200,378 and aurora_test.py only assert a coverage
threshold / `is not None`), and jax cannot be installed here.  Both are restated line by line from the source, including the
places where the source does something surprising:

  MOME  `cell_fitness - jnp.inf * mask`, `fitness * fitness_mask`, `x * new_mask_indices[0]`, `descriptors * descriptors_mask`
        multiply a float array by a BOOLEAN one.  IEEE arithmetic on the converted mask gives inf * 0 = NaN, i.e. every valid
        entry of a touched cell would become NaN.  `add` always runs inside jax.lax.scan, i.e. compiled, and XLA's algebraic
        simplifier rewrites Mul(x, Convert(pred)) into Select(pred, x, 0) -- which is why the reference's fronts hold finite
        fitnesses in practice.  `ieee_literal=False` (default) follows the compiled behaviour (select), `ieee_literal=True` the
        IEEE product; the CUDA path implements both behind the same flag.
  UNSTRUCTURED  `jnp.where(expand_dims(fitnesses == -inf, -1), full(Dd, inf), descriptors)` with fitnesses of shape (N, 1)
        broadcasts to (N, N, Dd): slot i sees ALL stored descriptors when it is occupied and all-inf when it is empty, and
        `jax.vmap(jnp.linalg.norm)` over it is a Frobenius norm.  Every occupied slot is therefore at the same distance
        F_b = sqrt(sum_j sum_d (x_bd - desc_jd)^2) from offspring b, top_k returns the two lowest-index occupied slots, and the
        l-value tests compare F_b.  Restated as written.  The summation order of F_b inside XLA is unknown; the spec here is
        per slot left to right over d, then slot after slot from +0 (sequential float32).
  Negative scatter / segment indices (empty_indexes padded with -1 when the archive is full) wrap around like NumPy / jnp
  indexing does (-1 = last slot).
"""

from __future__ import annotations

from typing import Tuple

import numpy as np

F32 = np.float32
INF = F32(np.inf)


def _seq_sum_last(a: np.ndarray) -> np.ndarray:
    """sequential float32 sum over the last axis, left to right from the first term"""
    a = np.asarray(a, dtype=F32)
    if a.shape[-1] == 0:
        return np.zeros(a.shape[:-1], F32)
    return np.cumsum(a, axis=-1, dtype=F32)[..., -1]


# ------------------------------------------------------------------------------------------------------------- MOME
def masked_pareto_front(crit: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """compute_masked_pareto_front (pareto_front.py:81-95): crit (n, C) float32, mask (n,) bool (True = no element)."""
    crit = np.asarray(crit, dtype=F32)
    n = crit.shape[0]
    out = np.zeros(n, dtype=bool)
    with np.errstate(invalid="ignore"):
        for i in range(n):
            diff = (crit - crit[i][None, :]).astype(F32)                       # :69
            diff = np.where(mask[:, None], F32(-1.0), diff)                    # :70-73
            dominated = np.any(np.any(diff > 0, axis=-1) & np.all(diff >= 0, axis=-1))     # :74-77
            out[i] = (not dominated) and (not mask[i])                         # :92-95
    return out


def _mul_mask(x: np.ndarray, m: np.ndarray, ieee_literal: bool) -> np.ndarray:
    """float array * boolean mask: IEEE product of the converted mask, or XLA's Select(mask, x, 0)."""
    if ieee_literal:
        with np.errstate(invalid="ignore"):
            return (x * m.astype(F32)).astype(F32)
    return np.where(m, x, F32(0.0)).astype(F32)


def mome_add_one(front_f, front_g, front_d, f, g, d, ieee_literal: bool = False):
    """_add_one (mome_repertoire.py:246-307) on one cell: front_f (L, C), front_g (L, D), front_d (L, Dd); new point f (C,),
    g (D,), d (Dd,).  Returns the new (front_f, front_g, front_d)."""
    L = front_f.shape[0]
    cell_mask = np.any(front_f == -INF, axis=-1)                               # :260
    cat_mask = np.concatenate([cell_mask, np.zeros(1, bool)])                  # :122 (new_mask = zeros)
    cat_f = np.concatenate([front_f, f[None]], axis=0).astype(F32)
    cat_g = np.concatenate([front_g, g[None]], axis=0).astype(F32)
    cat_d = np.concatenate([front_d, d[None]], axis=0).astype(F32)
    front = masked_pareto_front(cat_f, cat_mask)                               # :143-145
    idx = np.arange(L + 1) * front + (~front) * L                              # :148-151  (batch_size + L - 1 = L)
    idx = np.sort(idx)                                                         # :152
    nf, ng, nd = cat_f[idx], cat_g[idx], cat_d[idx]                            # :155-162
    num = int(front.sum())                                                     # :165
    new_mask = (num - np.arange(L + 1)) > 0                                    # :166-173
    nf = _mul_mask(nf, np.repeat(new_mask[:, None], nf.shape[1], axis=1), ieee_literal)[:L]          # :175-181
    ng = _mul_mask(ng, np.full(ng.shape, new_mask[0]), ieee_literal)[:L]       # :183-188 (scalar new_mask_indices[0])
    nd = _mul_mask(nd, np.repeat(new_mask[:, None], nd.shape[1], axis=1), ieee_literal)[:L]          # :190-194
    out_mask = ~new_mask[:L]                                                   # :203
    with np.errstate(invalid="ignore"):
        if ieee_literal:
            nf = (nf - (INF * out_mask[:, None].astype(F32))).astype(F32)      # :288: inf * 0 = NaN for the valid entries
        else:
            nf = (nf - np.where(out_mask[:, None], INF, F32(0.0))).astype(F32)
    return nf, ng, nd


def mome_add(rep_f, rep_g, rep_d, centroids, genotypes, descriptors, fitnesses, cells, ieee_literal: bool = False):
    """MOMERepertoire.add (:211-322): rep_f (K, L, C), rep_g (K, L, D), rep_d (K, L, Dd); the batch is scanned in order, each
    offspring updating the Pareto front of its cell.  `cells` = get_cells_indices(descriptors, centroids)."""
    rep_f, rep_g, rep_d = rep_f.astype(F32).copy(), rep_g.astype(F32).copy(), rep_d.astype(F32).copy()
    for b in range(genotypes.shape[0]):
        c = int(cells[b])
        rep_f[c], rep_g[c], rep_d[c] = mome_add_one(rep_f[c], rep_g[c], rep_d[c], fitnesses[b].astype(F32), genotypes[b].astype(F32),
                                                    descriptors[b].astype(F32), ieee_literal)
    return rep_f, rep_g, rep_d


# ------------------------------------------------------------------------------------------------------------- unstructured
def unstructured_frobenius(batch_desc: np.ndarray, rep_desc: np.ndarray) -> np.ndarray:
    """F_b of the header: per slot sum over d left to right, then slot after slot from +0, then sqrt."""
    x = np.asarray(batch_desc, dtype=F32)
    with np.errstate(invalid="ignore", over="ignore"):
        diff = (x[:, None, :] - np.asarray(rep_desc, dtype=F32)[None, :, :]).astype(F32)
        s = _seq_sum_last((diff * diff).astype(F32))                           # (B, N): per-slot squared distance
        tot = np.cumsum(np.concatenate([np.zeros((x.shape[0], 1), F32), s], axis=1), axis=1, dtype=F32)[:, -1]
        return np.sqrt(tot).astype(F32)


def _top2_neg(dist: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """jax.lax.top_k(-dist, 2) -> (indices, distances); equal values: lower index first; NaN never reaches here (see caller)."""
    order = np.argsort(dist, kind="stable")[:2]
    return order, dist[order]


def unstructured_intra_batch_keep(desc: np.ndarray, fit: np.ndarray, l_value: float) -> np.ndarray:
    """vmap(intra_batch_comp) (:69-129, :272-283) over the (re-ordered) batch: keep[i]."""
    desc = np.asarray(desc, dtype=F32)
    B = desc.shape[0]
    fit = np.asarray(fit, dtype=F32).reshape(B)
    l = F32(l_value)
    with np.errstate(invalid="ignore", over="ignore"):
        ev = np.where(np.isinf(fit), F32(np.nan), fit).astype(F32)             # :83-85
        if np.all(np.isnan(ev)):
            add = F32(0.0)                                                     # nanmax == nanmin is NaN == NaN: False
        else:
            add = F32(1.0) if np.nanmax(ev) == np.nanmin(ev) else F32(0.0)     # :89-91
        if B > 1:                                                              # jnp.linspace(0, add, B) in float32
            step = (np.arange(B - 1, dtype=F32) / F32(B - 1)).astype(F32)
            lin = np.concatenate([(F32(0.0) * (F32(1.0) - step) + add * step).astype(F32), np.array([add], F32)])
        else:
            lin = np.zeros(1, F32)
        ev = np.where(np.isnan(ev), -INF, ev).astype(F32)                      # :95-97
        ev = (ev + lin).astype(F32)                                            # :99
        keep = np.zeros(B, dtype=bool)
        for i in range(B):
            x = desc[i]
            not_existent = bool(np.isnan(x).any())                             # :78
            xi = np.where(np.isnan(x), INF, x).astype(F32)                     # :81
            diff = (xi[None, :] - desc).astype(F32)                            # :102-104 (normed_all is NOT NaN-filtered)
            dist = np.sqrt(_seq_sum_last((diff * diff).astype(F32))).astype(F32)
            close = dist < l                                                   # :109
            close[i] = False                                                   # :112
            discard = bool(np.any((ev > ev[i]) & close)) or not_existent       # :113-125
            keep[i] = not discard
    return keep


def unstructured_add(rep_g, rep_f, rep_d, l_value: float, genotypes, descriptors, fitnesses, tie_break: str = "first"):
    """UnstructuredRepertoire.add (:162-337).  rep_g (N, D), rep_f (N,) [the reference's (N, 1)], rep_d (N, Dd).  Duplicate
    scatter targets resolve to the first / last offspring (in the re-ordered batch) per `tie_break`."""
    rep_g, rep_f, rep_d = rep_g.astype(F32).copy(), np.asarray(rep_f, dtype=F32).reshape(-1).copy(), rep_d.astype(F32).copy()
    N = rep_f.shape[0]
    g, d, f = np.asarray(genotypes, dtype=F32), np.asarray(descriptors, dtype=F32), np.asarray(fitnesses, dtype=F32).reshape(-1)
    B = g.shape[0]
    l = F32(l_value)
    empty = rep_f == -INF                                                      # :188-194
    occ = np.nonzero(~empty)[0]
    F = unstructured_frobenius(d, rep_d)                                       # distance of offspring b to EVERY occupied slot
    with np.errstate(invalid="ignore"):
        dist0 = np.where(len(occ) >= 1, F, INF).astype(F32)                    # :196-207: top_k(-distances, 2)
        dist1 = np.where(len(occ) >= 2, F, INF).astype(F32)
        idx0 = np.full(B, occ[0] if len(occ) else 0, dtype=np.int64)
        # (F_b = inf or NaN: every slot ties / compares false -- the index is irrelevant because dist0 <= l is false)
        not_novel = dist1 <= l                                                 # :211-213
        empty_idx = np.nonzero(np.isinf(rep_f))[0][:B]                         # :223-229 (isinf: +inf counts too)
        empty_idx = np.concatenate([empty_idx, np.full(B - len(empty_idx), -1)]).astype(np.int64)
        near = dist0 <= l
        idx = np.where(near, idx0, -1)                                         # :230-234
        order = np.argsort(idx, kind="stable")                                 # :238-240: top_k(-idx, B)[1]
        idx = np.where(near[order], idx[order], empty_idx)                     # :241-247
    d, g, f, not_novel = d[order], g[order], f[order], not_novel[order]        # :252-263
    keep = unstructured_intra_batch_keep(d, f, l) & ~not_novel                 # :266-285
    idx = np.where(idx < 0, idx + N, idx)                                      # negative indices wrap
    best = np.full(N, -INF, F32)                                               # :288-292 segment_max (NaN-propagating)
    for b in range(B):
        c = idx[b]
        if np.isnan(f[b]) or np.isnan(best[c]):
            best[c] = F32(np.nan)
        elif f[b] > best[c]:
            best[c] = f[b]
    with np.errstate(invalid="ignore"):
        fm = np.where(f == best[idx], f, -INF).astype(F32)                     # :294-299
        cond = (fm > rep_f[idx]) & keep                                        # :302-306
    seq = range(B - 1, -1, -1) if tie_break == "first" else range(B)           # duplicate .at[].set targets
    for b in seq:
        if cond[b]:
            c = idx[b]
            rep_g[c], rep_f[c], rep_d[c] = g[b], fm[b], d[b]                   # :316-330
    return rep_g, rep_f, rep_d
