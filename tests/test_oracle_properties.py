"""Property tests (hypothesis) on the two CPU oracles -- the literal NumPy restatement and the exact-arithmetic C twin were
written independently; on adversarial random inputs (heavy ties, NaN, +-inf, +-0, duplicates, empty repertoires) they must
take the same decisions, and the insertion rule must have the algebraic properties SURVEY.md Appendix C lists.
Reference: qdax/core/containers/mapelites_repertoire.py:111-137,173-266; dns_repertoire.py:22-165; mels_repertoire.py:26-57."""
import numpy as np
import pytest

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from oracle import qdax_numpy as qn  # noqa: E402

SET = settings(max_examples=120, deadline=None, derandomize=True, database=None)      # the same examples on every run
SPECIAL = np.array([np.nan, -np.inf, np.inf, 0.0, -0.0], dtype=np.float32)


def batch(seed, B, K_side, D, levels, special_frac):
    rng = np.random.default_rng(seed)
    cent = qn.compute_euclidean_centroids((K_side, K_side), 0.0, 1.0)
    K = cent.shape[0]
    g = rng.random((B, D)).astype(np.float32)
    d = np.round(rng.random((B, 2)) * 1.2 - 0.1, 1 + seed % 3).astype(np.float32)        # many descriptors exactly on bisectors
    f = (rng.integers(0, levels, B) / max(levels - 1, 1) - 0.3).astype(np.float32)        # few distinct levels: same-cell ties
    hit = rng.random(B) < special_frac
    f[hit] = SPECIAL[rng.integers(0, len(SPECIAL), int(hit.sum()))]
    rep_f = np.where(rng.random(K) < 0.5, np.round(rng.standard_normal(K), 1), -np.inf).astype(np.float32)
    rep_g = rng.random((K, D)).astype(np.float32)
    rep_d = np.where(np.isinf(rep_f)[:, None], 0.0, cent).astype(np.float32)
    return cent, rep_g, rep_f, rep_d, g, d, f


@SET
@given(seed=st.integers(0, 10**6), B=st.integers(1, 300), K_side=st.integers(1, 9), levels=st.integers(1, 6),
       special=st.sampled_from([0.0, 0.05, 0.5]), tb=st.sampled_from(["first", "last"]))
def test_add_both_oracles_agree(co, seed, B, K_side, levels, special, tb):
    cent, rep_g, rep_f, rep_d, g, d, f = batch(seed, B, K_side, 3, levels, special)
    cells_np = qn.get_cells_indices(d, cent)
    assert np.array_equal(co.cells(d, cent), cells_np)                                    # first minimum, bisector ties included
    rep = qn.Repertoire(rep_g.copy(), rep_f.reshape(-1, 1).copy(), rep_d.copy(), cent)
    new, cells, sidx = qn.repertoire_add(rep, g, d, f, tb)
    G, F, Dn, sidx_c = co.add(rep_g, rep_f, rep_d, g, f, d, cells_np, tb)
    assert np.array_equal(new.genotypes, G) and np.array_equal(new.descriptors, Dn)
    assert np.array_equal(new.fitnesses.ravel().view(np.uint32), F.view(np.uint32))       # bit patterns: -0.0 vs +0.0, NaN never stored
    assert not np.isnan(F).any()
    # monotone: no cell ever gets worse, occupied cells stay occupied
    assert (F >= rep_f).all()


@SET
@given(seed=st.integers(0, 10**6), B=st.integers(1, 200), K_side=st.integers(1, 8))
def test_add_is_idempotent_and_order_free_without_ties(co, seed, B, K_side):
    cent, rep_g, rep_f, rep_d, g, d, _ = batch(seed, B, K_side, 2, 3, 0.0)
    rng = np.random.default_rng(seed + 1)
    f = rng.permutation(B).astype(np.float32)                                             # strictly distinct fitnesses
    cells = co.cells(d, cent)
    G, F, Dn, _ = co.add(rep_g, rep_f, rep_d, g, f, d, cells, "first")
    perm = rng.permutation(B)
    G2, F2, Dn2, _ = co.add(rep_g, rep_f, rep_d, g[perm], f[perm], d[perm], cells[perm], "last")
    assert np.array_equal(G, G2) and np.array_equal(F, F2) and np.array_equal(Dn, Dn2)    # no ties -> no order / tie-break dependence
    occ = F != -np.inf
    G3, F3, Dn3, _ = co.add(G, F, Dn, G[occ], F[occ], Dn[occ], co.cells(Dn[occ], cent), "first")
    assert np.array_equal(G3, G) and np.array_equal(F3, F)                                # re-adding its own contents changes nothing


@SET
@given(seed=st.integers(0, 10**6), K=st.integers(1, 400), occ=st.sampled_from([0.02, 0.5, 1.0]), n=st.integers(1, 500))
def test_uniform_selection_both_oracles_agree(co, seed, K, occ, n):
    from oracle import jax_prng as jr

    rng = np.random.default_rng(seed)
    f = np.where(rng.random(K) < occ, rng.standard_normal(K), -np.inf).astype(np.float32)
    if not np.isfinite(f).any():
        f[rng.integers(0, K)] = 1.0
    idx = co.select_indices(f, jr.key(seed), n)
    assert np.array_equal(idx, qn.uniform_select_indices(f.reshape(-1, 1), jr.key(seed), n))
    assert np.isfinite(f[idx]).all()                                                      # never an empty cell (uniform_selector.py:44-45)


@SET
@given(seed=st.integers(0, 10**6), P=st.integers(2, 120), B=st.integers(1, 40), Dd=st.integers(1, 3), k=st.integers(1, 5))
def test_dns_both_oracles_agree(co, seed, P, B, Dd, k):
    rng = np.random.default_rng(seed)
    pf = np.where(rng.random(P) < 0.8, np.round(rng.standard_normal(P), 1), -np.inf).astype(np.float32)
    pd = np.where((pf == -np.inf)[:, None], np.nan, np.round(rng.random((P, Dd)), 1)).astype(np.float32)
    pg = rng.random((P, 2)).astype(np.float32)
    bf = np.round(rng.standard_normal(B), 1).astype(np.float32)
    bd, bg = np.round(rng.random((B, Dd)), 1).astype(np.float32), rng.random((B, 2)).astype(np.float32)
    rep, meta, surv = qn.dns_add(qn.DNSRepertoire(pg, pf.reshape(-1, 1), pd, k), bg, bd, bf)
    G, F, Dn, meta_c, surv_c = co.dns_add(pg, pf, pd, bg, bf, bd, k)
    assert np.array_equal(surv, surv_c)                                                   # NaN first, descending, higher index first among equals
    assert np.allclose(meta, meta_c, rtol=1e-6, atol=1e-7, equal_nan=True)
    assert np.array_equal(rep.genotypes, G) and np.array_equal(rep.fitnesses.ravel(), F, equal_nan=True)


@SET
@given(seed=st.integers(0, 10**6), B=st.integers(1, 60), S=st.integers(1, 7), K_side=st.integers(1, 6))
def test_mels_reduction_both_oracles_agree(co, seed, B, S, K_side):
    rng = np.random.default_rng(seed)
    cent = qn.compute_euclidean_centroids((K_side, K_side), 0.0, 1.0)
    d = np.round(rng.random((B, S, 2)), 1).astype(np.float32)
    f = np.round(rng.standard_normal((B, S)), 1).astype(np.float32)
    cells_all = qn.get_cells_indices(d.reshape(B * S, 2), cent).reshape(B, S)
    cell, spread, fmean = co.mels_reduce(cells_all, d, f)
    assert np.array_equal(cell, [qn.mels_mode(c) for c in cells_all])
    if S > 1:
        assert np.allclose(spread, [qn.mels_dispersion(x) for x in d], rtol=1e-6, atol=1e-7)
    else:
        assert (spread == 0).all()
    assert np.allclose(fmean, qn.seq_sum(f) / np.float32(S), rtol=1e-7)


@settings(max_examples=25, deadline=None, derandomize=True, database=None)
@given(seed=st.integers(0, 10**6), R=st.sampled_from([1, 2, 4]), B_dev=st.integers(1, 96), task=st.sampled_from(["arm", "rastrigin", "sphere"]))
def test_distributed_update_is_emit_per_rank_plus_one_replicated_add(co, seed, R, B_dev, task):
    """DistributedMAPElites.update (distributed_map_elites.py:124-146): every rank emits from its own key (k, e = split(key);
    emit(e)), the all_gather concatenates in rank order (global index = rank * B_dev + i) and ONE add runs on the concatenation --
    both in the C oracle's fused routine and when spelled out with its single-rank pieces, and in the NumPy oracle."""
    from oracle import jax_prng as jr

    rng = np.random.default_rng(seed)
    D = 8
    cent = qn.compute_euclidean_centroids((5, 5), 0.0, 1.0)
    init = rng.random((20, D)).astype(np.float32)
    f0, d0 = co.score(task, init)
    K = cent.shape[0]
    g, f, d, _ = co.add(np.zeros((K, D)), np.full(K, -np.inf), np.zeros((K, 2)), init, f0, d0, co.cells(d0, cent))
    keys = jr.split(jr.key(seed), R)
    G, F, Dn, og, of, od, oc = co.distributed_update(g, f, d, cent, keys, B_dev, task)
    xs = [co.emit_isoline(g, f, jr.split(keys[r])[1], B_dev, 0.05, 0.1, 0.0, 1.0)[0] for r in range(R)]
    x = np.concatenate(xs, axis=0)
    assert np.array_equal(og, x)
    fx, dx = co.score(task, x)
    assert np.array_equal(of, fx) and np.array_equal(od, dx) and np.array_equal(oc, co.cells(dx, cent))
    G2, F2, D2, _ = co.add(g, f, d, x, fx, dx, oc)
    assert np.array_equal(G, G2) and np.array_equal(F, F2) and np.array_equal(Dn, D2)
    rep = qn.Repertoire(g.copy(), f.reshape(-1, 1).copy(), d.copy(), cent)
    cfg = qn.EmitterConfig(batch_size=B_dev, iso_sigma=0.05, line_sigma=0.1, minval=0.0, maxval=1.0)
    new = qn.distributed_update(rep, list(keys), cfg, task)
    new = new[0] if isinstance(new, tuple) else new
    assert np.array_equal(np.isinf(new.fitnesses.ravel()), np.isinf(F))
    assert np.allclose(new.genotypes, G, rtol=1e-5, atol=1e-6) and np.allclose(new.fitnesses.ravel(), F, rtol=1e-5, atol=1e-6)
