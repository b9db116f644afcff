"""The two CPU oracles against each other, stage by stage: the literal NumPy restatement (oracle/qdax_numpy.py, library
log1p / sin / cos) and its exact-arithmetic C twin (oracle/qdx_oracle.c, every float32 rounding written out -- the checker of
the bit-exact GPU tests).  Integers, indices and decisions must agree bit for bit; floating-point results within 1e-6
(transcendentals differ in the last bits between libm and the spec'd polynomials)."""
import numpy as np
import pytest

from oracle import jax_prng as jr
from oracle import qdax_numpy as qn

FTOL = dict(rtol=1e-6, atol=1e-6)


def _rep(rng, K, D, occ):
    fit = np.where(rng.random(K) < occ, np.round(rng.standard_normal(K), 2), -np.inf).astype(np.float32)
    fit[0] = 0.25
    g = np.where(np.isinf(fit)[:, None], 0, rng.random((K, D))).astype(np.float32)
    return g, fit


def test_random_streams(co):
    for seed, n in [(0, 1000), (42, 4097), (2**40 + 5, 333)]:
        k = jr.key(seed)
        assert np.array_equal(co.random_bits(k, n), jr.random_bits(k, (n,)))                  # integers: bit for bit
        assert np.array_equal(co.split(k, 7), jr.split(k, 7))
        assert np.array_equal(co.uniform(k, n), jr.uniform(k, (n,)))                          # bit trick only: exact
        assert np.allclose(co.normal(k, n), jr.normal(k, (n,)), **FTOL)                      # erfinv / log1p: 1e-6


@pytest.mark.parametrize("K,occ,n", [(16, 1.0, 100), (400, 0.3, 5000), (10000, 0.68, 20000)])
def test_selection_streams(co, K, occ, n):
    rng = np.random.default_rng(K)
    _, fit = _rep(rng, K, 4, occ)
    assert np.array_equal(co.select_indices(fit, jr.key(3), n), qn.uniform_select_indices(fit, jr.key(3), n))
    m = min(n, K)
    a = co.select_indices_without_replacement(fit, jr.key(4), m)
    b = qn.uniform_select_indices_without_replacement(fit, jr.key(4), m)
    assert len(set(a.tolist())) == m and (a == b).mean() > 0.999                             # Gumbel keys within 1 ulp may swap
    n_occ = int((fit != -np.inf).sum())
    assert np.all(fit[a[: min(m, n_occ)]] != -np.inf)                                        # occupied cells first


@pytest.mark.parametrize("B,D", [(1, 4), (257, 20), (1000, 100)])
def test_isoline_and_emit(co, B, D):
    rng = np.random.default_rng(B)
    x1, x2 = rng.random((B, D)).astype(np.float32), rng.random((B, D)).astype(np.float32)
    for lo, hi in [(0.0, 1.0), (None, None)]:
        assert np.allclose(co.isoline_variation(x1, x2, jr.key(9), 0.05, 0.1, lo, hi), qn.isoline_variation(x1, x2, jr.key(9), 0.05, 0.1, lo, hi), **FTOL)
    g, fit = _rep(rng, 300, D, 0.5)
    rep = qn.Repertoire(g, fit.reshape(-1, 1), np.zeros((300, 2), np.float32), np.zeros((300, 2), np.float32))
    x, p1, p2 = co.emit_isoline(g, fit, jr.key(5), B, 0.05, 0.1, 0.0, 1.0)
    xn, q1, q2 = qn.mixing_emit_isoline(rep, jr.key(5), B, 0.05, 0.1, 0.0, 1.0)
    assert np.array_equal(p1, q1) and np.array_equal(p2, q2) and np.allclose(x, xn, **FTOL)  # parent indices: bit for bit


@pytest.mark.parametrize("task", ["arm", "rastrigin", "sphere"])
def test_scoring(co, task):
    rng = np.random.default_rng(7)
    for B, D in [(5, 2), (64, 100), (33, 1000)]:
        g = rng.random((B, D)).astype(np.float32)
        f, d = co.score(task, g)
        fn, dn = qn.SCORING[task](g)
        assert np.allclose(f, fn, rtol=2e-6, atol=2e-5 if task == "rastrigin" else 2e-6) and np.allclose(d, dn, **FTOL)


def test_polynomial_operators(co):
    rng = np.random.default_rng(11)
    x, y = rng.random((40, 30)).astype(np.float32), rng.random((40, 30)).astype(np.float32)
    assert np.allclose(co.polynomial_mutation(x, jr.key(1), 0.2, 10.0, 0.0, 1.0), qn.polynomial_mutation(x, jr.key(1), 0.2, 10.0, 0.0, 1.0), rtol=1e-5, atol=1e-6)
    assert np.array_equal(co.polynomial_crossover(x, y, jr.key(2), 0.3), qn.polynomial_crossover(x, y, jr.key(2), 0.3))   # copies only: exact


@pytest.mark.parametrize("K_shape,Dd", [((10, 10), 2), ((4, 5, 6), 3)])
def test_cells_add_metrics(co, K_shape, Dd):
    rng = np.random.default_rng(Dd)
    cent = qn.compute_euclidean_centroids(K_shape, 0.0, 1.0)
    K = cent.shape[0]
    d = rng.random((3000, Dd)).astype(np.float32)
    d[:50] = cent[rng.integers(0, K, 50)]                                                    # exactly on centroids
    cells = co.cells(d, cent)
    assert np.array_equal(cells, qn.get_cells_indices(d, cent))                              # indices: bit for bit
    g = rng.random((3000, 6)).astype(np.float32)
    f = np.round(rng.standard_normal(3000), 1).astype(np.float32)                           # ties on purpose
    for tb in ("first", "last"):
        rep = qn.repertoire_init_default(6, cent)
        new, _, _ = qn.repertoire_add(rep, g, d, f, tb)
        G, F, Dn, _ = co.add(rep.genotypes, rep.fitnesses, rep.descriptors, g, f, d, cells, tb)
        assert np.array_equal(F, new.fitnesses.ravel()) and np.array_equal(G, new.genotypes) and np.array_equal(Dn, new.descriptors)
        m, mn = co.metrics(F, 0.5), qn.default_qd_metrics(new, 0.5)
        assert np.allclose(m, [mn["qd_score"], mn["max_fitness"], mn["coverage"]], rtol=1e-6)


def test_full_scan_decisions(co):
    """Ten generations of the README configuration at reduced size: same occupancy pattern, QD metrics within 1e-5.  (Genotypes
    are not compared gene by gene: a 1-ulp difference in a normal draw is a legitimate difference between the two oracles.)"""
    cent = qn.compute_euclidean_centroids((12, 12), 0.0, 1.0)
    init = jr.uniform(jr.key(1), (40, 16))
    rep, _ = qn.map_elites_init(init, cent, jr.key(2), "arm")
    G, F, Dn, k2, M, _ = co.map_elites_scan(rep.genotypes, rep.fitnesses, rep.descriptors, cent, jr.key(3), 10, 128, "arm")
    rep2, key2, hist = qn.map_elites_scan(rep, jr.key(3), 10, qn.EmitterConfig(128), "arm")
    assert np.array_equal(np.asarray(key2), k2)
    assert np.array_equal(np.isinf(F), np.isinf(rep2.fitnesses.ravel()))
    assert np.allclose(F[~np.isinf(F)], rep2.fitnesses.ravel()[~np.isinf(F)], rtol=1e-5, atol=1e-6)
    qd = np.array([h["qd_score"] for h in hist], np.float32)
    assert np.allclose(M[:, 0], qd, rtol=1e-5, atol=1e-5)


def test_dns(co):
    rng = np.random.default_rng(5)
    P, B = 300, 40
    f = np.round(rng.standard_normal(P + B), 1).astype(np.float32)
    d = np.round(rng.random((P + B, 2)), 2).astype(np.float32)
    dn = co.dns_dominated_novelty(f, d, 3)
    assert np.allclose(dn, qn.dominated_novelty(f, d, 3), rtol=1e-6, atol=1e-7, equal_nan=True)
    meta = np.where(f != -np.inf, dn, -np.inf).astype(np.float32)
    assert np.array_equal(co.dns_survivors(meta, P), qn.dns_survivor_order(meta)[:P])        # order: bit for bit
