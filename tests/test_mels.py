"""MELSRepertoire.add (reference qdax/core/containers/mels_repertoire.py:89-230).

CPU: the two oracles against the reference's own known-answer test (tests/core_test/containers_test/
mels_repertoire_test.py:8-163) and against each other on random batches.  GPU: the native path (qdx_cells + qdx_mels_offer +
qdx_commit + qdx_scatter_rows_by_source) against the same known answers and bit-exact against the C oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import qdax_numpy as qn  # noqa: E402

CENT = np.array([[1.0, 1.0], [2.0, 1.0], [2.0, 2.0], [1.0, 2.0]], dtype=np.float32)
INF = np.float32(np.inf)


def reference_kat_steps():
    """(genotypes, descriptors, fitnesses) of the reference test's two additions."""
    one = (np.ones((1, 12), np.float32), np.array([[[0.0, 1.0], [1.0, 1.0]]], np.float32), np.array([[0.0, 0.0]], np.float32))
    two = (np.concatenate([np.full((1, 12), 2.0), np.full((1, 12), 3.0)]).astype(np.float32),
           np.array([[[1.0, 0.25], [1.0, 1.0]], [[1.0, 0.5], [1.0, 1.0]]], np.float32), np.array([[1.0, 1.0], [0.5, 0.5]], np.float32))
    return one, two


def check_kat_state(g, f, d, s, step, tie_break):
    if step == 1:        # mels_repertoire_test.py:47-75
        assert np.allclose(g[0], 1.0) and np.allclose(g[1:], 0.0)
        assert np.allclose(f, [0.0, -INF, -INF, -INF]) and np.allclose(d, [[1, 1], [0, 0], [0, 0], [0, 0]])
        assert np.allclose(s, [1.0, INF, INF, INF])
    else:                # :103-163 -- either candidate is acceptable to the reference; ours is picked by tie_break
        val, fit, spr = (2.0, 1.0, 0.75) if tie_break == "first" else (3.0, 0.5, 0.5)
        assert np.allclose(g[0], val) and np.allclose(g[1:], 0.0)
        assert np.allclose(f, [fit, -INF, -INF, -INF]) and np.allclose(d, [[1, 1], [0, 0], [0, 0], [0, 0]])
        assert np.allclose(s, [spr, INF, INF, INF])


@pytest.mark.parametrize("tie_break", ["first", "last"])
def test_oracles_reproduce_reference_kat(co, tie_break):
    rep = qn.repertoire_init_default(12, CENT)
    spreads = np.full(4, INF, np.float32)
    # MELSRepertoire.init adds the (all -inf) initial batch with num_samples = 1: nothing is inserted (:16-29 of the test)
    rep, spreads = qn.mels_add(rep, spreads, np.zeros((4, 12)), np.zeros((4, 1, 2)), np.full((4, 1), -INF), tie_break)
    assert np.isinf(rep.fitnesses).all() and np.isinf(spreads).all()
    cg, cf, cd, cs = rep.genotypes.copy(), rep.fitnesses.ravel().copy(), rep.descriptors.copy(), spreads.copy()
    for step, (g, d, f) in enumerate(reference_kat_steps(), start=1):
        rep, spreads = qn.mels_add(rep, spreads, g, d, f, tie_break)
        check_kat_state(rep.genotypes, rep.fitnesses.ravel(), rep.descriptors, spreads, step, tie_break)
        cg, cf, cd, cs = co.mels_add(cg, cf, cd, cs, CENT, g, d, f, tie_break)
        check_kat_state(cg, cf, cd, cs, step, tie_break)


def random_batch(rng, B, S, D, Dd):
    g = rng.random((B, D)).astype(np.float32)
    base = rng.random((B, 1, Dd)).astype(np.float32)
    d = (base + 0.15 * rng.standard_normal((B, S, Dd))).astype(np.float32)
    d[: B // 8] = base[: B // 8]                    # some individuals with identical samples: spread 0
    f = rng.standard_normal((B, S)).astype(np.float32)
    return g, d, f


@pytest.mark.parametrize("S", [1, 2, 5])
def test_c_oracle_matches_numpy_oracle(co, S):
    rng = np.random.default_rng(S)
    cent = qn.compute_euclidean_centroids((6, 6), 0.0, 1.0)
    rep = qn.repertoire_init_default(8, cent)
    spreads = np.full(36, INF, np.float32)
    cg, cf, cd, cs = rep.genotypes.copy(), rep.fitnesses.ravel().copy(), rep.descriptors.copy(), spreads.copy()
    for it in range(3):
        g, d, f = random_batch(rng, 64, S, 8, 2)
        cell, spread, fmean = co.mels_reduce(qn.get_cells_indices(d.reshape(64 * S, 2), cent), d, f)
        assert np.array_equal(cell, [qn.mels_mode(c) for c in qn.get_cells_indices(d.reshape(64 * S, 2), cent).reshape(64, S)])
        if S > 1:
            assert np.allclose(spread, [qn.mels_dispersion(x) for x in d], rtol=1e-6, atol=1e-7)
        rep, spreads = qn.mels_add(rep, spreads, g, d, f)
        cg, cf, cd, cs = co.mels_add(cg, cf, cd, cs, cent, g, d, f)
        assert np.array_equal(np.isinf(cf), np.isinf(rep.fitnesses.ravel()))
        assert np.allclose(cg, rep.genotypes) and np.allclose(cs, spreads, rtol=1e-6) and np.allclose(cf, rep.fitnesses.ravel(), rtol=1e-6)


# ------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)


def N(t):
    return t.detach().cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("tie_break", ["first", "last"])
def test_gpu_mels_reference_kat(dev, tie_break):
    from qdax_b200.core.containers.mels_repertoire import MELSRepertoire

    rep = MELSRepertoire.init(genotypes=torch.zeros(4, 12, device=dev), fitnesses=torch.full((4, 1), -np.inf, device=dev),
                              descriptors=torch.zeros(4, 2, device=dev), centroids=T(CENT, dev), tie_break=tie_break)
    assert np.isinf(N(rep.fitnesses)).all() and np.isinf(N(rep.spreads)).all()
    for step, (g, d, f) in enumerate(reference_kat_steps(), start=1):
        before = N(rep.genotypes).copy()
        new = rep.add(T(g, dev), T(d, dev), T(f, dev), {})
        assert np.array_equal(N(rep.genotypes), before)                       # value semantics
        rep = new
        check_kat_state(N(rep.genotypes), N(rep.fitnesses).ravel(), N(rep.descriptors), N(rep.spreads), step, tie_break)


@pytest.mark.gpu
@pytest.mark.parametrize("S,tie_break", [(1, "first"), (2, "last"), (5, "first"), (16, "last")])
def test_gpu_mels_bit_exact_vs_oracle(dev, co, S, tie_break):
    from qdax_b200.core.containers.mapelites_repertoire import compute_euclidean_centroids
    from qdax_b200.core.containers.mels_repertoire import MELSRepertoire

    rng = np.random.default_rng(10 + S)
    cent = compute_euclidean_centroids((12, 12), 0.0, 1.0, device=dev)
    cent_h = N(cent)
    K, D = cent_h.shape[0], 20
    rep = MELSRepertoire.init_default(torch.zeros(D, device=dev), cent, tie_break=tie_break)
    cg, cf, cd, cs = np.zeros((K, D), np.float32), np.full(K, -INF, np.float32), np.zeros((K, 2), np.float32), np.full(K, INF, np.float32)
    for it in range(4):
        g, d, f = random_batch(rng, 300, S, D, 2)
        if it == 2:
            f[5] = np.nan                                # a NaN mean never passes `>`
        rep = rep.add(T(g, dev), T(d, dev), T(f, dev))
        cg, cf, cd, cs = co.mels_add(cg, cf, cd, cs, cent_h, g, d, f, tie_break)
        assert np.array_equal(N(rep.fitnesses).ravel(), cf), f"fitnesses differ at iteration {it}"
        assert np.array_equal(N(rep.genotypes), cg) and np.array_equal(N(rep.descriptors), cd) and np.array_equal(N(rep.spreads), cs)
    assert (~np.isinf(cf)).sum() > 20
