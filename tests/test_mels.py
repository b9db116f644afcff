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


# ------------------------------------------------------------------------------------------------- noisy arm + MELS end to end
def test_noisy_arm_oracles_agree(co):
    """noisy_arm_scoring_function (qdax/tasks/arm.py:53-81): literal NumPy restatement vs the exact-arithmetic C twin."""
    from oracle import jax_prng as jr

    rng = np.random.default_rng(3)
    g = rng.random((200, 24)).astype(np.float32)
    for key, (fv, dv, pv) in [(jr.key(1), (0.01, 0.01, 0.05)), (jr.key(2), (0.0, 0.0, 0.0)), (jr.key(9), (0.3, 0.0, 0.0))]:
        f1, d1 = co.noisy_arm(g, key, fv, dv, pv)
        f2, d2 = qn.noisy_arm_scoring_function(g, key, fv, dv, pv)
        assert np.allclose(f1, f2, rtol=1e-5, atol=1e-6) and np.allclose(d1, d2, rtol=1e-5, atol=1e-6)
    f0, d0 = co.score("arm", g)
    f3, d3 = co.noisy_arm(g, jr.key(2), 0.0, 0.0, 0.0)      # x + 0 * n = x: zero variances reduce to the deterministic arm
    assert np.array_equal(f0, f3) and np.array_equal(d0, d3)


@pytest.mark.gpu
@pytest.mark.parametrize("B,D", [(1, 4), (33, 20), (1000, 100), (130, 200)])
def test_gpu_noisy_arm_bit_exact(dev, co, B, D):
    from oracle import jax_prng as jr
    from qdax_b200.tasks.arm import noisy_arm_scoring_function

    g = np.random.default_rng(B).random((B, D)).astype(np.float32)
    f, d, extra = noisy_arm_scoring_function(T(g, dev), jr.key(5), fit_variance=0.01, desc_variance=0.02, params_variance=0.05)
    fo, do = co.noisy_arm(g, jr.key(5), 0.01, 0.02, 0.05)
    assert extra == {} and np.array_equal(N(f), fo) and np.array_equal(N(d), do)
    fn, dn = qn.noisy_arm_scoring_function(g, jr.key(5), 0.01, 0.02, 0.05)
    assert np.allclose(N(f), fn, rtol=1e-5, atol=1e-6) and np.allclose(N(d), dn, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("batch_size,custom_repertoire", [(1, False), (10, False), (10, True), (64, False)])
def test_gpu_mels_end_to_end(dev, co, batch_size, custom_repertoire):
    """reference tests/core_test/mels_test.py:25-183 (MELS, and MAPElites + multi_sample_scoring_function + MELSRepertoire.init;
    5 samples, 5 iterations; it only asserts `repertoire is not None`) with the noisy arm as the stochastic task -- and every
    iteration compared with the oracle: emit, num_samples noisy evaluations with split(key, num_samples), MELSRepertoire.add."""
    import functools

    from oracle import jax_prng as jr
    from qdax_b200 import lax as qlax
    from qdax_b200 import random as qr
    from qdax_b200.core.containers.mapelites_repertoire import compute_euclidean_centroids
    from qdax_b200.core.containers.mels_repertoire import MELSRepertoire
    from qdax_b200.core.emitters.mutation_operators import isoline_variation
    from qdax_b200.core.emitters.standard_emitters import MixingEmitter
    from qdax_b200.core.map_elites import MAPElites
    from qdax_b200.core.mels import MELS
    from qdax_b200.tasks.arm import noisy_arm_scoring_function
    from qdax_b200.utils.metrics import default_qd_metrics
    from qdax_b200.utils.sampling import multi_sample_scoring_function

    S, D, iters = 5, 20, 5
    var = dict(fit_variance=0.01, desc_variance=0.02, params_variance=0.05)
    scoring = functools.partial(noisy_arm_scoring_function, **var)
    em = MixingEmitter(lambda x, y: (x, y), functools.partial(isoline_variation, iso_sigma=0.05, line_sigma=0.1, minval=0.0, maxval=1.0), 1.0, batch_size)
    metrics_fn = functools.partial(default_qd_metrics, qd_offset=0.0)
    if custom_repertoire:
        mels = MAPElites(functools.partial(multi_sample_scoring_function, scoring_fn=scoring, num_samples=S), em, metrics_fn,
                         repertoire_init=MELSRepertoire.init)
    else:
        mels = MELS(scoring, em, metrics_fn, num_samples=S)
    cent = compute_euclidean_centroids((8, 8), 0.0, 1.0, device=dev)
    cent_h, K = N(cent), 64
    key = qr.key(42)
    key, subkey = qr.split(key)
    init = qr.uniform(subkey, (batch_size, D), device=dev)
    key, subkey = qr.split(key)
    rep, state, metrics0 = mels.init(init, cent, subkey)
    assert isinstance(rep, MELSRepertoire)

    def oracle_scores(x, score_key):
        ks = jr.split(score_key, S)                                  # sampling.py:137
        outs = [co.noisy_arm(x, ks[s], var["fit_variance"], var["desc_variance"], var["params_variance"]) for s in range(S)]
        return np.stack([o[0] for o in outs], axis=1), np.stack([o[1] for o in outs], axis=1)

    # init: key, s = split(subkey); scoring(init, s)  (map_elites.py:81-82)
    f_all, d_all = oracle_scores(N(init), jr.split(subkey)[1])
    pop = co.mels_add(np.zeros((K, D), np.float32), np.full(K, -INF, np.float32), np.zeros((K, 2), np.float32), np.full(K, INF, np.float32),
                      cent_h, N(init), d_all, f_all)
    assert np.array_equal(N(rep.fitnesses).ravel(), pop[1]) and np.array_equal(N(rep.genotypes), pop[0]) and np.array_equal(N(rep.spreads), pop[3])

    okey = np.array(key, dtype=np.uint32)
    carry = (rep, state, key)
    for it in range(iters):
        carry, m = mels.scan_update(carry, None)
        ks = jr.split(okey)
        okey, sub = ks[0], ks[1]
        ku, a = jr.split(sub)                                        # map_elites.py:177
        x, _, _ = co.emit_isoline(pop[0], pop[1], jr.split(a)[1], batch_size, 0.05, 0.1, 0.0, 1.0)    # :241
        f_all, d_all = oracle_scores(x, jr.split(ku)[1])             # :181
        pop = co.mels_add(pop[0], pop[1], pop[2], pop[3], cent_h, x, d_all, f_all)
        r = carry[0]
        assert np.array_equal(N(r.fitnesses).ravel(), pop[1]), it
        assert np.array_equal(N(r.genotypes), pop[0]) and np.array_equal(N(r.descriptors), pop[2]) and np.array_equal(N(r.spreads), pop[3]), it
        ref = co.metrics(pop[1], 0.0)
        assert np.isclose(float(m["qd_score"]), ref[0], rtol=1e-5, atol=1e-6) and np.isclose(float(m["coverage"]), ref[2], rtol=1e-6)
    assert (np.array(carry[2]) == okey).all()
    # the jax.lax.scan idiom of the reference test
    (rep_s, _, key_s), ms = qlax.scan(mels.scan_update, (rep, state, key), (), length=iters)
    assert torch.equal(rep_s.genotypes, carry[0].genotypes) and torch.equal(rep_s.spreads, carry[0].spreads) and (np.array(key_s) == okey).all()
