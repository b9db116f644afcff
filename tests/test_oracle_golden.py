"""The reference's own known-answer tests for the hot path, re-expressed against both
oracles, plus the committed golden vectors (tests/golden/hotpath_v1.npz, made by tools/make_golden.py)."""
import numpy as np

from oracle import jax_prng as jr
from oracle import qdax_numpy as qn


def test_ref_euclidean_centroids_2x2():
    # /root/reference/tests/core_test/containers_test/mapelites_repertoire_test.py:21-36
    c = qn.compute_euclidean_centroids((2, 2), 0.0, 1.0)
    assert np.allclose(c, [[0.25, 0.25], [0.75, 0.25], [0.25, 0.75], [0.75, 0.75]], atol=1e-6)


def test_ref_repertoire_add(co):
    # same file :39-81
    c = qn.compute_euclidean_centroids((2, 2), 0.0, 1.0)
    rep = qn.repertoire_init(np.zeros((4, 12)), np.full(4, -np.inf), np.zeros((4, 2)), c)
    assert (rep.fitnesses == -np.inf).all() and (rep.genotypes == 0).all()
    g, d, f = np.ones((2, 12), np.float32), np.array([[0.1, 0.1], [0.9, 0.9]], np.float32), np.zeros(2, np.float32)
    new, cells, _ = qn.repertoire_add(rep, g, d, f)
    exp_g = np.array([[1.0] * 12, [0.0] * 12, [0.0] * 12, [1.0] * 12])
    for G, F, D in [
        (new.genotypes, new.fitnesses.ravel(), new.descriptors),
        co.add(rep.genotypes, rep.fitnesses, rep.descriptors, g, f, d, co.cells(d, c))[:3],
    ]:
        assert np.allclose(G, exp_g, atol=1e-6)
        assert np.array_equal(F, np.array([0.0, -np.inf, -np.inf, 0.0], np.float32))
        assert np.allclose(D, [[0.1, 0.1], [0, 0], [0, 0], [0.9, 0.9]], atol=1e-6)


def test_ref_arm_descriptors(co):
    # /root/reference/tests/tasks_test/arm_test.py:123-163
    cases = [
        (np.ones((1, 4)) * 0.5, [1.0, 0.5]), (np.zeros((1, 6)), [0.5, 0.5]), (np.ones((1, 10)), [0.5, 0.5]),
        (np.array([[0, 0.5]]), [0.0, 0.5]), (np.array([[0.25, 0.5]]), [0.5, 0.0]),
        (np.array([[0.5, 0.5]]), [1.0, 0.5]), (np.array([[0.75, 0.5]]), [0.5, 1.0]),
    ]
    for g, exp in cases:
        assert np.array_equal(np.around(qn.arm_scoring_function(g)[1], 1) + 0.0, np.array([exp], np.float32))
        assert np.array_equal(np.around(co.score("arm", g)[1], 1) + 0.0, np.array([exp], np.float32))


def test_centroid_layout_matches_meshgrid_xy():
    c = qn.compute_euclidean_centroids((3, 5), 0.0, 1.0)  # cell = iy*nx + ix
    assert c.shape == (15, 2)
    assert np.allclose(c[1], [0.5, 0.1]) and np.allclose(c[3], [1 / 6, 0.3])
    c3 = qn.compute_euclidean_centroids((2, 3, 4), 0.0, 1.0)  # cell = i1*(n0*n2) + i0*n2 + i2
    assert np.allclose(c3[1 * 8 + 1 * 4 + 2], [0.75, 0.5, 0.625])


def test_golden_prng(golden):
    assert (golden["split_key42"] == jr.split(jr.key(42))).all()
    assert (golden["split3_key7"] == jr.split(jr.key(7), 3)).all()
    assert (golden["bits_key0_8"] == jr.random_bits(jr.key(0), (8,))).all()
    assert (golden["uniform_key0_8"] == jr.uniform(jr.key(0), (8,))).all()
    assert np.allclose(golden["normal_key0_8"], jr.normal(jr.key(0), (8,)), rtol=1e-6, atol=1e-7)


def test_golden_scenario_both_oracles(golden, co):
    g = golden
    key, cent = g["S_key"], g["S_centroids"]
    B = g["S_emit_x"].shape[0]
    assert (g["S_select_idx"] == co.select_indices(g["S_rep_f"], key, B)).all()
    assert (g["S_select_idx"] == qn.uniform_select_indices(g["S_rep_f"].reshape(-1, 1), key, B)).all()
    x, p1, p2 = co.emit_isoline(g["S_rep_g"], g["S_rep_f"], key, B, 0.05, 0.1, 0.0, 1.0)
    assert np.array_equal(x, g["S_emit_x"]) and (p1 == g["S_emit_p1"]).all() and (p2 == g["S_emit_p2"]).all()
    rep = qn.Repertoire(g["S_rep_g"], g["S_rep_f"].reshape(-1, 1), g["S_rep_d"], cent)
    xn, i1, i2 = qn.mixing_emit_isoline(rep, key, B, 0.05, 0.1, 0.0, 1.0)
    assert (i1 == p1).all() and (i2 == p2).all() and np.allclose(xn, x, rtol=0, atol=3e-7)
    for task in ("arm", "rastrigin", "sphere"):
        f, d = co.score(task, x)
        assert np.array_equal(f, g[f"S_{task}_f"]) and np.array_equal(d, g[f"S_{task}_d"])
        fn, dn = qn.SCORING[task](x)
        assert np.allclose(fn, f, rtol=2e-6) and np.allclose(dn, d, atol=1e-6)
    assert (co.cells(g["S_arm_d"], cent) == g["S_arm_cells"]).all()
    assert (qn.get_cells_indices(g["S_arm_d"], cent) == g["S_arm_cells"]).all()
    for tb in ("first", "last"):
        G, F, D, sidx = co.add(g["S_rep_g"], g["S_rep_f"], g["S_rep_d"], x, g["S_inj_f"], g["S_inj_d"], g["S_inj_cells"], tb)
        new, _, idx = qn.repertoire_add(rep, x, g["S_inj_d"], g["S_inj_f"], tb, cells=g["S_inj_cells"])
        for a, b, c in [(G, new.genotypes, g[f"S_add_{tb}_g"]), (F, new.fitnesses.ravel(), g[f"S_add_{tb}_f"]),
                        (D, new.descriptors, g[f"S_add_{tb}_d"]), (sidx, idx, g[f"S_add_{tb}_sidx"])]:
            assert np.array_equal(a, b, equal_nan=True) and np.array_equal(a, c, equal_nan=True)


def test_golden_c1mini_full_run(golden, co):
    g = golden
    K, D = g["C1mini_g"].shape
    cent = g["S_centroids"]
    init = g["C1mini_init"]
    f0, d0 = co.score("arm", init)
    g0, ff0, dd0, _ = co.add(np.zeros((K, D)), np.full(K, -np.inf), np.zeros((K, 2)), init, f0, d0, co.cells(d0, cent))
    gN, fN, dN, kN, mN, _ = co.map_elites_scan(g0, ff0, dd0, cent, jr.key(5), 10, 64, "arm")
    assert np.array_equal(gN, g["C1mini_g"]) and np.array_equal(fN, g["C1mini_f"]) and np.array_equal(dN, g["C1mini_d"])
    assert (kN == g["C1mini_key"]).all() and np.array_equal(mN, g["C1mini_metrics"])
    # the literal NumPy restatement follows the same trajectory on this run (library sin/cos differ by <= 1 ulp)
    rep = qn.repertoire_init(init, *qn.arm_scoring_function(init), cent)
    rep, key, mets = qn.map_elites_scan(rep, jr.key(5), 10, qn.EmitterConfig(batch_size=64))
    assert (key == kN).all()
    assert np.isclose(float(mets[-1]["coverage"]), float(mN[-1, 2]), atol=1.0)
    assert np.isclose(float(mets[-1]["qd_score"]), float(mN[-1, 0]), rtol=0.05)


def test_golden_dns(golden, co):
    g = golden
    G, F, D, meta, surv = co.dns_add(g["DNS_pg"], g["DNS_pf"], g["DNS_pd"], g["DNS_bg"], g["DNS_bf"], g["DNS_bd"], 3)
    assert np.array_equal(meta, g["DNS_meta"], equal_nan=True) and (surv == g["DNS_surv"]).all()
    rep = qn.DNSRepertoire(g["DNS_pg"], g["DNS_pf"].reshape(-1, 1), g["DNS_pd"], 3)
    new, meta_n, surv_n = qn.dns_add(rep, g["DNS_bg"], g["DNS_bd"], g["DNS_bf"])
    assert np.allclose(meta_n, meta, rtol=1e-6, equal_nan=True)
    assert (surv_n == surv).all()
    assert np.array_equal(new.genotypes, G) and np.array_equal(new.fitnesses.ravel(), F, equal_nan=True)
