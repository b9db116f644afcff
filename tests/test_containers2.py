"""MOMERepertoire.add and UnstructuredRepertoire.add (reference qdax/core/containers/mome_repertoire.py:
211-322, unstructured_repertoire.py:162-337).  PARITY UNPINNED: the reference has no value-level test for either and jax cannot
run here, so the known answers below are derived by hand from the source lines they cite.

CPU: the literal NumPy restatement (oracle/qdax_containers_numpy.py) against those hand-derived answers.  GPU: the native path
(qdx_mome_add; qdx_unstructured_plan / _offer + qdx_commit) bit-exact against the restatement on random batches."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import qdax_containers_numpy as qc  # noqa: E402
from oracle import qdax_numpy as qn  # noqa: E402

INF = np.float32(np.inf)
F32 = np.float32


# ------------------------------------------------------------------------------------------------- MOME, hand-derived
def test_mome_add_one_known_answers():
    L, C, D = 3, 2, 4
    f = np.full((L, C), -INF, F32)
    g = np.zeros((L, D), F32)
    d = np.zeros((L, 2), F32)
    # 1. first point into an empty cell: it is the front; every slot behind it receives a COPY of its genotype
    #    (indices = sort(i * front + L * !front), mome_repertoire.py:148-152), zero descriptors, -inf fitness
    f, g, d = qc.mome_add_one(f, g, d, np.array([1, 2], F32), np.full(D, 1.0, F32), np.array([0.1, 0.2], F32))
    assert np.array_equal(f, [[1, 2], [-INF, -INF], [-INF, -INF]]) and np.array_equal(g, np.ones((L, D)))
    assert np.array_equal(d, F32([[0.1, 0.2], [0, 0], [0, 0]]))
    # 2. a dominating point replaces it
    f, g, d = qc.mome_add_one(f, g, d, np.array([2, 3], F32), np.full(D, 2.0, F32), np.array([0.3, 0.3], F32))
    assert np.array_equal(f, [[2, 3], [-INF, -INF], [-INF, -INF]]) and np.array_equal(g, np.full((L, D), 2.0))
    # 3. a mutually non-dominated point joins the front behind it
    f, g, d = qc.mome_add_one(f, g, d, np.array([3, 1], F32), np.full(D, 3.0, F32), np.array([0.5, 0.5], F32))
    assert np.array_equal(f, [[2, 3], [3, 1], [-INF, -INF]]) and np.array_equal(g[:, 0], [2, 3, 3]) and np.array_equal(d[:, 0], F32([0.3, 0.5, 0]))
    # 4. a dominated point does not enter, but its genotype lands in the free slot (the source's behaviour)
    f, g, d = qc.mome_add_one(f, g, d, np.array([0, 0], F32), np.full(D, 4.0, F32), np.array([0.9, 0.9], F32))
    assert np.array_equal(f, [[2, 3], [3, 1], [-INF, -INF]]) and np.array_equal(g[:, 0], [2, 3, 4]) and np.array_equal(d[2], [0, 0])
    # 5. a third non-dominated point fills the front; 6. a fourth one is dropped by the truncation to L (:181)
    f, g, d = qc.mome_add_one(f, g, d, np.array([1, 4], F32), np.full(D, 5.0, F32), np.array([0.7, 0.7], F32))
    assert np.array_equal(f, [[2, 3], [3, 1], [1, 4]]) and np.array_equal(g[:, 0], [2, 3, 5])
    f2, g2, d2 = qc.mome_add_one(f, g, d, np.array([2.5, 2], F32), np.full(D, 6.0, F32), np.array([0.8, 0.8], F32))
    assert np.array_equal(f2, f) and np.array_equal(g2, g) and np.array_equal(d2, d)
    # equal points do not dominate each other (diff > 0 nowhere): both stay
    f3, g3, _ = qc.mome_add_one(np.array([[1, 1], [-INF, -INF], [-INF, -INF]], F32), np.zeros((L, D), F32), np.zeros((L, 2), F32),
                                np.array([1, 1], F32), np.ones(D, F32), np.zeros(2, F32))
    assert np.array_equal(f3, [[1, 1], [1, 1], [-INF, -INF]])
    # the IEEE reading of `cell_fitness - inf * mask` (:288): inf * 0 = NaN on every valid entry
    f4, _, _ = qc.mome_add_one(np.full((L, C), -INF, F32), np.zeros((L, D), F32), np.zeros((L, 2), F32), np.array([1, 2], F32), np.ones(D, F32),
                               np.zeros(2, F32), ieee_literal=True)
    assert np.isnan(f4[0]).all() and np.array_equal(f4[1:], np.full((2, C), -INF))


def test_mome_add_is_sequential_per_cell():
    """Two offspring of one cell in one batch: the second sees the front left by the first (lax.scan, :310-320)."""
    cent = qn.compute_euclidean_centroids((2, 2), 0.0, 1.0)
    K, L, C, D = 4, 2, 2, 3
    rf, rg, rd = np.full((K, L, C), -INF, F32), np.zeros((K, L, D), F32), np.zeros((K, L, 2), F32)
    desc = np.array([[0.1, 0.1], [0.2, 0.2], [0.9, 0.9]], F32)
    fit = np.array([[1, 1], [2, 2], [0, 5]], F32)
    gen = np.arange(9, dtype=F32).reshape(3, 3)
    cells = qn.get_cells_indices(desc, cent)
    assert list(cells) == [0, 0, 3]
    rf, rg, rd = qc.mome_add(rf, rg, rd, cent, gen, desc, fit, cells)
    assert np.array_equal(rf[0], [[2, 2], [-INF, -INF]]) and np.array_equal(rg[0], [gen[1], gen[1]])       # (1,1) was replaced by (2,2)
    assert np.array_equal(rf[3], [[0, 5], [-INF, -INF]]) and np.isinf(rf[1]).all() and np.isinf(rf[2]).all()


# ------------------------------------------------------------------------------------------------- unstructured, hand-derived
def test_unstructured_known_answers():
    N, D = 6, 3
    rg, rf, rd = np.full((N, D), np.nan, F32), np.full(N, -INF, F32), np.zeros((N, 2), F32)
    # empty archive: nothing is "near" (every distance inf), every offspring opens the next empty slot in batch order
    g = np.arange(9, dtype=F32).reshape(3, 3)
    d = np.array([[0.1, 0.1], [0.5, 0.5], [0.9, 0.9]], F32)
    f = np.array([1.0, 2.0, 3.0], F32)
    rg, rf, rd = qc.unstructured_add(rg, rf, rd, 0.1, g, d, f)
    assert np.array_equal(rf, [1, 2, 3, -INF, -INF, -INF]) and np.array_equal(rg[:3], g) and np.array_equal(rd[:3], d)
    # intra-batch competition (:69-129): two offspring closer than l, the less fit one is discarded -- but it had already
    # been dealt an empty slot (positions are fixed before the competition), which therefore stays empty
    g2 = np.arange(9, dtype=F32).reshape(3, 3) + 10
    d2 = np.array([[5.0, 5.0], [5.05, 5.0], [7.0, 7.0]], F32)
    f2 = np.array([1.0, 2.0, 0.5], F32)
    rg2, rf2, rd2 = qc.unstructured_add(rg, rf, rd, 0.1, g2, d2, f2)
    assert np.array_equal(rf2, [1, 2, 3, -INF, 2, 0.5]) and np.array_equal(rg2[4], g2[1]) and np.array_equal(rg2[5], g2[2])
    # equal fitnesses everywhere: the virtual fitness linspace(0, 1, B) (:89-99) lets the LATER of two close offspring win
    f3 = np.array([1.0, 1.0, 1.0], F32)
    _, rf3, rd3 = qc.unstructured_add(rg, rf, rd, 0.1, g2, d2, f3)
    assert np.array_equal(rf3, [1, 2, 3, -INF, 1, 1])
    # a NaN descriptor is never kept (:78, :125); -inf fitness never beats an empty slot
    d4 = np.array([[np.nan, 5.0], [6.0, 6.0], [8.0, 8.0]], F32)
    _, rf4, _ = qc.unstructured_add(rg, rf, rd, 0.1, g2, d4, np.array([9.0, -INF, 1.0], F32))
    assert np.array_equal(rf4, [1, 2, 3, -INF, -INF, 1])
    # full archive: empty_indexes is padded with -1, which wraps to the last slot like jnp indexing does
    full_f = np.arange(N, dtype=F32)
    _, rf5, _ = qc.unstructured_add(rg, full_f, rd, 0.1, g2[:1], np.array([[50.0, 50.0]], F32), np.array([100.0], F32))
    assert np.array_equal(rf5, [0, 1, 2, 3, 4, 100])


def test_unstructured_frobenius_reading():
    """`filtered_descriptors` broadcasts to (N, N, Dd) (:188-194): the distance to an occupied slot is the Frobenius norm over
    ALL stored descriptors; with two occupied slots the nearest and second-nearest distances are equal, so an offspring within
    l of that norm is "not novel enough" and never added (:211-213)."""
    rd = np.array([[0.0, 0.0], [0.3, 0.4], [0.0, 0.0]], F32)
    rf = np.array([1.0, 1.0, -INF], F32)
    F = qc.unstructured_frobenius(np.array([[0.0, 0.0]], F32), rd)
    assert np.isclose(F[0], 0.5)                                   # sqrt(0 + 0.25 + 0)
    _, rf2, _ = qc.unstructured_add(np.zeros((3, 2), F32), rf, rd, 0.6, np.ones((1, 2), F32), np.array([[0.0, 0.0]], F32), np.array([5.0], F32))
    assert np.array_equal(rf2, rf)                                 # near (0.5 <= 0.6) but not novel enough: dropped
    _, rf3, _ = qc.unstructured_add(np.zeros((3, 2), F32), rf, rd, 0.4, np.ones((1, 2), F32), np.array([[0.0, 0.0]], F32), np.array([5.0], F32))
    assert np.array_equal(rf3, [1, 1, 5])                          # not near: opens the empty slot


# ------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)


def N_(t):
    return t.detach().cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("K_shape,L,C,D,B,literal", [((4, 4), 3, 2, 8, 200, False), ((8, 8), 8, 2, 12, 1000, False), ((2, 2), 50, 3, 4, 600, False),
                                                   ((6, 6), 5, 1, 20, 500, False), ((4, 4), 4, 2, 8, 300, True)])
def test_gpu_mome_add_vs_restatement(dev, K_shape, L, C, D, B, literal):
    from qdax_b200.core.containers.mapelites_repertoire import compute_euclidean_centroids
    from qdax_b200.core.containers.mome_repertoire import MOMERepertoire

    rng = np.random.default_rng(L * 100 + C)
    cent = compute_euclidean_centroids(K_shape, 0.0, 1.0, device=dev)
    cent_h = N_(cent)
    K = cent_h.shape[0]

    def batch(n):
        f = np.round(rng.standard_normal((n, C)), 1).astype(F32)             # coarse values: ties and exact duplicates occur
        f[rng.random(n) < 0.02] = -INF                                       # an offspring with -inf fitness
        return rng.random((n, D)).astype(F32), rng.random((n, 2)).astype(F32), f

    g0, d0, f0 = batch(64)
    with pytest.warns(UserWarning):
        rep = MOMERepertoire.init(T(g0, dev), T(f0, dev), T(d0, dev), cent, L, ieee_literal=literal)
    rf, rg, rd = qc.mome_add(np.full((K, L, C), -INF, F32), np.zeros((K, L, D), F32), np.zeros((K, L, 2), F32), cent_h, g0, d0, f0,
                             qn.get_cells_indices(d0, cent_h), literal)
    assert np.array_equal(N_(rep.fitnesses), rf, equal_nan=True) and np.array_equal(N_(rep.genotypes), rg)
    for it in range(3):
        g, d, f = batch(B)
        before = N_(rep.fitnesses).copy()
        new = rep.add(T(g, dev), T(d, dev), T(f, dev))
        assert np.array_equal(N_(rep.fitnesses), before, equal_nan=True)     # value semantics
        rep = new
        rf, rg, rd = qc.mome_add(rf, rg, rd, cent_h, g, d, f, qn.get_cells_indices(d, cent_h), literal)
        assert np.array_equal(N_(rep.fitnesses), rf, equal_nan=True), f"fitnesses differ at iteration {it}"
        assert np.array_equal(N_(rep.genotypes), rg, equal_nan=True) and np.array_equal(N_(rep.descriptors), rd, equal_nan=True)
    if not literal:
        assert (rf[:, 0, 0] != -INF).sum() > K // 2 and not np.isnan(rf).any()


@pytest.mark.gpu
@pytest.mark.parametrize("Nmax,D,Dd,B,l,tie_break", [(64, 8, 2, 40, 0.15, "first"), (256, 12, 2, 100, 0.05, "last"), (128, 4, 3, 200, 0.3, "first"),
                                                     (32, 8, 2, 64, 0.1, "first"), (512, 300, 2, 50, 0.1, "first")])
def test_gpu_unstructured_add_vs_restatement(dev, Nmax, D, Dd, B, l, tie_break):
    from qdax_b200.core.containers.unstructured_repertoire import UnstructuredRepertoire

    rng = np.random.default_rng(Nmax + B)

    def batch(n, it):
        g = rng.random((n, D)).astype(F32)
        d = np.round(rng.random((n, Dd)) * 2.0, 2).astype(F32)               # coarse: offspring closer than l to each other occur
        f = np.round(rng.standard_normal(n), 1).astype(F32)
        if it == 1:
            f[:] = 0.5                                                       # all-equal fitness: the virtual-fitness branch
        if it == 2:
            d[3, 0] = np.nan
            f[5] = -INF
            f[7] = np.nan
        return g, d, f

    g0, d0, f0 = batch(min(B, 16), 0)
    rep = UnstructuredRepertoire.init(T(g0, dev), T(f0, dev), T(d0, dev), torch.tensor([l], device=dev), Nmax, tie_break=tie_break)
    rg, rf, rd = qc.unstructured_add(np.full((Nmax, D), np.nan, F32), np.full(Nmax, -INF, F32), np.zeros((Nmax, Dd), F32), l, g0, d0, f0, tie_break)
    assert np.array_equal(N_(rep.fitnesses).ravel(), rf, equal_nan=True) and np.array_equal(N_(rep.genotypes), rg, equal_nan=True)
    for it in range(1, 5):
        g, d, f = batch(B, it)
        rep = rep.add(T(g, dev), T(d, dev), T(f, dev))
        rg, rf, rd = qc.unstructured_add(rg, rf, rd, l, g, d, f, tie_break)
        assert np.array_equal(N_(rep.fitnesses).ravel(), rf, equal_nan=True), f"fitnesses differ at iteration {it}"
        assert np.array_equal(N_(rep.genotypes), rg, equal_nan=True) and np.array_equal(N_(rep.descriptors), rd, equal_nan=True)
    assert int(rep.get_number_genotypes()) == int((rf != -INF).sum()) and rep.get_maximal_size() == Nmax
