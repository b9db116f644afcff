"""GPU parity at the BASELINE.json sizes: the CUDA path through the public classes vs the
exact-arithmetic C oracle on the same seeded inputs, bit for bit.  The oracle needs a few seconds per case on the
host cores of the GPU box (brute-force cells over K centroids, dense N^2 DNS competition)."""
import functools

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import jax_prng as jr  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def T(a, dev, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dtype)


def N(t):
    return t.detach().cpu().numpy()


def _driver(task, B, desc_dim=2):
    from qdax_b200.core.emitters.mutation_operators import isoline_variation
    from qdax_b200.core.emitters.standard_emitters import MixingEmitter
    from qdax_b200.core.map_elites import MAPElites
    from qdax_b200.tasks.arm import arm_scoring_function
    from qdax_b200.tasks.standard_functions import rastrigin_scoring_function, sphere_scoring_function
    from qdax_b200.utils.metrics import default_qd_metrics

    scoring = {"arm": arm_scoring_function, "rastrigin": rastrigin_scoring_function,
               "sphere": functools.partial(sphere_scoring_function, desc_dim=desc_dim)}[task]
    em = MixingEmitter(lambda x, y: x, functools.partial(isoline_variation, iso_sigma=0.05, line_sigma=0.1, minval=0.0, maxval=1.0), 1.0, B)
    return MAPElites(scoring, em, functools.partial(default_qd_metrics, qd_offset=0.0))


def _oracle_init(co, init, cent, task, Dd=2):
    K, D = cent.shape[0], init.shape[1]
    f0, d0 = co.score(task, init, Dd)
    g, f, d, _ = co.add(np.zeros((K, D)), np.full(K, -np.inf), np.zeros((K, Dd)), init, f0, d0, co.cells(d0, cent))
    return g, f, d


def _run_and_compare(dev, co, task, cent, D, B, gens, init_n, Dd=2):
    from qdax_b200 import lax as qlax
    from qdax_b200 import random as qr

    me = _driver(task, B, Dd)
    init = N(qr.uniform(jr.key(1), (init_n, D), device=dev))
    rep, state, _ = me.init(T(init, dev), T(cent, dev), jr.key(2))
    assert me._fused_config(rep) is not None, "the BASELINE configuration must take the fused native path"
    g, f, d = _oracle_init(co, init, cent, task, Dd)
    assert np.array_equal(N(rep.genotypes), g) and np.array_equal(N(rep.fitnesses).ravel(), f)
    (rep2, _, key2), metrics = qlax.scan(me.scan_update, (rep, state, jr.key(3)), (), length=gens)
    G, F, Dn, k2, M, _ = co.map_elites_scan(g, f, d, cent, jr.key(3), gens, B, task)
    assert (np.array(key2) == k2).all()
    assert np.array_equal(N(rep2.fitnesses).ravel(), F), "insertion decisions / fitnesses differ from the oracle"
    assert np.array_equal(N(rep2.descriptors), Dn), "descriptors differ from the oracle"
    assert np.array_equal(N(rep2.genotypes), G), "genotypes differ from the oracle"
    assert np.allclose(N(metrics["qd_score"]), M[:, 0], rtol=1e-5) and np.allclose(N(metrics["coverage"]), M[:, 2], rtol=1e-6)
    return rep2, metrics


def test_c3_full_size_vs_oracle(dev, co):
    """BASELINE configs[2] on one GPU: arm 100-DoF, grid 100x100, B = 2^20, two generations, whole repertoire bit-exact."""
    from oracle import qdax_numpy as qn

    cent = qn.compute_euclidean_centroids((100, 100), 0.0, 1.0)
    rep, m = _run_and_compare(dev, co, "arm", cent, 100, 1 << 20, 2, 100)
    assert float(m["coverage"][-1]) > 10.0


def test_c2_full_size_vs_oracle(dev, co):
    """BASELINE configs[1]: rastrigin 100-D, K = 10^4 non-grid centroids (bucket-index cells on the GPU, brute force in the
    oracle), B = 65 536, three generations."""
    cent = np.random.default_rng(0).random((10000, 2)).astype(np.float32)
    _run_and_compare(dev, co, "rastrigin", cent, 100, 65536, 3, 200)


def test_c4_full_size_vs_oracle(dev, co):
    """BASELINE configs[3]: sphere 1000-D, descriptor = first 32 genes (declared extension), K = 50 000 centroids in 32-D
    (tcgen05 TF32 pass + exact FP32 re-rank on the GPU, FP32 brute force in the oracle), B = 65 536, one generation."""
    cent = np.random.default_rng(0).random((50000, 32)).astype(np.float32)
    _run_and_compare(dev, co, "sphere", cent, 1000, 65536, 1, 200, Dd=32)


def test_c5_full_size_dns_add_vs_oracle(dev, co):
    """BASELINE configs[4]: Dominated Novelty Search, rastrigin 100-D, population 100 000 (all valid), batch 1024, k = 3:
    one DominatedNoveltyRepertoire.add -- meta fitness of all N = 101 024 candidates and the survivor order, bit-exact."""
    from qdax_b200 import random as qr
    from qdax_b200.core.containers.dns_repertoire import DominatedNoveltyRepertoire
    from qdax_b200.tasks.standard_functions import rastrigin_scoring_function

    P, B, D, k = 100000, 1024, 100, 3
    pg = qr.uniform(jr.key(2), (P, D), device=dev)
    pf, pd, _ = rastrigin_scoring_function(pg)
    bg = qr.uniform(jr.key(3), (B, D), device=dev)
    bf, bd, _ = rastrigin_scoring_function(bg)
    rep = DominatedNoveltyRepertoire(genotypes=pg, fitnesses=pf.reshape(P, 1), descriptors=pd, k=k)
    new = rep.add(bg, bd, bf)
    G, F, Dn, meta, surv = co.dns_add(N(pg), N(pf), N(pd), N(bg), N(bf), N(bd), k)
    assert np.array_equal(N(new._last_meta_fitness), meta, equal_nan=True), "dominated novelty differs from the oracle"
    assert np.array_equal(N(new._last_survivors), surv), "survivor order differs from the oracle"
    assert np.array_equal(N(new.fitnesses).ravel(), F, equal_nan=True) and np.array_equal(N(new.descriptors), Dn, equal_nan=True)
    assert np.array_equal(N(new.genotypes), G)
    assert np.array_equal(N(rep.fitnesses).ravel(), N(pf))     # value semantics: the input repertoire is untouched
