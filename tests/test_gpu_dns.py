"""Dominated Novelty Search through its classes (reference tests/core_test/dns_test.py:33-143, :151-276 drive
DominatedNoveltySearch.init + 5 x scan_update / ask-tell and only assert `repertoire is not None`; here every
generation is compared with the oracle bit for bit): DominatedNoveltySearch (qdax/core/dns.py:21-253) over a
DominatedNoveltyRepertoire (qdax/core/containers/dns_repertoire.py:79-273) with a MixingEmitter."""
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


def N(t):
    return t.detach().cpu().numpy()


def _make(task_fn, B, P, k, clip=True):
    from qdax_b200.core.dns import DominatedNoveltySearch
    from qdax_b200.core.emitters.mutation_operators import isoline_variation
    from qdax_b200.core.emitters.standard_emitters import MixingEmitter
    from qdax_b200.utils.metrics import default_qd_metrics

    kw = dict(iso_sigma=0.05, line_sigma=0.1)
    if clip:
        kw.update(minval=0.0, maxval=1.0)
    em = MixingEmitter(lambda x, y: (x, y), functools.partial(isoline_variation, **kw), 1.0, B)
    return DominatedNoveltySearch(task_fn, em, functools.partial(default_qd_metrics, qd_offset=0.0), population_size=P, k=k)


def _oracle_generation(co, pop, sub, B, task, k, clip):
    """One DominatedNoveltySearch.update(key=sub) on the oracle side (dns.py:151-186: key, s1 = split(sub); ask(s1):
    _, e = split(s1); emit(e); scoring; repertoire.add)."""
    g, f, d = pop
    emit_key = jr.split(jr.split(sub)[1])[1]
    lo, hi = (0.0, 1.0) if clip else (None, None)
    x, _, _ = co.emit_isoline(g, f, emit_key, B, 0.05, 0.1, lo, hi)
    fx, dx = co.score(task, x)
    G, F, Dn, meta, surv = co.dns_add(g, f, d, x, fx, dx, k)
    return (G, F, Dn), x, meta, surv


@pytest.mark.parametrize("task,B,P,k,D,clip", [("rastrigin", 10, 128, 3, 20, True), ("rastrigin", 1, 128, 3, 20, False),
                                                ("sphere", 64, 256, 5, 100, True), ("arm", 300, 1000, 3, 20, True)])
def test_dns_init_and_scan_update_vs_oracle(dev, co, task, B, P, k, D, clip):
    from qdax_b200 import lax as qlax
    from qdax_b200 import random as qr
    from qdax_b200.core.containers.dns_repertoire import DominatedNoveltyRepertoire
    from qdax_b200.tasks.arm import arm_scoring_function
    from qdax_b200.tasks.standard_functions import rastrigin_scoring_function, sphere_scoring_function

    fn = {"arm": arm_scoring_function, "rastrigin": rastrigin_scoring_function, "sphere": sphere_scoring_function}[task]
    dns = _make(fn, B, P, k, clip)
    key = qr.key(42)
    key, subkey = qr.split(key)
    init = qr.uniform(subkey, (B, D), device=dev)          # dns_test.py: the initial population has batch_size individuals
    key, subkey = qr.split(key)
    rep, state, metrics0 = dns.init(init, subkey)
    assert isinstance(rep, DominatedNoveltyRepertoire) and rep.genotypes.shape == (P, D) and rep.fitnesses.shape == (P, 1)

    # oracle: init_default (fitness -inf, genotypes 0, descriptors NaN; dns_repertoire.py:214-273) + add of the scored init
    f0, d0 = co.score(task, N(init))
    pop = co.dns_add(np.zeros((P, D), np.float32), np.full(P, -np.inf, np.float32), np.full((P, 2), np.nan, np.float32), N(init), f0, d0, k)[:3]
    assert np.array_equal(N(rep.genotypes), pop[0]) and np.array_equal(N(rep.fitnesses).ravel(), pop[1], equal_nan=True)
    assert np.array_equal(N(rep.descriptors), pop[2], equal_nan=True)
    ref0 = co.metrics(pop[1], 0.0)
    assert np.isclose(float(metrics0["qd_score"]), ref0[0], rtol=1e-5) and np.isclose(float(metrics0["coverage"]), ref0[2], rtol=1e-6)

    # five scan_update steps (dns_test.py:131-141), each compared with the oracle
    okey = np.array(key, dtype=np.uint32)
    carry = (rep, state, key)
    for it in range(5):
        carry, m = dns.scan_update(carry, None)
        ks = jr.split(okey)
        okey, sub = ks[0], ks[1]
        pop, x, meta, surv = _oracle_generation(co, pop, sub, B, task, k, clip)
        r = carry[0]
        assert np.array_equal(N(r._last_meta_fitness), meta, equal_nan=True), (it, "meta fitness")
        assert np.array_equal(N(r._last_survivors), surv), (it, "survivors")
        assert np.array_equal(N(r.genotypes), pop[0]) and np.array_equal(N(r.fitnesses).ravel(), pop[1], equal_nan=True), it
        assert np.array_equal(N(r.descriptors), pop[2], equal_nan=True), it
        ref = co.metrics(pop[1], 0.0)
        assert np.isclose(float(m["qd_score"]), ref[0], rtol=1e-5) and np.isclose(float(m["max_fitness"]), ref[1], rtol=1e-6)
    assert (np.array(carry[2]) == okey).all()

    # the lax.scan idiom of the reference test gives the same population
    (rep_s, _, key_s), ms = qlax.scan(dns.scan_update, (rep, state, key), (), length=5)
    assert torch.equal(rep_s.genotypes, carry[0].genotypes) and torch.equal(rep_s.fitnesses, carry[0].fitnesses)
    assert (np.array(key_s) == okey).all() and ms["qd_score"].shape[0] == 5


def test_dns_ask_tell_vs_oracle(dev, co):
    """reference dns_test.py:151-276: init_ask_tell + ask / user scoring / tell."""
    from qdax_b200 import random as qr
    from qdax_b200.tasks.standard_functions import rastrigin_scoring_function

    B, P, k, D = 10, 128, 3, 20
    dns = _make(rastrigin_scoring_function, B, P, k)
    key = qr.key(7)
    key, subkey = qr.split(key)
    init = qr.uniform(subkey, (B, D), device=dev)
    f, d, _ = rastrigin_scoring_function(init)
    key, subkey = qr.split(key)
    rep, state, _ = dns.init_ask_tell(genotypes=init, fitnesses=f, descriptors=d, key=subkey)
    f0, d0 = co.score("rastrigin", N(init))
    pop = co.dns_add(np.zeros((P, D), np.float32), np.full(P, -np.inf, np.float32), np.full((P, 2), np.nan, np.float32), N(init), f0, d0, k)[:3]
    for it in range(5):
        key, subkey = qr.split(key)
        x, info = dns.ask(rep, state, subkey)
        xo, _, _ = co.emit_isoline(pop[0], pop[1], jr.split(subkey)[1], B, 0.05, 0.1, 0.0, 1.0)
        assert np.array_equal(N(x), xo), it
        fx, dx, extra = rastrigin_scoring_function(x)
        rep, state, m = dns.tell(x, fx, dx, rep, state, extra, info)
        pop = co.dns_add(pop[0], pop[1], pop[2], xo, N(fx), N(dx), k)[:3]
        assert np.array_equal(N(rep.genotypes), pop[0]) and np.array_equal(N(rep.fitnesses).ravel(), pop[1], equal_nan=True), it


def test_dns_errors_mirror_reference(dev):
    """dns.py:103-104, :171-172: ValueError("Scoring function is not set.")."""
    dns = _make(None, 4, 16, 3)
    from qdax_b200 import random as qr

    with pytest.raises(ValueError, match="Scoring function is not set"):
        dns.init(qr.uniform(qr.key(0), (4, 8), device=dev), qr.key(1))
