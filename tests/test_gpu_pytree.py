"""GPU parity tests for pytree genotypes (reference qdax/core/emitters/mutation_operators.py:205-224,
qdax/core/containers/mapelites_repertoire.py:234-240,342-347): the packed-row kernels vs the oracle, bit-exact against the C
oracle and within 1e-6 of the literal NumPy restatement."""
import functools

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import jax_prng as jr  # noqa: E402
from oracle import qdax_numpy as qn  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)


def N(t):
    return t.detach().cpu().numpy()


# an MLP-like parameter tree (flax layout: {"params": {"Dense_0": {"bias", "kernel"}, ...}}); jax.tree.leaves order is
# Dense_0/bias, Dense_0/kernel, Dense_1/bias, Dense_1/kernel
SHAPES = {"Dense_0": {"kernel": (6, 5), "bias": (5,)}, "Dense_1": {"kernel": (5, 3), "bias": (3,)}}
ORDER = [("Dense_0", "bias"), ("Dense_0", "kernel"), ("Dense_1", "bias"), ("Dense_1", "kernel")]
SIZES = [5, 30, 3, 15]          # total 53 -> not a multiple of 4: generic emit path; see SHAPES4 for the fused one
SHAPES4 = {"Dense_0": {"kernel": (6, 5), "bias": (5,)}, "Dense_1": {"kernel": (5, 3), "bias": (3,)}, "scale": (3,)}
ORDER4 = ORDER + [("scale",)]
SIZES4 = SIZES + [3]            # total 56


def make_tree(rng, n, shapes, dev):
    def rec(node):
        if isinstance(node, dict):
            return {k: rec(v) for k, v in node.items()}
        return T(rng.random((n,) + node), dev)
    return {"params": rec(shapes)}


def packed(tree, order):
    cols = []
    for path in order:
        node = tree["params"]
        for k in path:
            node = node[k]
        a = N(node)
        cols.append(a.reshape(a.shape[0], -1))
    return np.concatenate(cols, axis=1)


def test_tree_order_pack_unpack_roundtrip(dev):
    from qdax_b200 import tree_util as tu

    rng = np.random.default_rng(0)
    tree = make_tree(rng, 9, SHAPES, dev)
    leaves, _ = tu.tree_flatten(tree)
    assert [tuple(l.shape[1:]) for l in leaves] == [(5,), (6, 5), (3,), (5, 3)]       # sorted dict keys, like jax
    flat, spec = tu.pack(tree)
    assert spec.offsets == (0, 5, 35, 38, 53)
    assert np.array_equal(N(flat), packed(tree, ORDER))
    back = tu.unpack(flat, spec)
    assert all(np.array_equal(N(a), N(b)) for a, b in zip(tu.tree_leaves(back), leaves))
    flat2, _ = tu.pack(back)
    assert flat2 is flat                                                              # views of one buffer: zero copy
    back["params"]["Dense_0"]["bias"] = back["params"]["Dense_0"]["bias"].clone()     # a foreign leaf forces a real pack
    flat3, _ = tu.pack(back)
    assert flat3 is not flat and np.array_equal(N(flat3), N(flat))


@pytest.mark.parametrize("clip", [(None, None), (0.0, 1.0), (0.2, None)])
def test_isoline_variation_pytree(dev, co, clip):
    from qdax_b200 import tree_util as tu
    from qdax_b200.core.emitters.mutation_operators import isoline_variation

    rng = np.random.default_rng(1)
    B = 37
    x1, x2 = make_tree(rng, B, SHAPES, dev), make_tree(rng, B, SHAPES, dev)
    key = jr.key(11)
    out = isoline_variation(x1, x2, key, iso_sigma=0.05, line_sigma=0.1, minval=clip[0], maxval=clip[1])
    got = packed(out, ORDER)
    ref = co.isoline_variation_leaves(packed(x1, ORDER), packed(x2, ORDER), key, SIZES, 0.05, 0.1, clip[0], clip[1])
    assert np.array_equal(got, ref)
    lit = qn.isoline_variation_tree([N(l) for l in tu.tree_leaves(x1)], [N(l) for l in tu.tree_leaves(x2)], key, 0.05, 0.1, clip[0], clip[1])
    for a, b in zip(tu.tree_leaves(out), lit):
        assert a.shape == b.shape and np.allclose(N(a), b, rtol=1e-6, atol=1e-6)
    # a tree with one leaf draws exactly what the single-tensor call draws
    one = isoline_variation({"w": x1["params"]["Dense_0"]["kernel"]}, {"w": x2["params"]["Dense_0"]["kernel"]}, key, 0.05, 0.1)
    flat = isoline_variation(x1["params"]["Dense_0"]["kernel"].reshape(B, -1).contiguous(), x2["params"]["Dense_0"]["kernel"].reshape(B, -1).contiguous(), key, 0.05, 0.1)
    assert np.array_equal(N(one["w"]).reshape(B, -1), N(flat))


def test_polynomial_operators_pytree(dev, co):
    from qdax_b200.core.emitters.mutation_operators import polynomial_crossover, polynomial_mutation

    rng = np.random.default_rng(2)
    B = 19
    x1 = {"a": T(rng.random((B, 12)), dev), "b": T(rng.random((B, 7)), dev)}
    x2 = {"a": T(rng.random((B, 12)), dev), "b": T(rng.random((B, 7)), dev)}
    key = jr.key(5)
    m = polynomial_mutation(x1, key, proportion_to_mutate=0.5, eta=0.05, minval=0.0, maxval=1.0)
    c = polynomial_crossover(x1, x2, key, proportion_var_to_change=0.5)
    for k in ("a", "b"):                     # the same per-individual keys for every leaf (mutation_operators.py:107-116, :164-171)
        assert np.array_equal(N(m[k]), co.polynomial_mutation(N(x1[k]), key, 0.5, 0.05, 0.0, 1.0))
        assert np.array_equal(N(c[k]), co.polynomial_crossover(N(x1[k]), N(x2[k]), key, 0.5))


def _scoring(tree, key):
    """A user-supplied Python scoring function on the leaves (the generic path of MAPElites)."""
    p = tree["params"]
    w0, b0 = p["Dense_0"]["kernel"], p["Dense_0"]["bias"]
    fit = -(w0 * w0).sum(dim=(1, 2)) - (b0 * b0).sum(dim=1)
    desc = torch.stack([b0[:, 0], p["Dense_1"]["bias"][:, 1]], dim=1).contiguous()
    return fit, desc, {}


@pytest.mark.parametrize("shapes,order,sizes", [(SHAPES, ORDER, SIZES), (SHAPES4, ORDER4, SIZES4)], ids=["generic-emit-53", "fused-emit-56"])
def test_map_elites_with_pytree_genotypes(dev, co, shapes, order, sizes):
    """MAPElites.init / update on a pytree genotype with a Python scoring function: offspring bit-exact against the oracle's
    emit (same key chain, map_elites.py:177,241; standard_emitters.py:55; mutation_operators.py:205,220), insertion
    bit-exact on the injected identical (fitness, descriptor) values, repertoire leaves = the oracle's packed rows."""
    from qdax_b200 import random as qr
    from qdax_b200 import tree_util as tu
    from qdax_b200.core.containers.mapelites_repertoire import MapElitesRepertoire, compute_euclidean_centroids
    from qdax_b200.core.emitters.mutation_operators import isoline_variation
    from qdax_b200.core.emitters.standard_emitters import MixingEmitter
    from qdax_b200.core.map_elites import MAPElites
    from qdax_b200.utils.metrics import default_qd_metrics

    rng = np.random.default_rng(3)
    B, D = 64, sum(sizes)
    init = make_tree(rng, 40, shapes, dev)
    cent = compute_euclidean_centroids((8, 8), 0.0, 1.0, device=dev)
    emitter = MixingEmitter(lambda x, k: x, functools.partial(isoline_variation, iso_sigma=0.05, line_sigma=0.1, minval=0.0, maxval=1.0), 1.0, B)
    me = MAPElites(_scoring, emitter, functools.partial(default_qd_metrics, qd_offset=0.0))
    rep, state, _ = me.init(init, cent, jr.key(0))
    assert isinstance(rep, MapElitesRepertoire) and tu.is_tree(rep.genotypes)

    cent_h = N(cent)
    K = cent_h.shape[0]
    f0, d0, _ = _scoring(init, None)
    g, f, d, _ = co.add(np.zeros((K, D)), np.full(K, -np.inf), np.zeros((K, 2)), packed(init, order), N(f0), N(d0), co.cells(N(d0), cent_h))
    assert np.array_equal(packed(rep.genotypes, order), g) and np.array_equal(N(rep.fitnesses).ravel(), f)

    key = jr.key(7)
    for it in range(4):
        ks = qr.split(key)
        key, sub = ks[0], ks[1]
        old_rows = packed(rep.genotypes, order)
        rep2, state, metrics = me.update(rep, state, sub)
        assert np.array_equal(packed(rep.genotypes, order), old_rows)            # value semantics: the input is untouched
        # oracle: update :177 key, s1 = split(sub); ask :241 _, e = split(s1); emit(e)
        e = jr.split(jr.split(sub)[1])[1]
        x, p1, p2 = co.emit_isoline_leaves(g, f, e, B, sizes, 0.05, 0.1, 0.0, 1.0)
        off_tree = tu.unpack(T(x, dev), tu.spec_of(init))
        fo, do, _ = _scoring(off_tree, None)
        g, f, d, _ = co.add(g, f, d, x, N(fo), N(do), co.cells(N(do), cent_h))
        assert np.array_equal(packed(rep2.genotypes, order), g), f"genotypes differ at iteration {it}"
        assert np.array_equal(N(rep2.fitnesses).ravel(), f) and np.array_equal(N(rep2.descriptors), d)
        assert np.allclose(float(metrics["qd_score"]), co.metrics(f, 0.0)[0], rtol=1e-5)
        rep = rep2
    # select() gathers every leaf with the same indices (uniform_selector.py:57-60)
    sel = rep.select(jr.key(9), 10)
    idx = co.select_indices(f, jr.key(9), 10)
    assert np.array_equal(packed(sel.genotypes, order), g[idx]) and np.array_equal(N(sel.fitnesses).ravel(), f[idx])
