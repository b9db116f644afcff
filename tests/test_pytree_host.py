"""CPU tests for pytree genotypes: leaf ordering / structure handling of qdax_b200.tree_util, the
leaf table handed to the kernels, and the two oracles (literal NumPy, exact-arithmetic C) against each other on
isoline_variation over a tree (reference qdax/core/emitters/mutation_operators.py:205-224)."""
import collections

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import jax_prng as jr  # noqa: E402
from oracle import qdax_numpy as qn  # noqa: E402


def test_tree_flatten_follows_jax_leaf_order():
    from qdax_b200 import tree_util as tu

    Pt = collections.namedtuple("Pt", ["y", "x"])
    t = {"b": torch.zeros(2, 3), "a": [torch.ones(2, 1), None, (torch.full((2, 2, 2), 2.0),)], "c": Pt(torch.full((2,), 3.0), torch.full((2, 4), 4.0))}
    leaves, treedef = tu.tree_flatten(t)
    # dict keys sorted ("a" < "b" < "c"), sequences and namedtuple fields by position, None contributes no leaf
    assert [float(l.reshape(-1)[0]) for l in leaves] == [1.0, 2.0, 0.0, 3.0, 4.0]
    back = tu.tree_unflatten(treedef, leaves)
    assert isinstance(back["a"], list) and back["a"][1] is None and isinstance(back["a"][2], tuple) and isinstance(back["c"], Pt)
    assert tu.tree_structure(back) == treedef
    doubled = tu.tree_map(lambda a, b: a + b, t, t)
    assert [float(l.reshape(-1)[0]) for l in tu.tree_leaves(doubled)] == [2.0, 4.0, 0.0, 6.0, 8.0]
    spec = tu.spec_of(t)
    assert spec.shapes == ((1,), (2, 2), (3,), (), (4,)) and spec.offsets == (0, 1, 5, 8, 9, 13) and spec.sizes == (1, 4, 3, 1, 4)
    assert tu.spec_of(tu.tree_map(lambda x: x[0], t), batched=False) == spec
    with pytest.raises(ValueError):
        tu.tree_map(lambda a, b: a, t, {"a": torch.zeros(1)})


def test_unpack_views_alias_the_packed_rows():
    from qdax_b200 import tree_util as tu

    spec = tu.spec_of({"w": torch.zeros(3, 2, 3), "b": torch.zeros(3, 2)})
    flat = torch.arange(24, dtype=torch.float32).reshape(3, 8)
    tree = tu.unpack(flat, spec)
    assert tree["b"].tolist() == [[0, 1], [8, 9], [16, 17]]
    assert tree["w"][1].tolist() == [[10, 11, 12], [13, 14, 15]]
    flat[2, 7] = -1.0
    assert float(tree["w"][2, 1, 2]) == -1.0          # a view, not a copy


def test_leaf_table_layout():
    from qdax_b200 import tree_util as tu

    spec = tu.spec_of({"w": torch.zeros(3, 2, 3), "b": torch.zeros(3, 2)})
    keys = jr.split(jr.key(1), 2)
    lt = tu.leaf_table(spec, keys)
    assert lt.n == 2 and list(lt.off[:3]) == [0, 2, 8] and list(lt.key[:4]) == [int(k) for k in keys.reshape(-1)]
    big = tu.spec_of([torch.zeros(1, 1) for _ in range(33)])
    with pytest.raises(NotImplementedError):
        tu.leaf_table(big, jr.split(jr.key(1), 33))


@pytest.mark.parametrize("clip", [(None, None), (0.0, 1.0)])
def test_oracles_agree_on_pytree_isoline(co, clip):
    rng = np.random.default_rng(0)
    B, shapes = 23, [(2, 3), (5,), (3, 2, 2), ()]
    sizes = [int(np.prod(s)) for s in shapes]
    x1 = [rng.random((B,) + s).astype(np.float32) for s in shapes]
    x2 = [rng.random((B,) + s).astype(np.float32) for s in shapes]
    key = jr.key(3)
    lit = qn.isoline_variation_tree(x1, x2, key, 0.05, 0.1, clip[0], clip[1])
    pk = lambda xs: np.concatenate([a.reshape(B, -1) for a in xs], axis=1)
    c = co.isoline_variation_leaves(pk(x1), pk(x2), key, sizes, 0.05, 0.1, clip[0], clip[1])
    assert np.allclose(c, pk(lit), rtol=1e-6, atol=1e-6)
    # one leaf == the single-leaf operator, and the line noise is shared: leaf 0 of the tree differs from a lone leaf only
    # through its key (split(key', 4)[0] vs split(key', 1)[0] are the same block in partitionable mode)
    assert np.array_equal(co.isoline_variation_leaves(pk(x1), pk(x2), key, [sum(sizes)], 0.05, 0.1), co.isoline_variation(pk(x1), pk(x2), key, 0.05, 0.1))
    lone = co.isoline_variation(x1[0].reshape(B, -1), x2[0].reshape(B, -1), key, 0.05, 0.1, clip[0], clip[1])
    assert np.array_equal(c[:, :sizes[0]], lone)
