"""Pytree genotypes -- the slice of `jax.tree_util` that QDax's MAP-Elites path relies on
(/root/reference: qdax/core/emitters/mutation_operators.py:206,219-224; qdax/core/containers/mapelites_repertoire.py:
234-240,342-347; qdax/core/emitters/repertoire_selectors/uniform_selector.py:57-60), plus the packed row layout the
kernels work on.

A pytree is any nesting of dict / list / tuple (and None) whose leaves are CUDA float32 tensors with a common leading
batch dimension.  Leaves are enumerated in `jax.tree.leaves` order: dict entries by sorted key, sequences by position.

On the device an individual is ONE packed row: the concatenation of its flattened leaves in that order, so that parent
selection, isoline variation and the repertoire scatter each stay a single kernel over (N, D_total) rows whatever the
tree looks like.  `unpack` hands out the leaves as strided VIEWS of the packed buffer (tagged, so that `pack` of the same
leaves is free); `pack` of foreign leaves copies them with qdx_copy_2d."""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Callable, List, Sequence, Tuple

import numpy as np
import torch

MAX_LEAVES = 32


def is_leaf(x: Any) -> bool:
    return isinstance(x, torch.Tensor)


def is_tree(x: Any) -> bool:
    """True for a genotype that is a container of leaves rather than a single tensor."""
    return isinstance(x, (dict, list, tuple))


def tree_flatten(tree: Any) -> Tuple[List[torch.Tensor], Any]:
    """(leaves in jax.tree.leaves order, treedef)."""
    leaves: List[torch.Tensor] = []

    def rec(node):
        if node is None:
            return ("none",)
        if isinstance(node, dict):
            keys = sorted(node.keys())
            return ("dict", tuple(keys), tuple(rec(node[k]) for k in keys))
        if isinstance(node, (list, tuple)):
            kind = "list" if isinstance(node, list) else "tuple"
            if isinstance(node, tuple) and hasattr(node, "_fields"):      # namedtuple
                return ("namedtuple", type(node), tuple(rec(c) for c in node))
            return (kind, tuple(rec(c) for c in node))
        leaves.append(node)
        return ("leaf",)

    return leaves, rec(tree)


def tree_unflatten(treedef: Any, leaves: Sequence[Any]) -> Any:
    it = iter(leaves)

    def rec(d):
        tag = d[0]
        if tag == "none":
            return None
        if tag == "leaf":
            return next(it)
        if tag == "dict":
            return {k: rec(c) for k, c in zip(d[1], d[2])}
        if tag == "namedtuple":
            return d[1](*[rec(c) for c in d[2]])
        seq = [rec(c) for c in d[1]]
        return seq if tag == "list" else tuple(seq)

    out = rec(treedef)
    return out


def tree_leaves(tree: Any) -> List[torch.Tensor]:
    return tree_flatten(tree)[0]


def tree_structure(tree: Any) -> Any:
    return tree_flatten(tree)[1]


def tree_map(fn: Callable, tree: Any, *rest: Any) -> Any:
    leaves, treedef = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    for o in others:
        if len(o) != len(leaves):
            raise ValueError("tree_map: trees do not have the same structure")
    return tree_unflatten(treedef, [fn(*xs) for xs in zip(leaves, *others)])


@dataclass(frozen=True)
class TreeSpec:
    """Structure of a packed genotype: treedef, per-individual leaf shapes, leaf offsets inside the packed row."""

    treedef: Any
    shapes: Tuple[Tuple[int, ...], ...]
    offsets: Tuple[int, ...]          # len = n_leaves + 1; offsets[-1] = D_total

    @property
    def sizes(self) -> Tuple[int, ...]:
        return tuple(self.offsets[i + 1] - self.offsets[i] for i in range(len(self.shapes)))

    @property
    def total(self) -> int:
        return self.offsets[-1]

    @property
    def n_leaves(self) -> int:
        return len(self.shapes)


def spec_of(tree: Any, batched: bool = True) -> TreeSpec:
    leaves, treedef = tree_flatten(tree)
    if not leaves:
        raise ValueError("genotype pytree has no leaves")
    shapes = tuple(tuple(l.shape[1:] if batched else l.shape) for l in leaves)
    off = [0]
    for s in shapes:
        off.append(off[-1] + int(np.prod(s)) if len(s) else off[-1] + 1)
    return TreeSpec(treedef, shapes, tuple(off))


def unpack(flat: torch.Tensor, spec: TreeSpec) -> Any:
    """Leaves as views of the packed (N, D_total) buffer, tagged for a zero-copy `pack`."""
    N = flat.shape[0]
    leaves = []
    for l, shape in enumerate(spec.shapes):
        o, sz = spec.offsets[l], spec.offsets[l + 1] - spec.offsets[l]
        strides = [1] * len(shape)
        for i in range(len(shape) - 2, -1, -1):
            strides[i] = strides[i + 1] * shape[i + 1]
        v = flat.as_strided((N,) + tuple(shape), (flat.stride(0),) + tuple(strides), flat.storage_offset() + o) if sz > 0 \
            else flat.new_empty((N,) + tuple(shape))
        v._qdx_pack = (flat, o)
        leaves.append(v)
    return tree_unflatten(spec.treedef, leaves)


def pack(tree: Any, spec: TreeSpec = None) -> Tuple[torch.Tensor, TreeSpec]:
    """(packed (N, D_total) float32 tensor, spec).  Free when the leaves are the views `unpack` handed out (views always
    show the buffer's current content); otherwise each leaf is copied into place (qdx_copy_2d)."""
    from qdax_b200 import _native

    leaves, treedef = tree_flatten(tree)
    if spec is None:
        spec = spec_of(tree)
    elif treedef != spec.treedef or tuple(tuple(l.shape[1:]) for l in leaves) != spec.shapes:
        raise ValueError("genotype pytree does not match the repertoire's structure")
    N = leaves[0].shape[0]
    for l in leaves:
        _native.require_cuda(l, "genotype leaf")
        if l.shape[0] != N:
            raise ValueError("genotype leaves disagree on the batch size")
    tag0 = getattr(leaves[0], "_qdx_pack", None)
    if tag0 is not None:
        flat = tag0[0]
        if flat.shape == (N, spec.total) and all(
                (t := getattr(l, "_qdx_pack", None)) is not None and t[0] is flat and t[1] == spec.offsets[i]
                for i, l in enumerate(leaves)):
            return flat, spec
    flat = torch.empty((N, spec.total), dtype=torch.float32, device=leaves[0].device)
    for i, l in enumerate(leaves):
        sz = spec.offsets[i + 1] - spec.offsets[i]
        if sz == 0 or N == 0:
            continue
        src = l.reshape(N, sz)
        if src.stride(1) != 1:
            src = src.contiguous()
        _native.call("qdx_copy_2d", C.c_void_p(src.data_ptr()), C.c_int64(src.stride(0) if N > 1 else sz),
                     C.c_void_p(flat.data_ptr() + 4 * spec.offsets[i]), C.c_int64(spec.total), C.c_int64(N), C.c_int64(sz),
                     _native._stream())
    return flat, spec


def is_packed_view(tree: Any, flat: torch.Tensor) -> bool:
    """True when every leaf of `tree` is a view `unpack` handed out of the packed buffer `flat` (so writes into `flat` are
    writes into the tree)."""
    return all((t := getattr(l, "_qdx_pack", None)) is not None and t[0] is flat for l in tree_leaves(tree))


def leaf_table(spec: TreeSpec, keys: np.ndarray):
    """ctypes qdx_leaf_table for the kernels: offsets + the per-leaf noise keys split(key', n_leaves)."""
    from qdax_b200._lib import LeafTable

    if spec.n_leaves > MAX_LEAVES:
        raise NotImplementedError(f"genotype pytrees with more than {MAX_LEAVES} leaves: ravel groups of leaves on the host side")
    lt = LeafTable()
    lt.n = spec.n_leaves
    for i, o in enumerate(spec.offsets):
        lt.off[i] = o
    k = np.asarray(keys, dtype=np.uint32).reshape(spec.n_leaves, 2)
    for i in range(spec.n_leaves):
        lt.key[2 * i], lt.key[2 * i + 1] = int(k[i, 0]), int(k[i, 1])
    return lt
