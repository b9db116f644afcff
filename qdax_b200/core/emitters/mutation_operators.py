"""Variation operators -- mirrors qdax/core/emitters/mutation_operators.py of the reference.

isoline_variation (:175-226), polynomial_mutation (:81-117) and polynomial_crossover (:139-172) run as hand-written
kernels driven by the exact jax.random streams (Threefry-2x32, counter-based)."""

from __future__ import annotations

from typing import Optional

import torch

from qdax_b200 import _native
from qdax_b200 import random as qrandom
from qdax_b200 import tree_util


def isoline_variation(x1: torch.Tensor, x2: torch.Tensor, key, iso_sigma: float, line_sigma: float,
                      minval: Optional[float] = None, maxval: Optional[float] = None) -> torch.Tensor:
    """Iso+Line-DD variation (reference :175-226): x = (x1 + N(0, iso)) + (x2 - x1) * N(0, line)[:, None], clipped
    when either bound is given.  A single float32 tensor (batch, ...) or a pytree of such leaves: the line noise is
    shared, leaf l draws its iso noise from split(key', n_leaves)[l] (:219-224); the leaves are processed as packed rows
    by one kernel."""
    if tree_util.is_tree(x1):
        f1, spec = tree_util.pack(x1)
        f2, _ = tree_util.pack(x2, spec)
        ks = qrandom.split(key)                                       # :205
        leaves = tree_util.leaf_table(spec, qrandom.split(ks[0], spec.n_leaves))     # :220
        out = _native.isoline_variation_leaves(f1, f2, ks[1], leaves, float(iso_sigma), float(line_sigma), minval, maxval)
        return tree_util.unpack(out, spec)
    return _native.isoline_variation(x1, x2, key, float(iso_sigma), float(line_sigma), minval, maxval)


def polynomial_mutation(x: torch.Tensor, key, proportion_to_mutate: float, eta: float, minval: float, maxval: float) -> torch.Tensor:
    """Polynomial mutation over a batch of genotypes (reference :81-117 / :12-78): per row, `int(proportion * D)` genes
    chosen without replacement (jax.random.choice = first entries of a random permutation) receive the polynomial
    perturbation driven by one uniform draw each; the row is clipped to [minval, maxval].  On a pytree every leaf is
    mutated with the SAME per-individual keys (:107-116: tree.map over the leaves with one `mutation_keys`); leaves must
    be (batch, n) -- for higher-rank leaves the reference mutates whole sub-arrays along the first axis, not built."""
    if tree_util.is_tree(x):
        def one(leaf):
            if leaf.dim() != 2:
                raise NotImplementedError("polynomial_mutation on pytree leaves of rank > 2")
            return _native.polynomial_mutation(leaf, key, float(proportion_to_mutate), float(eta), float(minval), float(maxval))
        return tree_util.tree_map(one, x)
    return _native.polynomial_mutation(x, key, float(proportion_to_mutate), float(eta), float(minval), float(maxval))


def polynomial_crossover(x1: torch.Tensor, x2: torch.Tensor, key, proportion_var_to_change: float) -> torch.Tensor:
    """Crossover over pairs of genotypes (reference :139-172 / :120-136): per row, `int(proportion * D)` positions drawn
    WITH replacement (jax.random.randint) are copied from x2 into x1.  Pytrees: leaf by leaf with the same keys (:164-171),
    (batch, n) leaves only."""
    if tree_util.is_tree(x1):
        def one(a, b):
            if a.dim() != 2:
                raise NotImplementedError("polynomial_crossover on pytree leaves of rank > 2")
            return _native.polynomial_crossover(a, b, key, float(proportion_var_to_change))
        return tree_util.tree_map(one, x1, x2)
    return _native.polynomial_crossover(x1, x2, key, float(proportion_var_to_change))
