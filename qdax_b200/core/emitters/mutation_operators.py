"""Variation operators -- mirrors qdax/core/emitters/mutation_operators.py of the reference.

isoline_variation (:175-226), polynomial_mutation (:81-117) and polynomial_crossover (:139-172) run as hand-written
kernels driven by the exact jax.random streams (Threefry-2x32, counter-based)."""

from __future__ import annotations

from typing import Optional

import torch

from qdax_b200 import _native


def isoline_variation(x1: torch.Tensor, x2: torch.Tensor, key, iso_sigma: float, line_sigma: float,
                      minval: Optional[float] = None, maxval: Optional[float] = None) -> torch.Tensor:
    """Iso+Line-DD variation (reference :175-226): x = (x1 + N(0, iso)) + (x2 - x1) * N(0, line)[:, None], clipped
    when either bound is given.  Single-leaf float32 genotypes of shape (batch, ...)."""
    if isinstance(x1, (dict, list, tuple)):
        raise NotImplementedError("pytree genotypes are a SURVEY 8(f) 'next' row; pass a single (batch, D) tensor")
    return _native.isoline_variation(x1, x2, key, float(iso_sigma), float(line_sigma), minval, maxval)


def polynomial_mutation(x: torch.Tensor, key, proportion_to_mutate: float, eta: float, minval: float, maxval: float) -> torch.Tensor:
    """Polynomial mutation over a batch of genotypes (reference :81-117 / :12-78): per row, `int(proportion * D)` genes
    chosen without replacement (jax.random.choice = first entries of a random permutation) receive the polynomial
    perturbation driven by one uniform draw each; the row is clipped to [minval, maxval]."""
    if isinstance(x, (dict, list, tuple)):
        raise NotImplementedError("pytree genotypes are a SURVEY 8(f) 'next' row; pass a single (batch, D) tensor")
    return _native.polynomial_mutation(x, key, float(proportion_to_mutate), float(eta), float(minval), float(maxval))


def polynomial_crossover(x1: torch.Tensor, x2: torch.Tensor, key, proportion_var_to_change: float) -> torch.Tensor:
    """Crossover over pairs of genotypes (reference :139-172 / :120-136): per row, `int(proportion * D)` positions drawn
    WITH replacement (jax.random.randint) are copied from x2 into x1."""
    if isinstance(x1, (dict, list, tuple)):
        raise NotImplementedError("pytree genotypes are a SURVEY 8(f) 'next' row; pass a single (batch, D) tensor")
    return _native.polynomial_crossover(x1, x2, key, float(proportion_var_to_change))
