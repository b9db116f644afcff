"""Variation operators -- mirrors qdax/core/emitters/mutation_operators.py of the reference.

isoline_variation (:175-226) runs as a hand-written kernel (Threefry-2x32 counter-based normal draws, one
32-bit draw per gene, exactly jax.random.normal's stream).  polynomial_mutation / polynomial_crossover
(:81-172) are SURVEY.md section 8(f) "next" rows and are not built yet."""

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


def polynomial_mutation(x, key, proportion_to_mutate: float, eta: float, minval: float, maxval: float):
    raise NotImplementedError("polynomial_mutation (reference :81-117) is a SURVEY 8(f) 'next' row, not built in round 1")


def polynomial_crossover(x1, x2, key, proportion_var_to_change: float):
    raise NotImplementedError("polynomial_crossover (reference :139-172) is a SURVEY 8(f) 'next' row, not built in round 1")
