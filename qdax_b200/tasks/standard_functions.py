"""Rastrigin and sphere tasks -- mirrors qdax/tasks/standard_functions.py:9-48.  x = 10 p - 5;
rastrigin f = -(10 D + sum(x^2 - 10 cos(2 pi x))), sphere f = -sum(x^2); descriptor = (p[0], p[1]).
`desc_dim` other than 2 is a declared extension (descriptor = first desc_dim genes) used by the
high-dimensional CVT configuration of BASELINE.json; the reference always returns 2 descriptors."""

from __future__ import annotations

from typing import Tuple

import torch

from qdax_b200 import _native


def rastrigin_scoring_function(params: torch.Tensor, key=None, desc_dim: int = 2) -> Tuple[torch.Tensor, torch.Tensor, dict]:
    fitnesses, descriptors = _native.score("rastrigin", params, desc_dim)
    return fitnesses, descriptors, {}


def sphere_scoring_function(params: torch.Tensor, key=None, desc_dim: int = 2) -> Tuple[torch.Tensor, torch.Tensor, dict]:
    fitnesses, descriptors = _native.score("sphere", params, desc_dim)
    return fitnesses, descriptors, {}


def rastrigin(params: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    f, d = _native.score("rastrigin", params.reshape(1, -1), 2)
    return f[0], d[0]


def sphere(params: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    f, d = _native.score("sphere", params.reshape(1, -1), 2)
    return f[0], d[0]
