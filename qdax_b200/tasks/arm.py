"""Planar arm task -- mirrors qdax/tasks/arm.py:9-50.  fitness = -std of the clipped joint angles,
descriptor = end-effector position from the running sum of joint angles (sequential float32, as the spec)."""

from __future__ import annotations

from typing import Tuple

import torch

from qdax_b200 import _native


def arm_scoring_function(params: torch.Tensor, key=None) -> Tuple[torch.Tensor, torch.Tensor, dict]:
    """reference arm.py:41-50: (fitnesses (B,), descriptors (B, 2), {}); the key is unused."""
    fitnesses, descriptors = _native.score("arm", params, 2)
    return fitnesses, descriptors, {}


def arm(params: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference arm.py:9-38 for ONE genotype of shape (D,)."""
    f, d = _native.score("arm", params.reshape(1, -1), 2)
    return f[0], d[0]
