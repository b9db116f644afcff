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


def noisy_arm_scoring_function(params: torch.Tensor, key, fit_variance: float, desc_variance: float,
                               params_variance: float) -> Tuple[torch.Tensor, torch.Tensor, dict]:
    """reference arm.py:53-81: key, f_subkey, d_subkey, p_subkey = split(key, 4); Gaussian noise on the parameters before
    the arm is evaluated, on the fitnesses and on the descriptors afterwards.  One kernel (qdx_score_noisy_arm)."""
    fitnesses, descriptors = _native.score_noisy_arm(params, key, float(fit_variance), float(desc_variance), float(params_variance))
    return fitnesses, descriptors, {}
