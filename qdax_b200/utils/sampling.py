"""Multi-sample evaluation -- mirrors qdax/utils/sampling.py:111-152 of the reference (multi_sample_scoring_function).

The reference vmaps the scoring function over `num_samples` keys (out_axes = 1); here the scoring function is called once per
key -- each call is the native scoring kernel (or the user's function) -- and the results are stacked on axis 1."""

from __future__ import annotations

from typing import Any, Callable, Dict, Tuple

import torch

from qdax_b200 import random as qrandom


def multi_sample_scoring_function(policies_params, key, scoring_fn: Callable, num_samples: int) -> Tuple[torch.Tensor, torch.Tensor, Dict[str, Any]]:
    """(n, num_samples) fitnesses, (n, num_samples, num_descriptors) descriptors, extra scores with the same extra axis;
    sample s is evaluated with jax.random.split(key, num_samples)[s] (reference :137-150)."""
    keys = qrandom.split(key, num_samples)                                   # :137
    outs = [scoring_fn(policies_params, keys[s]) for s in range(num_samples)]
    fitnesses = torch.stack([o[0] for o in outs], dim=1)
    descriptors = torch.stack([o[1] for o in outs], dim=1)
    extra: Dict[str, Any] = {}
    for name in (outs[0][2] or {}):
        vals = [o[2][name] for o in outs]
        extra[name] = torch.stack(vals, dim=1) if isinstance(vals[0], torch.Tensor) else vals
    return fitnesses, descriptors, extra
