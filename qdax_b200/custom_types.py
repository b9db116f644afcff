"""Type aliases mirroring qdax/custom_types.py (reference :1-86), with torch.Tensor in place of jax.Array."""
from typing import Any, Dict, Union

import numpy as np
import torch

Array = torch.Tensor
Genotype = Any  # single float32 (B, D) tensor on the fast path; pytrees are flattened host-side
Fitness = torch.Tensor
Descriptor = torch.Tensor
Centroid = torch.Tensor
ExtraScores = Dict[str, Any]
Metrics = Dict[str, torch.Tensor]
RNGKey = Union[np.ndarray, "torch.Tensor"]  # two uint32 words (jax.random.key_data layout)
Params = Any
Observation = torch.Tensor
Action = torch.Tensor
Reward = torch.Tensor
Done = torch.Tensor
EnvState = Any
Mask = torch.Tensor
