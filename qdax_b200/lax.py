"""`scan` shim: `qdax_b200.lax.scan(map_elites.scan_update, carry, (), length=n)` is how the reference's
`jax.lax.scan(map_elites.scan_update, ...)` idiom (tests/tasks_test/arm_test.py:100-109, examples) is written
here.  A bound `scan_update` of a driver that offers a native multi-generation path is dispatched to it;
anything else runs as a plain Python loop."""

from __future__ import annotations

from typing import Any, Callable, Optional

import torch


def scan(f: Callable, init: Any, xs: Any = (), length: Optional[int] = None, **kwargs):
    if length is None:
        length = len(xs)
    owner = getattr(f, "__self__", None)
    if owner is not None and getattr(f, "__name__", "") == "scan_update" and hasattr(owner, "scan"):
        return owner.scan(init, length, **kwargs)
    carry, ys = init, []
    for i in range(length):
        carry, y = f(carry, xs[i] if (xs is not None and len(xs)) else None)
        ys.append(y)
    if ys and isinstance(ys[0], dict):
        ys = {k: torch.stack([torch.as_tensor(y[k]) for y in ys]) for k in ys[0]}
    return carry, ys
