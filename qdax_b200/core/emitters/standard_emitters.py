"""MixingEmitter -- mirrors qdax/core/emitters/standard_emitters.py:13-90.

emit(): n_variation = int(batch * pct); k1, k2, kv = split(key, 3); x1 = select(k1), x2 = select(k2);
variation_fn(x1, x2, kv); mutation branch re-splits the SAME key (reference :65).  When the emitter is the
configuration every MAP-Elites example uses (variation only, isoline_variation, uniform selector) the two
selections, the row gathers and the variation run as ONE kernel (qdx_generate) without materialising x1 / x2.
"""

from __future__ import annotations

import functools
from typing import Callable, Optional, Tuple

import torch

from qdax_b200 import _native
from qdax_b200 import random as qrandom
from qdax_b200 import tree_util
from qdax_b200.core.containers.ga_repertoire import GARepertoire
from qdax_b200.core.emitters.emitter import Emitter, EmitterState
from qdax_b200.core.emitters.mutation_operators import isoline_variation
from qdax_b200.core.emitters.repertoire_selectors.selector import Selector
from qdax_b200.core.emitters.repertoire_selectors.uniform_selector import UniformSelector


def isoline_config(variation_fn) -> Optional[dict]:
    """Recognise functools.partial(isoline_variation, iso_sigma=..., line_sigma=..., [minval, maxval])."""
    if not isinstance(variation_fn, functools.partial) or variation_fn.func is not isoline_variation or variation_fn.args:
        return None
    kw = dict(variation_fn.keywords)
    if not {"iso_sigma", "line_sigma"} <= set(kw) or set(kw) - {"iso_sigma", "line_sigma", "minval", "maxval"}:
        return None
    return {"iso_sigma": float(kw["iso_sigma"]), "line_sigma": float(kw["line_sigma"]),
            "minval": kw.get("minval"), "maxval": kw.get("maxval")}


class MixingEmitter(Emitter):
    def __init__(self, mutation_fn: Callable, variation_fn: Callable, variation_percentage: float, batch_size: int,
                 selector: Optional[Selector] = None) -> None:
        self._mutation_fn = mutation_fn
        self._variation_fn = variation_fn
        self._variation_percentage = variation_percentage
        self._batch_size = batch_size
        self._selector = selector

    def _fused_isoline(self) -> Optional[dict]:
        """Config of the fused select+isoline kernel, or None when this emitter needs the generic path."""
        n_variation = int(self._batch_size * self._variation_percentage)
        if n_variation != self._batch_size:
            return None
        sel = self._selector
        if sel is not None and not (type(sel) is UniformSelector and sel.select_with_replacement):
            return None
        return isoline_config(self._variation_fn)

    def emit(self, repertoire: GARepertoire, emitter_state: Optional[EmitterState], key) -> Tuple[torch.Tensor, dict]:
        """reference :27-82."""
        cfg = self._fused_isoline()
        g = repertoire.genotypes
        if cfg is not None and isinstance(g, torch.Tensor) and g.is_cuda and g.dtype == torch.float32 \
                and (g.numel() // g.shape[0]) % 4 == 0 and repertoire.fitnesses.shape[-1] == 1:
            K = g.shape[0]
            ws = repertoire._workspace()
            _native.ensure_selection(repertoire.fitnesses.reshape(-1), ws)
            out = torch.empty((self._batch_size,) + tuple(g.shape[1:]), dtype=torch.float32, device=g.device)
            _native.generate(g.reshape(K, -1), repertoire.fitnesses.reshape(-1), None, ws, self._batch_size, cfg["iso_sigma"],
                             cfg["line_sigma"], cfg["minval"], cfg["maxval"], None, 1, None, False, 0, True, out, None, None,
                             gen_keys=_native.host_generation_keys(_native.KEYMODE_EMIT, key))
            return out, {}

        if cfg is not None and tree_util.is_tree(g) and repertoire.fitnesses.shape[-1] == 1:
            # pytree genotype: same fused select x2 + isoline over the packed rows, per-leaf noise keys (:219-224)
            flat, spec = repertoire._packed_genotypes()
            if spec.total % 4 == 0 and spec.n_leaves <= tree_util.MAX_LEAVES:
                ws = repertoire._workspace()
                _native.ensure_selection(repertoire.fitnesses.reshape(-1), ws)
                kv = qrandom.split(key, 3)[2]                                            # :55
                leaves = tree_util.leaf_table(spec, qrandom.split(qrandom.split(kv)[0], spec.n_leaves))    # mutation_operators.py:205, :220
                out = torch.empty((self._batch_size, spec.total), dtype=torch.float32, device=flat.device)
                _native.generate_leaves(flat, repertoire.fitnesses.reshape(-1), ws, self._batch_size, cfg["iso_sigma"], cfg["line_sigma"],
                                        cfg["minval"], cfg["maxval"], out, _native.host_generation_keys(_native.KEYMODE_EMIT, key), leaves)
                return tree_util.unpack(out, spec), {}

        n_variation = int(self._batch_size * self._variation_percentage)
        n_mutation = self._batch_size - n_variation
        if n_variation > 0:
            k = qrandom.split(key, 3)                                                    # :55
            x1 = repertoire.select(k[0], n_variation, selector=self._selector).genotypes  # :56
            x2 = repertoire.select(k[1], n_variation, selector=self._selector).genotypes  # :59
            x_variation = self._variation_fn(x1, x2, k[2])                               # :62
        if n_mutation > 0:
            k = qrandom.split(key)                                                       # :65 (same key re-split)
            x1 = repertoire.select(k[0], n_mutation, selector=self._selector).genotypes
            x_mutation = self._mutation_fn(x1, k[1])                                     # :69
        if n_variation == 0:
            genotypes = x_mutation
        elif n_mutation == 0:
            genotypes = x_variation
        else:
            if tree_util.is_tree(x_variation):                                           # :75-80
                genotypes = tree_util.tree_map(lambda a, b: torch.cat([a, b], dim=0), x_variation, x_mutation)
            else:
                genotypes = torch.cat([x_variation, x_mutation], dim=0)
        return genotypes, {}

    @property
    def batch_size(self) -> int:
        return self._batch_size
