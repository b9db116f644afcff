"""Unstructured repertoire (AURORA-style archive with an l-value) -- mirrors qdax/core/containers/unstructured_repertoire.py:
132-440 of the reference for the insertion rule: genotypes (max_size, D), fitnesses (max_size, 1) with
-inf for empty slots, descriptors (max_size, Dd), l_value.  `add` = nearest / second-nearest stored individual + l-value
tests + intra-batch competition + segment_max + scatter, executed by qdx_unstructured_plan -> qdx_gather_rows ->
qdx_unstructured_offer -> qdx_commit (the same packed-key election and commit as MapElitesRepertoire.add).

PARITY UNPINNED (no reference test fixes a value; jax is not installable here): oracle/qdax_containers_numpy.py is the literal
restatement; its header spells out the (N, N, Dd) broadcast of `filtered_descriptors` and the other places where the source
is restated as written rather than as probably intended."""

from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional, Tuple

import torch

from qdax_b200 import _native
from qdax_b200.core.containers.ga_repertoire import GARepertoire
from qdax_b200.core.containers.mapelites_repertoire import TIE_BREAKS


class UnstructuredRepertoire(GARepertoire):
    def __init__(self, genotypes, fitnesses, descriptors, l_value, max_size: int, extra_scores: Optional[Dict[str, Any]] = None,
                 keys_extra_scores: Tuple[str, ...] = (), tie_break: str = "first"):
        super().__init__(genotypes, fitnesses, extra_scores, keys_extra_scores)
        self.descriptors = descriptors
        self.l_value = l_value
        self.max_size = int(max_size)
        if tie_break not in TIE_BREAKS:
            raise ValueError(f"tie_break must be one of {TIE_BREAKS}")
        self.tie_break = tie_break

    def get_maximal_size(self) -> int:
        """reference :155-157."""
        return self.max_size

    def get_number_genotypes(self) -> torch.Tensor:
        """reference :159-161."""
        return (self.fitnesses != float("-inf")).sum()

    def _clone_state(self) -> "UnstructuredRepertoire":
        return self.replace(genotypes=self.genotypes.clone(), fitnesses=self.fitnesses.clone(), descriptors=self.descriptors.clone())

    def _l_value(self) -> float:
        lv = self.l_value
        return float(lv.reshape(-1)[0]) if isinstance(lv, torch.Tensor) else float(lv)          # self.l_value[0] (:211)

    def add(self, batch_of_genotypes, batch_of_descriptors, batch_of_fitnesses, batch_of_extra_scores=None, *,
            _donate: bool = False) -> "UnstructuredRepertoire":
        """reference :162-337."""
        if self.keys_extra_scores:
            raise NotImplementedError("extra scores in the unstructured repertoire are outside the accelerated path")
        self._raise_if_error()
        g = _native.require_cuda(batch_of_genotypes, "batch_of_genotypes")
        d = _native.require_cuda(batch_of_descriptors, "batch_of_descriptors")
        f = _native.require_cuda(batch_of_fitnesses, "batch_of_fitnesses").reshape(-1)          # :184 reshape(-1, 1)
        B = g.shape[0]
        if f.numel() != B or d.shape[0] != B:
            raise ValueError("batch size mismatch between genotypes, descriptors and fitnesses")
        new = self if _donate else self._clone_state()
        if B == 0:
            return new
        N = new.max_size
        g2 = g.reshape(B, -1)
        rep_g = new.genotypes.reshape(N, -1)
        rep_f = new.fitnesses.reshape(-1)
        if rep_g.shape[1] != g2.shape[1]:
            raise ValueError("genotype dimension mismatch")
        dev = g.device
        l = new._l_value()
        nbytes = C.c_int64(0)
        _native.call("qdx_unstructured_scratch", C.c_int64(N), C.c_int64(B), C.byref(nbytes))
        scratch = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        order = torch.empty(B, dtype=torch.int32, device=dev)
        Dd = d.shape[1]
        _native.call("qdx_unstructured_plan", _native._ptr(rep_f), _native._ptr(new.descriptors), C.c_int64(N), C.c_int32(Dd), _native._ptr(f),
                     _native._ptr(d), C.c_int64(B), C.c_float(l), _native._ptr(scratch), _native._ptr(order), _native._stream())
        gs, ds = _native.gather_rows(g2, order), _native.gather_rows(d, order)                  # :252-263 re-indexing
        fs = _native.gather_rows(f.reshape(B, 1), order).reshape(B)
        ws = new._workspace()
        first = new.tie_break == "first"
        _native.call("qdx_unstructured_offer", _native._ptr(fs), _native._ptr(ds), C.c_int64(B), C.c_int32(Dd), C.c_int64(N), C.c_float(l),
                     _native._ptr(scratch), _native._ptr(order), ws.ptr, _native._ptr(rep_f), C.c_int32(first), _native._stream())
        _native.commit(ws, gs, fs, ds, rep_g, rep_f, new.descriptors, first_wins=first)         # :313-330
        if _native.DEBUG_SYNC:
            ws.check()
        return new

    @classmethod
    def init(cls, genotypes, fitnesses, descriptors, l_value, max_size: int, *args, extra_scores=None,
             keys_extra_scores: Tuple[str, ...] = (), tie_break: str = "first", **kwargs) -> "UnstructuredRepertoire":
        """reference :372-440: fitness -inf (max_size, 1), genotypes NaN, descriptors 0, then add the first batch."""
        genotypes = _native.require_cuda(genotypes, "genotypes")
        dev = genotypes.device
        rep = cls(
            genotypes=torch.full((max_size,) + tuple(genotypes.shape[1:]), float("nan"), dtype=torch.float32, device=dev),
            fitnesses=torch.full((max_size, 1), float("-inf"), dtype=torch.float32, device=dev),
            descriptors=torch.zeros((max_size, descriptors.shape[-1]), dtype=torch.float32, device=dev),
            l_value=l_value, max_size=max_size, extra_scores={}, keys_extra_scores=keys_extra_scores, tie_break=tie_break,
        )
        return rep.add(genotypes, descriptors, fitnesses, extra_scores, _donate=True)
