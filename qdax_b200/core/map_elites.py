"""MAP-Elites driver -- mirrors qdax/core/map_elites.py:22-286 of the reference: init, init_ask_tell, update,
scan_update, ask, tell, with the same key-splitting order (the key chain is observable behaviour).

Two execution paths behind the same methods:

  fused    emitter = MixingEmitter(variation only, isoline_variation, uniform selector), scoring function in
           {arm, rastrigin, sphere}, metrics = partial(default_qd_metrics, qd_offset=...), MapElitesRepertoire with
           a single float32 leaf: one generation = prepare -> generate(+score+cell+offer) [-> cells] -> commit, all
           in libqdx.so, keys derived on the device, metrics written by the commit kernel.
  generic  anything else: emitter.emit / scoring_function / repertoire.add / metrics_function are called exactly
           as the reference does (user-supplied Python emitters and scoring functions keep working; the
           repertoire insertion still runs natively).

`scan(carry, length)` is the equivalent of `jax.lax.scan(map_elites.scan_update, carry, (), length)`.
"""

from __future__ import annotations

import functools
from typing import Any, Callable, Dict, Optional, Tuple

import numpy as np
import torch

from qdax_b200 import _native
from qdax_b200 import random as qrandom
from qdax_b200.tasks.arm import arm_scoring_function
from qdax_b200.tasks.standard_functions import rastrigin_scoring_function, sphere_scoring_function
from qdax_b200.utils.metrics import default_qd_metrics
from qdax_b200.core.containers.mapelites_repertoire import MapElitesRepertoire
from qdax_b200.core.emitters.emitter import Emitter, EmitterState
from qdax_b200.core.emitters.standard_emitters import MixingEmitter

_TASKS = {arm_scoring_function: "arm", rastrigin_scoring_function: "rastrigin", sphere_scoring_function: "sphere"}


def _task_of(scoring_function) -> Optional[Tuple[str, int]]:
    """(task name, descriptor dim) when the scoring function is one of the natively fused tasks."""
    fn, desc_dim = scoring_function, 2
    if isinstance(fn, functools.partial) and not fn.args and set(fn.keywords) <= {"desc_dim"}:
        desc_dim = int(fn.keywords.get("desc_dim", 2))
        fn = fn.func
    name = _TASKS.get(fn)
    if name is None or (name == "arm" and desc_dim != 2):
        return None
    return name, desc_dim


def _qd_offset_of(metrics_function) -> Optional[float]:
    if isinstance(metrics_function, functools.partial) and metrics_function.func is default_qd_metrics \
            and not metrics_function.args and set(metrics_function.keywords) == {"qd_offset"}:
        return float(metrics_function.keywords["qd_offset"])
    return None


def _metrics_destination(metrics_out: Optional[torch.Tensor], device) -> torch.Tensor:
    """Where the commit kernel writes (qd_score, max_fitness, coverage, inserted): a fresh device tensor, or the caller's -- a
    device tensor, or a PINNED host tensor, which the kernel then writes straight over PCIe (the device -> host transfer of a
    logging loop without a copy node in the stream: order the read with an event recorded after the call)."""
    if metrics_out is None:
        return torch.empty(4, dtype=torch.float32, device=device)
    if metrics_out.dtype != torch.float32 or metrics_out.numel() != 4 or not metrics_out.is_contiguous():
        raise ValueError("metrics_out must be a contiguous float32 tensor of 4 elements")
    if not (metrics_out.is_cuda and metrics_out.device == device) and not (metrics_out.device.type == "cpu" and metrics_out.is_pinned()):
        raise ValueError("metrics_out must live on the repertoire's device or in pinned host memory")
    return metrics_out


class MAPElites:
    """Core elements of the MAP-Elites algorithm (reference :22-55)."""

    def __init__(self, scoring_function: Optional[Callable], emitter: Emitter, metrics_function: Callable,
                 repertoire_init: Callable = MapElitesRepertoire.init) -> None:
        self._scoring_function = scoring_function
        self._emitter = emitter
        self._metrics_function = metrics_function
        self._repertoire_init = repertoire_init
        self._buffers: Dict[Tuple, Dict[str, torch.Tensor]] = {}
        self._timeline: Optional[list] = None   # bench.py: [(label, cuda event)] recorded around each kernel
        self._step_cache = None                 # (identity of the buffers, _native.GenerationStep, cfg)
        self._cfg_cache = None                  # (repertoire object, fused configuration)

    # ------------------------------------------------------------------------------------------ fused path
    def _fused_config(self, repertoire) -> Optional[dict]:
        """Configuration of the fused native path for this repertoire, or None (generic path).  Cached per repertoire OBJECT:
        repertoires are values (`replace` returns a new object), and an in-place (donated) update keeps every property this
        decision depends on."""
        c = self._cfg_cache
        if c is not None and c[0] is repertoire:
            return c[1]
        cfg = self._fused_config_uncached(repertoire)
        self._cfg_cache = (repertoire, cfg)
        return cfg

    def _fused_config_uncached(self, repertoire) -> Optional[dict]:
        if self._scoring_function is None or not isinstance(self._emitter, MixingEmitter) or type(repertoire) is not MapElitesRepertoire:
            return None
        iso = self._emitter._fused_isoline()
        task = _task_of(self._scoring_function)
        qd_offset = _qd_offset_of(self._metrics_function)
        g = repertoire.genotypes
        if iso is None or task is None or qd_offset is None or not isinstance(g, torch.Tensor):
            return None
        if not g.is_cuda or g.dtype != torch.float32 or g.dim() != 2 or g.shape[1] % 4 != 0:
            return None
        if repertoire.keys_extra_scores or repertoire.fitnesses.shape[-1] != 1:
            return None
        D = g.shape[1]
        if task[0] == "arm" and D > 128:
            return None
        if task[1] > min(D, 128) or repertoire.centroids.shape[1] != task[1]:
            return None
        return {**iso, "task": task[0], "desc_dim": task[1], "qd_offset": qd_offset}

    def _offspring_buffers(self, B: int, D: int, Dd: int, device) -> Dict[str, torch.Tensor]:
        k = (B, D, Dd, str(device))
        if k not in self._buffers:
            self._buffers[k] = {
                "g": torch.empty((B, D), dtype=torch.float32, device=device),
                "f": torch.empty((B,), dtype=torch.float32, device=device),
                "d": torch.empty((B, Dd), dtype=torch.float32, device=device),
                "c": torch.empty((B,), dtype=torch.int32, device=device),
            }
        return self._buffers[k]

    def _mark(self, label: str) -> None:
        if self._timeline is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._timeline.append((label, e))

    def _fused_generation(self, rep: MapElitesRepertoire, cfg: dict, key_mode: int, key, metrics_out: torch.Tensor,
                          carry: Optional[np.ndarray] = None) -> None:
        """One generation in place on `rep` (launch-only; no host synchronisation).  Steady state = two launches:
        generate (select + variation + scoring + cell + offer) and commit (whose service CTA also leaves the next
        generation's parent-selection tables in the workspace); the jax.random.split chain runs on the host.  The whole
        generation is enqueued by ONE C-ABI call (qdx_map_elites_step); the kernel-by-kernel path below is kept for the
        instrumented pass of bench.py (a CUDA event after every kernel)."""
        ws = rep._workspace()
        ws.raise_if_error()                     # sticky device error of an earlier generation (async mirror, never blocks)
        rep_f = rep.fitnesses.reshape(-1)
        if self._timeline is None:
            step = self._generation_step(rep, cfg, ws, rep_f)
            _native.ensure_selection(rep_f, ws)
            step.run(key_mode, key, carry, metrics_out)
            return
        K, D = rep.genotypes.shape
        B = self._emitter.batch_size
        buf = self._offspring_buffers(B, D, cfg["desc_dim"], rep.genotypes.device)
        grid = rep._grid()
        first = rep.tie_break == "first"
        gen_keys = _native.host_generation_keys(key_mode, key, carry)
        self._mark("begin")
        _native.ensure_selection(rep_f, ws)
        self._mark("prepare")
        index = None if grid is not None else _native.cvt_index_of(rep.centroids)   # low-dimensional CVT: bucket index
        fused_cells = grid is not None or index is not None
        _native.generate(rep.genotypes, rep_f, rep.centroids, ws, B, cfg["iso_sigma"], cfg["line_sigma"], cfg["minval"],
                         cfg["maxval"], cfg["task"], cfg["desc_dim"], grid, fused_cells, 0, first,
                         buf["g"], buf["f"], buf["d"], buf["c"], gen_keys=gen_keys, index=index, fired_rows_only=True)
        self._mark("generate")
        if not fused_cells:
            _native.cells(buf["d"], rep.centroids, None, ws, rep_f, buf["f"], offer=True, first_wins=first, out=buf["c"])
            self._mark("cells")
        _native.commit(ws, buf["g"], buf["f"], buf["d"], rep.genotypes, rep_f, rep.descriptors, first_wins=first,
                       qd_offset=cfg["qd_offset"], metrics_out=metrics_out)
        self._mark("commit")

    def _generation_step(self, rep: MapElitesRepertoire, cfg: dict, ws, rep_f, rank: int = 0, nranks: int = 1):
        """The filled qdx_step_desc of this (repertoire buffers, configuration), cached while the repertoire is updated in place."""
        ident = (rep.genotypes.data_ptr(), rep_f.data_ptr(), rep.descriptors.data_ptr(), rep.centroids.data_ptr(), ws.buf.data_ptr(),
                 rep.tie_break, rank, nranks, id(cfg))
        cached = self._step_cache
        if cached is not None and cached[0] == ident:
            return cached[1]
        K, D = rep.genotypes.shape
        B = self._emitter.batch_size
        buf = self._offspring_buffers(B, D, cfg["desc_dim"], rep.genotypes.device)
        grid = rep._grid()
        index = None if grid is not None else _native.cvt_index_of(rep.centroids)
        step = _native.GenerationStep(rep.genotypes, rep_f, rep.descriptors, rep.centroids, ws, B, cfg, grid, index,
                                      rep.tie_break == "first", buf, rank, nranks)
        self._step_cache = (ident, step, cfg)
        return step

    @staticmethod
    def _metrics_dict(m: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {"qd_score": m[..., 0], "max_fitness": m[..., 1], "coverage": m[..., 2]}

    # ------------------------------------------------------------------------------------------ reference API
    def init(self, genotypes, centroids, key) -> Tuple[MapElitesRepertoire, Optional[EmitterState], Dict]:
        """reference :57-92."""
        if self._scoring_function is None:
            raise ValueError("Scoring function is not set.")
        ks = qrandom.split(key)                                             # :81
        key, subkey = ks[0], ks[1]
        fitnesses, descriptors, extra_scores = self._scoring_function(genotypes, subkey)
        return self.init_ask_tell(genotypes=genotypes, fitnesses=fitnesses, descriptors=descriptors, centroids=centroids,
                                  key=key, extra_scores=extra_scores)

    def init_ask_tell(self, genotypes, fitnesses, descriptors, centroids, key, extra_scores=None):
        """reference :94-146."""
        if extra_scores is None:
            extra_scores = {}
        repertoire = self._repertoire_init(genotypes, fitnesses, descriptors, centroids, extra_scores)
        ks = qrandom.split(key)                                             # :133
        key, subkey = ks[0], ks[1]
        emitter_state = self._emitter.init(key=subkey, repertoire=repertoire, genotypes=genotypes, fitnesses=fitnesses,
                                           descriptors=descriptors, extra_scores=extra_scores)
        metrics = self._metrics_function(repertoire)
        return repertoire, emitter_state, metrics

    def update(self, repertoire: MapElitesRepertoire, emitter_state: Optional[EmitterState], key, *, donate: bool = False,
               metrics_out: Optional[torch.Tensor] = None):
        """One MAP-Elites iteration (reference :148-195).  `donate=True` updates the repertoire's buffers in
        place (jax buffer donation); by default the input repertoire is left untouched.  `metrics_out` (fused path only):
        see `_metrics_destination`."""
        if self._scoring_function is None:
            raise ValueError("Scoring function is not set.")
        cfg = self._fused_config(repertoire)
        if cfg is not None:
            rep = repertoire if donate else repertoire._clone_state()
            m = _metrics_destination(metrics_out, rep.genotypes.device)
            self._fused_generation(rep, cfg, _native.KEYMODE_UPDATE, key, m)
            self._last_metrics = m          # (qd_score, max_fitness, coverage, inserted) as one device tensor
            return rep, emitter_state, self._metrics_dict(m)

        ks = qrandom.split(key)                                             # :177
        key, subkey = ks[0], ks[1]
        genotypes, extra_info = self.ask(repertoire, emitter_state, subkey)
        ks = qrandom.split(key)                                             # :181
        key, subkey = ks[0], ks[1]
        fitnesses, descriptors, extra_scores = self._scoring_function(genotypes, subkey)
        return self.tell(genotypes=genotypes, fitnesses=fitnesses, descriptors=descriptors, repertoire=repertoire,
                         emitter_state=emitter_state, extra_scores=extra_scores, extra_info=extra_info)

    def scan_update(self, carry, _: Any = None):
        """reference :197-225 (the body handed to jax.lax.scan)."""
        repertoire, emitter_state, key = carry
        ks = qrandom.split(key)                                             # :214
        key, subkey = ks[0], ks[1]
        repertoire, emitter_state, metrics = self.update(repertoire, emitter_state, subkey)
        return (repertoire, emitter_state, key), metrics

    def scan(self, carry, length: int, *, donate: bool = False, graph: bool = False):
        """Equivalent of `jax.lax.scan(self.scan_update, carry, (), length=length)`: returns
        ((repertoire, emitter_state, key), metrics) with every metric stacked to shape (length,).
        On the fused path the key chain runs on the device and nothing synchronises with the host until the
        final carry key is read back; `graph=True` replays the whole scan as one CUDA graph."""
        repertoire, emitter_state, key = carry
        cfg = self._fused_config(repertoire) if self._scoring_function is not None else None
        if cfg is None:
            out = []
            for _ in range(length):
                (repertoire, emitter_state, key), m = self.scan_update((repertoire, emitter_state, key))
                out.append(m)
            stacked = {k: torch.stack([torch.as_tensor(m[k]) for m in out]) for k in out[0]} if out else {}
            return (repertoire, emitter_state, key), stacked
        rep = repertoire if donate else repertoire._clone_state()
        ws = rep._workspace()
        carry = np.array(_native.key_words(key), dtype=np.uint32)      # the scan carry key, advanced on the host (:214)
        metrics = torch.empty((length, 4), dtype=torch.float32, device=rep.genotypes.device)
        if graph and length > 0:
            # warm-up outside the capture (module load, offspring buffers) on a throw-away copy of the state
            self._fused_generation(repertoire._clone_state(), cfg, _native.KEYMODE_SCAN, None, metrics[0], carry.copy())
            _native.ensure_selection(rep.fitnesses.reshape(-1), ws)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for it in range(length):
                    self._fused_generation(rep, cfg, _native.KEYMODE_SCAN, None, metrics[it], carry)
            g.replay()
        else:
            for it in range(length):
                self._fused_generation(rep, cfg, _native.KEYMODE_SCAN, None, metrics[it], carry)
        _, _, err = ws.read()
        if err != 0:
            from .._lib import QdxError
            raise QdxError("MAPElites.scan", err)
        return (rep, emitter_state, carry), self._metrics_dict(metrics)

    def ask(self, repertoire: MapElitesRepertoire, emitter_state: Optional[EmitterState], key):
        """reference :227-243."""
        ks = qrandom.split(key)                                             # :241
        key, subkey = ks[0], ks[1]
        return self._emitter.emit(repertoire, emitter_state, subkey)

    def tell(self, genotypes, fitnesses, descriptors, repertoire: MapElitesRepertoire, emitter_state: Optional[EmitterState],
             extra_scores=None, extra_info=None):
        """reference :245-286."""
        if extra_scores is None:
            extra_scores = {}
        if extra_info is None:
            extra_info = {}
        repertoire = repertoire.add(genotypes, descriptors, fitnesses, extra_scores)
        emitter_state = self._emitter.state_update(emitter_state=emitter_state, repertoire=repertoire, genotypes=genotypes,
                                                   fitnesses=fitnesses, descriptors=descriptors,
                                                   extra_scores={**extra_scores, **extra_info})
        metrics = self._metrics_function(repertoire)
        return repertoire, emitter_state, metrics
