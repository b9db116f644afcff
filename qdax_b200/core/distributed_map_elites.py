"""Distributed MAP-Elites -- mirrors qdax/core/distributed_map_elites.py:16-244 of the reference.

The reference replicates the repertoire on every device under jax.pmap, lets each device emit and score its own
shard with its own key, all-gathers (genotypes, fitnesses, descriptors) and applies the identical `add` everywhere.
Here: one process per GPU (torch.distributed, NCCL over NVLink), same sharding (global offspring index =
rank * B_dev + i), same per-rank key chain (`key, subkey = split(key)`; emit(subkey); :124-131), two exchanges:

  exchange="allgather"  reference-faithful: the offspring tuples (+ their cells, computed on the shard) are
                        all-gathered, every rank offers the full batch and commits from the gathered rows.
  exchange="regen"      nothing but 64-bit keys travels: each rank offers its shard into its local key table (global
                        indices), ONE all-reduce(max) of the K keys elects the per-cell winners and carries every
                        rank's generation keys in tail slots; every rank then REGENERATES the winners from (owner's
                        keys, local index) -- the RNG is counter-based and the repertoire replicated -- scores them
                        and commits.  No genotype crosses NVLink.
  exchange="p2p"        no collective library, the exchange is fused into the two kernels of a generation: an offer that
                        improves its cell's local best is max-merged by the offering thread straight into every peer's
                        key table (system-scope 64-bit atomicMax over NVLink; buffers mapped with cudaIpc,
                        double-buffered by generation parity) and its genotype row is left in the rank's offspring block
                        of the same buffer; the last CTA of the generate kernel raises an arrival flag in every peer; the
                        commit kernel acquire-spins on its local flags, then copies every elected winner straight out
                        of its OWNER's offspring block (NVLink loads) -- the reference's all_gather reduced to the rows
                        that change the repertoire.  generate -> commit: two launches (one C-ABI call), no host
                        involvement, no collective, no recomputation: every rank copies the same bits.
  exchange="winners"    only what can change the repertoire travels: each rank offers its shard into its local
                        64-bit key table, one all-reduce(max) of the K keys elects the global per-cell winners
                        (the global best of a cell is always a local best), winners' rows are merged through a
                        K-row staging buffer.  Same repertoire, bit for bit (tests/test_gpu_distributed.py).

Replicas stay bit-identical because the insertion is deterministic: integer atomicMax on packed keys with the
tie broken on the GLOBAL offspring index.
"""

from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Tuple

import torch

from qdax_b200 import _native, parallel
from qdax_b200 import random as qrandom
from qdax_b200.core.containers.mapelites_repertoire import MapElitesRepertoire
from qdax_b200.core.emitters.emitter import EmitterState
from qdax_b200.core.map_elites import MAPElites, _metrics_destination


class DistributedMAPElites(MAPElites):
    def __init__(self, *args, exchange: str = "allgather", group=None, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        if exchange not in ("allgather", "winners", "regen", "p2p"):
            raise ValueError("exchange must be 'allgather', 'winners', 'regen' or 'p2p'")
        self._exchange = exchange
        self._group = group
        self._dist_buffers: Dict[Tuple, Dict[str, torch.Tensor]] = {}
        self._xchg: Optional[_native.PeerExchange] = None
        self.exchange_fallback: Optional[str] = None

    # ------------------------------------------------------------------------------------------ reference API
    def init(self, genotypes, centroids, key) -> Tuple[MapElitesRepertoire, Optional[EmitterState], Dict]:
        """reference :17-90: score the local genotypes with `key` itself (no split, :47), gather, init."""
        if self._scoring_function is None:
            raise ValueError("Scoring function is not set.")
        fitnesses, descriptors, extra_scores = self._scoring_function(genotypes, key)
        g = parallel.all_gather_rows(genotypes, self._group)
        f = parallel.all_gather_rows(fitnesses, self._group)
        d = parallel.all_gather_rows(descriptors, self._group)
        repertoire = MapElitesRepertoire.init(genotypes=g, fitnesses=f, descriptors=d, centroids=centroids)
        emitter_state = self._emitter.init(key=key, repertoire=repertoire, genotypes=genotypes, fitnesses=fitnesses,
                                           descriptors=descriptors, extra_scores=extra_scores)
        emitter_state = self._emitter.state_update(emitter_state=emitter_state, repertoire=repertoire, genotypes=genotypes,
                                                   fitnesses=fitnesses, descriptors=descriptors, extra_scores=extra_scores)
        return repertoire, emitter_state, self._metrics_function(repertoire)

    def _gather_buffers(self, R: int, B: int, D: int, Dd: int, K: int, device, kind: str) -> Dict[str, torch.Tensor]:
        """kind "gather": receive buffers of the all-gather exchange; "stage": per-cell staging rows of winners / regen."""
        k = (kind, R, B, D, Dd, K, str(device))
        if k not in self._dist_buffers:
            f32, i32 = torch.float32, torch.int32
            if kind == "gather":
                self._dist_buffers[k] = {
                    "G": torch.empty((R * B, D), dtype=f32, device=device), "F": torch.empty((R * B,), dtype=f32, device=device),
                    "Dn": torch.empty((R * B, Dd), dtype=f32, device=device), "C": torch.empty((R * B,), dtype=i32, device=device)}
            else:   # staging rows by cell: [genotype | descriptor | fitness] per cell
                self._dist_buffers[k] = {"stage": torch.zeros((K, D + Dd + 1), dtype=f32, device=device)}
        return self._dist_buffers[k]

    def _fused_distributed_generation(self, rep: MapElitesRepertoire, cfg: dict, key_mode: int, key, metrics_out) -> None:
        rank, R = parallel.world(self._group)
        K, D = rep.genotypes.shape
        B = self._emitter.batch_size
        Dd = cfg["desc_dim"]
        dev = rep.genotypes.device
        buf = self._offspring_buffers(B, D, Dd, dev)
        ws = rep._workspace()
        rep_f = rep.fitnesses.reshape(-1)
        grid = rep._grid()
        first = rep.tie_break == "first"
        winners = self._exchange in ("winners", "regen", "p2p")
        p2p = self._exchange == "p2p" and R > 1
        base = rank * B
        if p2p:
            if self._xchg is None or self._xchg.K != K or self._xchg.shape != (B, D, Dd):
                try:
                    if self._xchg is not None:
                        self._xchg.close()
                    self._xchg = _native.PeerExchange(K, self._group, B, D, Dd)
                except _native.PeerExchangeUnavailable as e:     # same decision on every rank: NCCL carries the keys instead
                    self.exchange_fallback = f"p2p unavailable ({e}); using regen"
                    self._exchange, p2p = "regen", False
            if p2p:
                self._xchg.attach(ws)
        ws.raise_if_error()            # sticky device error of an earlier generation (e.g. QDX_ERR_PEER_TIMEOUT): host mirror, never blocks
        if p2p and self._timeline is None:
            # the whole generation (generate -> [cells -> publish] -> elect -> commit) behind ONE C-ABI call
            step = self._generation_step(rep, cfg, ws, rep_f, rank, R)
            _native.ensure_selection(rep_f, ws)
            step.run(key_mode, key, None, metrics_out)
            return
        gen_keys = _native.host_generation_keys(key_mode, key)
        self._mark("begin")
        if self._exchange == "regen" and R > 1:
            # the generation keys ride behind the key table in the all-reduce: prepare publishes them (and clears the
            # other ranks' slots), so this exchange keeps its prepare launch
            _native.select_prepare(rep_f, ws, key_mode, key, rank_slot=rank)
        else:
            _native.ensure_selection(rep_f, ws)
        self._mark("prepare")
        index = None if grid is not None else _native.cvt_index_of(rep.centroids)
        fused_cells = grid is not None or index is not None
        _native.generate(rep.genotypes, rep_f, rep.centroids, ws, B, cfg["iso_sigma"], cfg["line_sigma"], cfg["minval"],
                         cfg["maxval"], cfg["task"], Dd, grid, winners and fused_cells, base, first,
                         buf["g"], buf["f"], buf["d"], buf["c"], gen_keys=gen_keys, index=index, fired_rows_only=p2p, out_xchg=p2p)
        if not fused_cells:   # cell assignment stays sharded: each rank assigns only its own offspring
            _native.cells(buf["d"], rep.centroids, None, ws, rep_f, buf["f"], offer=winners, idx_base=base, first_wins=first, out=buf["c"])
        self._mark("generate")
        if not winners:
            gb = self._gather_buffers(R, B, D, Dd, K, dev, "gather")
            G = parallel.all_gather_rows(buf["g"], self._group, gb["G"] if R > 1 else None)
            F = parallel.all_gather_rows(buf["f"], self._group, gb["F"] if R > 1 else None)
            Dn = parallel.all_gather_rows(buf["d"], self._group, gb["Dn"] if R > 1 else None)
            Cc = parallel.all_gather_rows(buf["c"], self._group, gb["C"] if R > 1 else None)
            self._mark("exchange")
            _native.offer_cells(Cc, F, ws, rep_f, 0, first)
            _native.commit(ws, G, F, Dn, rep.genotypes, rep_f, rep.descriptors, first_wins=first, qd_offset=cfg["qd_offset"],
                           metrics_out=metrics_out)
            self._mark("commit")
            return
        if R == 1:
            _native.commit(ws, buf["g"], buf["f"], buf["d"], rep.genotypes, rep_f, rep.descriptors, idx_base=base, first_wins=first,
                           qd_offset=cfg["qd_offset"], metrics_out=metrics_out)
            self._mark("commit")
            return
        if p2p:
            # the offers already pushed their keys into every peer's table (qdx_offer) and the rows of the fired offers sit in
            # this rank's offspring block; with fused cell assignment the last CTA of the generate kernel also raised the
            # arrival flags, otherwise a 1-thread kernel does.  The commit waits for every rank's flag and reads each
            # winner from its owner's block over NVLink.
            if not fused_cells:
                _native.xchg_push(ws, gen_keys)
                self._mark("exchange")
            _native.commit(ws, None, None, None, rep.genotypes, rep_f, rep.descriptors, first_wins=first, qd_offset=cfg["qd_offset"],
                           metrics_out=metrics_out, mode=3)
            self._mark("commit")
            return
        st = self._gather_buffers(R, B, D, Dd, K, dev, "stage")["stage"]
        sg, sd, sf = _stage_views(st, D, Dd)
        if self._exchange == "regen":
            parallel.all_reduce_max_i64_(ws.keytab(with_key_slots=True), self._group)
            self._mark("exchange_keys")
            _native.elect_winners(ws, rep.genotypes, cfg["task"], Dd, B, R, cfg["iso_sigma"], cfg["line_sigma"], cfg["minval"],
                                  cfg["maxval"], first, sg, sf, sd)
            self._mark("exchange")
            _native.commit(ws, sg, sf, sd, rep.genotypes, rep_f, rep.descriptors, first_wins=first, qd_offset=cfg["qd_offset"],
                           metrics_out=metrics_out, mode=2)
            self._mark("commit")
            return
        parallel.all_reduce_max_i64_(ws.keytab(), self._group)
        st.zero_()
        _native.commit(ws, buf["g"], buf["f"], buf["d"], sg, sf, sd, idx_base=base, first_wins=first, mode=1)
        parallel.all_reduce_disjoint_rows_(st, self._group)
        self._mark("exchange")
        _native.commit(ws, sg, sf, sd, rep.genotypes, rep_f, rep.descriptors, first_wins=first, qd_offset=cfg["qd_offset"],
                       metrics_out=metrics_out, mode=2)
        self._mark("commit")

    def update(self, repertoire: MapElitesRepertoire, emitter_state: Optional[EmitterState], key, *, donate: bool = False,
               metrics_out: Optional[torch.Tensor] = None):
        """reference :92-161.  `key` is this rank's key (examples/distributed_mapelites.ipynb cell 23:
        keys = split(key, num_devices))."""
        if self._scoring_function is None:
            raise ValueError("Scoring function is not set.")
        cfg = self._fused_config(repertoire)
        if cfg is not None:
            rep = repertoire if donate else repertoire._clone_state()
            m = _metrics_destination(metrics_out, rep.genotypes.device)
            self._fused_distributed_generation(rep, cfg, _native.KEYMODE_DIST_UPDATE, key, m)
            self._last_metrics = m
            return rep, emitter_state, self._metrics_dict(m)
        ks = qrandom.split(key)                                             # :124
        key, subkey = ks[0], ks[1]
        genotypes, extra_info = self._emitter.emit(repertoire, emitter_state, subkey)
        ks = qrandom.split(key)                                             # :128
        key, subkey = ks[0], ks[1]
        fitnesses, descriptors, extra_scores = self._scoring_function(genotypes, subkey)
        g = parallel.all_gather_rows(genotypes, self._group)               # :134-141
        f = parallel.all_gather_rows(fitnesses, self._group)
        d = parallel.all_gather_rows(descriptors, self._group)
        repertoire = repertoire.add(g, d, f)                                # :144-146
        emitter_state = self._emitter.state_update(emitter_state=emitter_state, repertoire=repertoire, genotypes=genotypes,
                                                   fitnesses=fitnesses, descriptors=descriptors,
                                                   extra_scores={**extra_scores, **extra_info})
        return repertoire, emitter_state, self._metrics_function(repertoire)

    def scan(self, carry, length: int, *, donate: bool = False, graph: bool = False):
        """`num_iterations` updates as in get_distributed_update_fn (:181-244): per iteration
        `key, subkey = split(key)` (:215) then update(subkey)."""
        repertoire, emitter_state, key = carry
        cfg = self._fused_config(repertoire) if self._scoring_function is not None else None
        rep = repertoire if (donate or cfg is None) else repertoire._clone_state()
        out = []
        for _ in range(length):
            ks = qrandom.split(key)
            key, subkey = ks[0], ks[1]
            if cfg is not None:
                m = torch.empty(4, dtype=torch.float32, device=rep.genotypes.device)
                self._fused_distributed_generation(rep, cfg, _native.KEYMODE_DIST_UPDATE, subkey, m)
                out.append(m)
            else:
                rep, emitter_state, md = self.update(rep, emitter_state, subkey)
                out.append(torch.stack([md["qd_score"], md["max_fitness"], md["coverage"], md["coverage"] * 0]))
        metrics = torch.stack(out) if out else torch.empty((0, 4))
        if cfg is not None:
            self.check_errors(rep)
        return (rep, emitter_state, key), self._metrics_dict(metrics)

    def check_errors(self, repertoire: MapElitesRepertoire) -> None:
        """Blocking: read this rank's device error flag and all-reduce it, so that a rank-local failure (a peer that did not
        arrive in time, QDX_ERR_PEER_TIMEOUT; a bad index) raises on EVERY rank together instead of leaving the others to
        run on with diverged replicas.  Called at the end of every scan; call it after a loop of update()s as well."""
        _, _, err = repertoire._workspace().read()
        worst = parallel.all_reduce_min_int(err, repertoire.fitnesses.device, self._group)
        if worst != 0:
            from qdax_b200._lib import QdxError
            raise QdxError("DistributedMAPElites (rank %d reports %d)" % (parallel.world(self._group)[0], err), worst)

    def get_distributed_init_fn(self, centroids, devices: Optional[List[Any]] = None) -> Callable:
        """reference :163-179.  One process per GPU: the returned function is called by every rank."""
        return lambda genotypes, key: self.init(genotypes, centroids, key)

    def get_distributed_update_fn(self, num_iterations: int, devices: Optional[List[Any]] = None) -> Callable:
        """reference :181-244."""
        def update_fn(repertoire, emitter_state, key):
            (rep, st, _), metrics = self.scan((repertoire, emitter_state, key), num_iterations)
            return rep, st, metrics
        return update_fn


def _stage_views(st: torch.Tensor, D: int, Dd: int):
    """The staging buffer is laid out as three contiguous blocks so each is a valid dense (K, *) array."""
    K = st.shape[0]
    flat = st.reshape(-1)
    sg = flat[: K * D].view(K, D)
    sd = flat[K * D: K * (D + Dd)].view(K, Dd)
    sf = flat[K * (D + Dd):].view(K)
    return sg, sd, sf
