#!/usr/bin/env python
"""bench.py -- offspring evaluated + inserted per second for the MAP-Elites generation step.

Workload (BASELINE.json configs[2], the configuration the headline metric and the 1e9/s target are quoted on; it
fits one GPU, so the same workload is used at every N): arm 100-DoF, grid 100x100 (K = 10^4), Iso+LineDD
(iso 0.05, line 0.1, clip [0,1]), B_total = 2^20 offspring per generation sharded over N GPUs (strong scaling),
replicated repertoire, per-rank keys split(key, N)[rank].  A step = ONE full generation through the public API
(MAPElites.update / DistributedMAPElites.update): select parents -> variation -> arm scoring -> cell assignment ->
per-cell best -> [exchange] -> commit into the repertoire -> QD metrics.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--exchange p2p|regen|winners|allgather] [--config c1..c5]
    python bench.py --impl reference ...      # CPU arm: the oracle port of the reference path on ALL host cores

One JSON line on stdout (rank 0).  Timing: CUDA events on the launching stream, barrier + synchronize on both
sides, max over ranks; W >= 3 warm-up steps; the 419 MB offspring buffer written and re-read every step exceeds the
126 MB L2 (inputs larger than L2).  The line also carries: `e2e` (same loop + the step's metrics read back to the host),
`insertions` (offspring actually inserted per second), `rooflines` (one entry per kernel, each against the bound that
applies to it; per-kernel counters come from profiles/ncu_traffic.json, written by tools/ncu_summary.py from ncu captures
of this binary), `other_configs` (BASELINE configs[0], [1], [3], [4] measured in the same run, N = 1) and `oracle_parity`
(the active code path checked against the C oracle inside this very process, every rank).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: task, D, grid shape / K, B_total, centroids kind
    "c1": dict(task="arm", D=100, grid=(100, 100), B=1024, cvt=False, note="BASELINE configs[0] (README example)"),
    "c2": dict(task="rastrigin", D=100, K=10000, B=65536, cvt=True, note="BASELINE configs[1] (CVT 10k centroids, bucket-index cells)"),
    "c3": dict(task="arm", D=100, grid=(100, 100), B=1 << 20, cvt=False, note="BASELINE configs[2] (arm 100-DoF, batch 2^20)"),
    "c4": dict(task="sphere", D=1000, K=50000, Dd=32, B=65536, cvt=True,
               note="BASELINE configs[3] (sphere 1000-D, desc = p[:32] (declared extension), 50k centroids, tensor-core cell assignment)"),
    "c5": dict(task="rastrigin", D=100, P=100000, B=1024, k=3, dns=True, cvt=False,
               note="BASELINE configs[4] (Dominated Novelty Search, population 100k, batch 1024, k=3)"),
}
METRIC = "offspring evaluated+inserted/sec"
UNIT = "offspring/s"
SMS = 148


def algorithmic_bytes(D, Dd):
    """g = 4D; stage (a) reads 2g (parents) and writes g; fused (b) adds the write of fitness (4) +
    descriptor (4 Dd); the cell id written for the exchange / debugging adds 4."""
    g = 4 * D
    return {"generate_per_offspring": 3 * g + 4 + 4 * Dd + 4,
            # insert (d): B*8 (fitness+cell, folded into generate on the fused path), K*(8+8+4), W*2*(g+4Dd+4)
            "commit_fixed_per_cell": 8 + 8 + 4, "commit_per_winner": 2 * (g + 4 * Dd + 4)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])), mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ CPU arm
def _oracle_state(co, cfg):
    """Initial repertoire of the CPU legs (same seeds as the GPU arm)."""
    import numpy as np
    from oracle import jax_prng as jr

    D, task = cfg["D"], cfg["task"]
    cent, K = _centroids_np(cfg)
    Dd = cfg.get("Dd", 2)
    init = co.uniform(jr.split(jr.key(42))[1], 100 * D).reshape(100, D)
    f0, d0 = co.score(task, init, Dd)
    g, f, d, _ = co.add(np.zeros((K, D)), np.full(K, -np.inf), np.zeros((K, Dd)), init, f0, d0, co.cells(d0, cent))
    return g, f, d, cent, K


def run_reference(args, cfg):
    """The reference's CPU path for this workload.  jax / jaxlib cannot be installed in this image (no wheel, no
    network), so this is the oracle PORT (oracle/qdx_oracle.c, OpenMP over all host cores) of exactly what the
    reference executes: UniformSelector x2 + isoline_variation (Threefry normal draws), arm scoring, brute-force
    get_cells_indices over K centroids, segment_max insertion, QD metrics.  A step is one generation of the FULL workload
    batch (same config as the GPU arm); the thread count is set explicitly to the cores this process may use, whatever
    OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)."""
    from oracle import c_oracle as co
    from oracle import jax_prng as jr

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    co.build()
    cores = host_cores()
    co.set_threads(cores)
    if cfg.get("dns"):
        return run_reference_dns(args, cfg, co, cores)
    task = cfg["task"]
    B = cfg["B"] if args.cpu_sample <= 0 else min(cfg["B"], args.cpu_sample)
    g, f, d, cent, K = _oracle_state(co, cfg)
    key = jr.key(7)
    g, f, d, key, _, _ = co.map_elites_scan(g, f, d, cent, key, args.warmup, B, task)
    t0 = time.perf_counter()
    g, f, d, key, m, secs = co.map_elites_scan(g, f, d, cent, key, args.steps, B, task)
    dt = time.perf_counter() - t0
    value = args.steps * B / dt
    sample = f"{args.steps} generations of {B} offspring (workload batch {cfg['B']}), K={K} brute-force cells, after {args.warmup} warm-up, {dt:.1f} s"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config_dict(args, cfg, B_step=B),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": co.get_threads(), "kind": "port", "sample": sample,
                         "stage_seconds": {"emit": secs[0], "score": secs[1], "cells": secs[2], "add+metrics": secs[3]},
                         "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
                         "label": "C/OpenMP restatement of QDax 0.5.1 (not jax[cpu]: jax is not installable here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "final_coverage": float(m[-1, 2]),
    }
    emit(line)


def _dns_inputs_np(co, cfg, P):
    from oracle import jax_prng as jr

    D, B = cfg["D"], cfg["B"]
    pg = co.uniform(jr.key(2), P * D).reshape(P, D)
    bg = co.uniform(jr.key(3), B * D).reshape(B, D)
    pf, pd = co.score(cfg["task"], pg)
    bf, bd = co.score(cfg["task"], bg)
    return pg, pf, pd, bg, bf, bd


def run_reference_dns(args, cfg, co, cores):
    """c5 on the CPU: the oracle's dense N^2 competition (the reference materialises (N, N) arrays and cannot run at all at
    this N on a 62 GB host); one add is ~10 s of CPU work, so the step count is capped."""
    P, B, k = cfg["P"], cfg["B"], cfg["k"]
    pg, pf, pd, bg, bf, bd = _dns_inputs_np(co, cfg, P)
    steps = max(1, min(args.steps, 2))
    t0 = time.perf_counter()
    for _ in range(steps):
        co.dns_add(pg, pf, pd, bg, bf, bd, k)
    dt = (time.perf_counter() - t0) / steps
    emit({"impl": "reference", "metric": METRIC, "value": B / dt, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": 0,
          "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": _config_dict(args, cfg, B_step=B),
          "cpu_baseline": {"value": B / dt, "unit": UNIT, "cores": co.get_threads(), "kind": "port",
                           "sample": f"{steps} DominatedNoveltyRepertoire.add of {B} offspring into a population of {P} (N = {P + B})"},
          "e2e": {"value": B / dt, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def _cvt_points(K, Dd=2):
    """seeded U[0,1)^Dd points standing in for k-means centroids (same work)"""
    import numpy as np

    return np.random.default_rng(0).random((K, Dd)).astype(np.float32)


def _centroids_np(cfg):
    """CPU legs only (imports the oracle)."""
    if cfg["cvt"]:
        return _cvt_points(cfg["K"], cfg.get("Dd", 2)), cfg["K"]
    from oracle import qdax_numpy as qn

    cent = qn.compute_euclidean_centroids(cfg["grid"], 0.0, 1.0)
    return cent, cent.shape[0]


def _config_dict(args, cfg, B_step, name=None, gpus=None):
    gpus = args.gpus if gpus is None else gpus
    if cfg.get("dns"):
        return {"workload": f"Dominated Novelty Search {cfg['task']} {cfg['D']}-D, population {cfg['P']}, batch {cfg['B']}, k={cfg['k']}: one "
                            f"DominatedNoveltyRepertoire.add per step; {cfg['note']}", "name": name or args.config, "global_batch": cfg["B"],
                "batch_per_step": B_step, "genotype_dim": cfg["D"], "population": cfg["P"], "descriptor_dim": 2, "parallelism": "dp1",
                "l2": "working set (41 MB population + candidates) fits L2; not flushed"}
    K = cfg["K"] if cfg["cvt"] else cfg["grid"][0] * cfg["grid"][1]
    return {"workload": f"MAP-Elites {cfg['task']} {cfg['D']}-D, K={K} {'CVT(seeded uniform)' if cfg['cvt'] else 'grid'} cells, "
                        f"batch {cfg['B']} per generation, iso 0.05 / line 0.1 / clip [0,1]; {cfg['note']}",
            "name": name or args.config, "global_batch": cfg["B"], "batch_per_step": B_step, "genotype_dim": cfg["D"], "cells": K,
            "descriptor_dim": cfg.get("Dd", 2), "parallelism": f"dp{gpus} (offspring sharded, repertoire replicated)",
            "exchange": args.exchange if gpus > 1 else "none", "l2": "working set > L2 (offspring buffer 400 B x batch)"
            if cfg["B"] * cfg["D"] * 4 > 126e6 else "L2 flushed between steps by a 256 MB write" if args.flush_l2 else "working set fits L2; not flushed"}


# ------------------------------------------------------------------------------------------------ GPU arm
def _peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p["hbm_gbs"], "of measured (MEASURED_PEAKS.json hbm_gbs)", p.get("bf16_tflops"), p.get("bf16_tflops_sustained")
    except Exception:
        return 6650.0, "of fallback", None, None


def _ncu_entry(kernel, config, n=1):
    """Counters of `kernel` at `config` from the committed ncu capture of this binary (profiles/ncu_traffic.json, written by
    tools/ncu_summary.py --update); None when no capture exists."""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return None
    return tr.get(f"{kernel}@{config}@n{n}")


def make_driver(cfg, B, world, exchange):
    import functools

    from qdax_b200.core.distributed_map_elites import DistributedMAPElites
    from qdax_b200.core.emitters.mutation_operators import isoline_variation
    from qdax_b200.core.emitters.standard_emitters import MixingEmitter
    from qdax_b200.core.map_elites import MAPElites
    from qdax_b200.tasks.arm import arm_scoring_function
    from qdax_b200.tasks.standard_functions import rastrigin_scoring_function, sphere_scoring_function
    from qdax_b200.utils.metrics import default_qd_metrics

    Dd = cfg.get("Dd", 2)
    scoring = {"arm": arm_scoring_function, "rastrigin": rastrigin_scoring_function,
               "sphere": functools.partial(sphere_scoring_function, desc_dim=Dd)}[cfg["task"]]
    emitter = MixingEmitter(lambda x, k: x, functools.partial(isoline_variation, iso_sigma=0.05, line_sigma=0.1, minval=0.0, maxval=1.0), 1.0, B)
    metrics_fn = functools.partial(default_qd_metrics, qd_offset=0.0)
    base = MAPElites(scoring, emitter, metrics_fn)
    me = DistributedMAPElites(scoring, emitter, metrics_fn, exchange=exchange) if world > 1 else base
    return me, base


def make_repertoire(cfg, base, dev):
    import torch

    from qdax_b200 import random as qr
    from qdax_b200.core.containers.mapelites_repertoire import compute_euclidean_centroids

    Dd = cfg.get("Dd", 2)
    if cfg["cvt"]:
        cent = torch.from_numpy(_cvt_points(cfg["K"], Dd)).to(dev)
    else:
        cent = compute_euclidean_centroids(cfg["grid"], 0.0, 1.0, device=dev)
    key = qr.key(42)
    key, subkey = qr.split(key)
    init = qr.uniform(subkey, (100, cfg["D"]), device=dev)          # identical on every rank -> identical replicas
    rep, state, _ = base.init(init, cent, qr.key(1))
    return rep, state, cent


def run_generations(me, rep, state, key, n, keep=None):
    """n generations through the public API, launch-only: host key chain (README.md:133 / notebook cell 25), update(donate)."""
    from qdax_b200 import random as qr

    for _ in range(n):
        ks = qr.split(key)
        key, sub = ks[0], ks[1]
        rep, state, _md = me.update(rep, state, sub, donate=True)
        if keep is not None:
            keep.append(me._last_metrics)
    return rep, state, key


def oracle_parity(world, rank, dev, exchange, cfg, gens=3, B_dev=4096):
    """The ACTIVE code path (same driver class, same exchange, same C entry point) against the C oracle, inside the bench
    process: `gens` generations of B_dev offspring per rank from a fresh repertoire; every rank compares ITS replica with
    oracle.distributed_update (global offspring index = rank * B_dev + i) bit for bit, and the verdicts are AND-ed."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from oracle import c_oracle as co
    from oracle import jax_prng as jr
    from qdax_b200 import random as qr

    co.build()
    co.set_threads(max(1, min(8, host_cores() // max(world, 1))))
    N = lambda t: t.detach().cpu().numpy()
    me, base = make_driver(cfg, B_dev, world, exchange)
    rep, state, cent_t = make_repertoire(cfg, base, dev)
    g, f, d, cent, K = _oracle_state(co, cfg)
    ok = bool(np.array_equal(N(rep.genotypes), g) and np.array_equal(N(rep.fitnesses).ravel(), f))
    keys = qr.split(qr.key(5), world)
    key = keys[rank]
    okeys = [np.array(k) for k in keys]
    for _ in range(gens):
        ks = qr.split(key)
        key, sub = ks[0], ks[1]
        rep, state, m = me.update(rep, state, sub, donate=True)
        if world > 1:
            subs = []
            for r in range(world):
                s2 = jr.split(okeys[r])
                okeys[r] = s2[0]
                subs.append(s2[1])
            g, f, d, *_ = co.distributed_update(g, f, d, cent, np.stack(subs), B_dev, cfg["task"])
        else:
            s2 = jr.split(okeys[0])
            okeys[0] = s2[0]
            emit_key = jr.split(jr.split(s2[1])[1])[1]            # update(sub): _, s1 = split(sub); ask(s1): _, e = split(s1)  (map_elites.py:177, :241)
            x, _, _ = co.emit_isoline(g, f, emit_key, B_dev, 0.05, 0.1, 0.0, 1.0)
            fx, dx = co.score(cfg["task"], x, cfg.get("Dd", 2))
            g, f, d, _ = co.add(g, f, d, x, fx, dx, co.cells(dx, cent))
        ok = ok and bool(np.array_equal(N(rep.fitnesses).ravel(), f) and np.array_equal(N(rep.genotypes), g) and np.array_equal(N(rep.descriptors), d))
        ref = co.metrics(f, 0.0)
        ok = ok and bool(np.allclose([float(m["qd_score"]), float(m["max_fitness"]), float(m["coverage"])], ref, rtol=1e-5))
    if world > 1:
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(int(t[0]))
    return {"ok": ok, "generations": gens, "offspring_per_rank": B_dev, "ranks": world,
            "checked": "every rank's repertoire (genotypes, fitnesses, descriptors: bit-exact; metrics 1e-5) vs oracle/qdx_oracle.c "
                       + ("distributed_update" if world > 1 else "emit + score + cells + add")}


def timed_region(me, rep, state, key, steps, barrier, world, dev):
    """K generations, device timed (events around the whole region only), max over ranks.  Returns (ms_total, host enqueue ms
    per step, per-step metrics tensors, rep, state, key)."""
    import torch
    import torch.distributed as dist

    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    keep = []
    ev0.record()
    t_host0 = time.perf_counter()
    rep, state, key = run_generations(me, rep, state, key, steps, keep)
    host_ms = (time.perf_counter() - t_host0) * 1e3 / steps
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), host_ms, keep, rep, state, key


def instrumented_pass(me, rep, state, key, steps, barrier):
    """Same loop with a CUDA event after every kernel (kernel-by-kernel enqueue path): per-kernel average ms.  Separate from
    the timed region: an event record between two kernels stops the next launch from being staged behind the running one."""
    import numpy as np
    import torch

    me._timeline = []
    barrier()
    iv0, iv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iv0.record()
    rep, state, key = run_generations(me, rep, state, key, steps)
    iv1.record()
    barrier()
    tl, me._timeline = me._timeline, None
    per, prev = {}, None
    for label, e in tl:
        if label != "begin" and prev is not None:
            per.setdefault(label, []).append(prev.elapsed_time(e))
        prev = e
    return {k: float(np.mean(v)) for k, v in per.items()}, iv0.elapsed_time(iv1) / steps, rep, state, key


def kernel_rooflines(cfg, name, kern_ms, B, K, world, clocks_mhz, W):
    """One entry per kernel of the generation, each against the bound that applies to IT."""
    hbm_peak, peak_src, bf16_burst, bf16_sust = _peaks()
    D, Dd = cfg["D"], cfg.get("Dd", 2)
    ab = algorithmic_bytes(D, Dd)
    out = {}
    if "generate" in kern_ms:
        t = kern_ms["generate"] * 1e-3
        gb = ab["generate_per_offspring"] * B
        ent = _ncu_entry("qdx_generate_kernel", name, world) or _ncu_entry("qdx_generate_kernel", name, 1)
        r = {"kernel": "qdx_generate_kernel", "bound": "issue", "avg_launch_ms": kern_ms["generate"],
             "hbm": {"achieved": gb / t / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": gb / t / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": gb}}
        if ent and ent.get("warp_instr_per_row"):
            clk = (clocks_mhz or 1965.0) * 1e6
            peak_issue = SMS * 4 * clk                                        # one warp-instruction per SM sub-partition per clock
            ach = ent["warp_instr_per_row"] * B / t
            r.update({"achieved": ach / 1e9, "peak": peak_issue / 1e9, "unit": "G warp-instr/s", "frac": ach / peak_issue,
                      "warp_instr_per_row": ent["warp_instr_per_row"], "pipes_pct_ncu": ent.get("pipes_pct"), "source": ent.get("source"),
                      "note": "instruction-issue bound: one Threefry-2x32-20 block + erfinv per gene, sincos per joint; the ALU pipe (LOP3 / SHF / "
                              "IADD3, half rate) is the busiest unit; DRAM traffic = the offspring rows written (ncu)"})
        out["generate"] = r
    if "cells" in kern_ms:
        t = kern_ms["cells"] * 1e-3
        if Dd >= 8:
            flop = 2.0 * B * K * Dd                                               # GEMM formulation: -2 x.c (+ norms)
            tf32_peak = (bf16_burst or 1609.4) / 2.0                              # dense TF32 = half the measured bf16 figure
            ent = _ncu_entry("qdx_cells_tc_kernel", name, 1)
            out["cells"] = {"kernel": "qdx_cells_tc_kernel (tcgen05 kind::tf32 + exact FP32 re-rank)", "bound": "tensor", "avg_launch_ms": kern_ms["cells"],
                            "achieved": flop / t / 1e12, "peak": tf32_peak, "unit": "TFLOP/s", "frac": flop / t / 1e12 / tf32_peak,
                            "peak_source": "half of MEASURED_PEAKS.json bf16_tflops (TF32 runs at half the bf16 rate)", "flop_per_launch": flop,
                            "pipes_pct_ncu": ent.get("pipes_pct") if ent else None, "source": ent.get("source") if ent else None}
        else:
            flop = 3.0 * B * K * Dd
            out["cells"] = {"kernel": "qdx_cells_bf_kernel", "bound": "fp32", "avg_launch_ms": kern_ms["cells"], "achieved": flop / t / 1e12,
                            "peak": SMS * 128 * 2 * (clocks_mhz or 1965.0) * 1e6 / 1e12, "unit": "TFLOP/s (FP32, FMA = 2)", "flop_per_launch": flop}
    if "commit" in kern_ms:
        t = kern_ms["commit"] * 1e-3
        cb = K * ab["commit_fixed_per_cell"] + W * ab["commit_per_winner"]
        out["commit"] = {"kernel": "qdx_commit_lean_kernel" if D <= 256 else "qdx_commit_stream_kernel", "bound": "hbm", "achieved": cb / t / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": cb / t / 1e9 / hbm_peak, "winners_last_step": W, "algorithmic_bytes_per_launch": cb, "avg_launch_ms": kern_ms["commit"],
                         "note": "launch / latency bound whenever W * row bytes is small (SURVEY.md 8d caveat); see insert_roofline.large_rows"}
    return out


def measure_config(name, args, dev, steps):
    """One of the other BASELINE configurations on one GPU, same method as the headline: public API, CUDA events, instrumented
    pass for the per-kernel split."""
    import torch

    from qdax_b200 import random as qr

    cfg = CONFIGS[name]
    if cfg.get("dns"):
        return measure_dns(cfg, args, dev, steps)
    me, base = make_driver(cfg, cfg["B"], 1, "none")
    rep, state, cent = make_repertoire(cfg, base, dev)
    K = cent.shape[0]
    rep = rep._clone_state()
    key = qr.key(7)
    barrier = torch.cuda.synchronize
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if cfg["B"] * cfg["D"] * 4 <= 126e6 else None
    rep, state, key = run_generations(me, rep, state, key, max(args.warmup, 3))
    if flush is not None:
        flush.fill_(1)                       # working set fits L2: start the timed region from a flushed L2
    ms_total, host_ms, keep, rep, state, key = timed_region(me, rep, state, key, steps, barrier, 1, dev)
    kern_ms, ms_instr, rep, state, key = instrumented_pass(me, rep, state, key, steps, barrier)
    m = torch.stack(keep).cpu()
    inserted = float(m[:, 3].sum())
    value = steps * cfg["B"] / (ms_total * 1e-3)
    scan = None
    if name == "c1":
        # the README idiom: jax.lax.scan(map_elites.scan_update, ...) = MAPElites.scan; wall clock from the call to the final carry
        # key and the stacked metrics on the host (README.md:121-140), 50 iterations, replayed as one CUDA graph
        import time
        n_it = 50
        (rep, state, key), _ = me.scan((rep, state, key), n_it, donate=True)                    # one-time costs outside the timing
        torch.cuda.synchronize()
        loops = []
        for _ in range(3):
            t0 = time.perf_counter()
            (rep, state, key), ms_ = me.scan((rep, state, key), n_it, donate=True)
            host = {k: v.cpu() for k, v in ms_.items()}
            loops.append(time.perf_counter() - t0)
        t_loop = sorted(loops)[1]
        (rep, state, key), _ = me.scan((rep, state, key), n_it, donate=True, graph=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        (rep, state, key), ms_ = me.scan((rep, state, key), n_it, donate=True, graph=True)     # captured again on every call
        host = {k: v.cpu() for k, v in ms_.items()}
        t_graph = time.perf_counter() - t0
        scan = {"iterations": n_it, "api": "MAPElites.scan (= jax.lax.scan(scan_update)) + metrics to the host, wall clock",
                "offspring_per_s": n_it * cfg["B"] / t_loop, "ms_per_iteration": t_loop / n_it * 1e3,
                "graph_offspring_per_s_incl_capture": n_it * cfg["B"] / t_graph, "graph_ms_per_iteration_incl_capture": t_graph / n_it * 1e3}
    return {"config": _config_dict(args, cfg, cfg["B"], name=name, gpus=1), "value": value, "unit": UNIT, "ms_per_step": ms_total / steps, "steps": steps,
            "readme_scan": scan,
            "kernel_ms": kern_ms, "host_enqueue_ms_per_step": host_ms, "insertions_per_s": inserted / (ms_total * 1e-3),
            "rooflines": kernel_rooflines(cfg, name, kern_ms, cfg["B"], K, 1, None, float(m[-1, 3])),
            "final": {"coverage": float(m[-1, 2]), "qd_score": float(m[-1, 0])},
            "l2": "flushed once before the timed region (working set fits L2)" if flush is not None else "working set > L2"}


def measure_dns(cfg, args, dev, steps):
    """BASELINE configs[4]: one DominatedNoveltyRepertoire.add of 1024 offspring into a full population of 100 000 (k = 3)."""
    import torch

    from qdax_b200 import random as qr
    from qdax_b200.core.containers.dns_repertoire import DominatedNoveltyRepertoire
    from qdax_b200.tasks.standard_functions import rastrigin_scoring_function

    P, B, D, k = cfg["P"], cfg["B"], cfg["D"], cfg["k"]
    pg = qr.uniform(qr.key(2), (P, D), device=dev)
    pf, pd, _ = rastrigin_scoring_function(pg)
    bg = qr.uniform(qr.key(3), (B, D), device=dev)
    bf, bd, _ = rastrigin_scoring_function(bg)
    rep = DominatedNoveltyRepertoire(genotypes=pg, fitnesses=pf.reshape(P, 1), descriptors=pd, k=k)
    for _ in range(3):
        new = rep.add(bg, bd, bf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        new = rep.add(bg, bd, bf)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    Nn = P + B
    pairs = Nn * (Nn - 1) / 2.0                      # triangular competition (candidates sorted by fitness)
    fp32_peak = SMS * 128 * 1965.0e6 / 1e12         # T op/s of single FP32 instructions (no FMA contraction in the spec)
    flop = pairs * 8.0                               # Dd = 2: 2 sub, 2 mul, 1 add + compare / select of the k-list
    return {"config": _config_dict(args, cfg, B, name="c5", gpus=1), "value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "candidates": Nn, "pairs_per_s": Nn * Nn / (ms * 1e-3), "survivors_changed": int((new._last_survivors >= P).sum()),
            "rooflines": {"dns_knn": {"kernel": "qdx_dns_knn_sorted_kernel", "bound": "issue (FP32 pipe)", "achieved": flop / (ms * 1e-3) / 1e12,
                                      "peak": fp32_peak, "unit": "T FP32 instr/s", "frac": flop / (ms * 1e-3) / 1e12 / fp32_peak,
                                      "note": "whole add (rank + k-NN + survivors + gather) timed; the k-NN kernel is ~90 % of it"}}}


def run_gpu(args, cfg):
    import numpy as np
    import torch
    import torch.distributed as dist

    from qdax_b200 import _lib
    from qdax_b200 import random as qr

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N bench.py --gpus N ...")
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if cfg.get("dns"):
        if rank == 0:
            r = measure_dns(cfg, args, dev, args.steps)
            emit({"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": 3, "ms_per_step": r["ms_per_step"],
                  "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", **{k: v for k, v in r.items() if k not in ("value", "unit", "ms_per_step", "steps")}})
        return

    if args.weak and world > 1:          # weak scaling (not the BASELINE configuration): every rank keeps the 1-GPU batch
        cfg = dict(cfg, B=cfg["B"] * world, note=cfg["note"] + " -- WEAK scaling variant: batch x N GPUs")
    D, task, B_total = cfg["D"], cfg["task"], cfg["B"]
    assert B_total % world == 0
    B = B_total // world
    Dd = cfg.get("Dd", 2)
    me, base = make_driver(cfg, B, world, args.exchange)
    rep, state, cent = make_repertoire(cfg, base, dev)
    K = cent.shape[0]
    assert me._fused_config(rep) is not None, "bench configuration must take the fused native path"
    rank_key = qr.split(qr.key(7), world)[rank]               # examples/distributed_mapelites.ipynb cell 23

    # ---- cold start: the first generations after init insert thousands of offspring per step (insertions/s where it is not ~0)
    warm = rep._clone_state()
    run_generations(me, warm, state, rank_key, 2)             # module load / buffer allocation on a throw-away copy
    rep = rep._clone_state()
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cold_keep = []
    c0.record()
    rep, state, key = run_generations(me, rep, state, rank_key, 3, cold_keep)
    c1.record()
    barrier()
    cold_ms = c0.elapsed_time(c1)
    cold_inserted = float(torch.stack(cold_keep)[:, 3].sum())

    for _ in range(max(args.warmup, 3)):
        rep, state, key = run_generations(me, rep, state, key, 1)
    barrier()
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record()
    rep, state, key = run_generations(me, rep, state, key, 5)   # rough ms per generation (after the one-time costs), untimed
    w1.record()
    barrier()
    warm_ms = torch.tensor([w0.elapsed_time(w1) / 5], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(warm_ms, op=dist.ReduceOp.MAX)        # the same figure, hence the same step count, on every rank

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # the recipe's clocks line (-lms 200): started before, killed after the timed region
    # nvidia-smi needs ~0.3 s to initialise and its first query stalls the GPU for ~0.5 ms -- 15 % of a 4 ms timed region (20
    # generations at N = 8).  Keep the GPU under load with untimed generations (the same count on every rank) while it does.
    rep, state, key = run_generations(me, rep, state, key, min(int(600.0 / max(float(warm_ms[0]), 1e-3)) + 1, 20000))
    launches0 = _lib.launch_count
    ms_total, host_enqueue_ms, keep, rep, state, key = timed_region(me, rep, state, key, args.steps, barrier, world, dev)
    launches = _lib.launch_count - launches0
    kern_ms, ms_instr, rep, state, key = instrumented_pass(me, rep, state, key, args.steps, barrier)
    clocks = sampler.stop() if rank == 0 else None
    m_all = torch.stack(keep).cpu()
    m = m_all[-1]
    added_last, coverage, qd = float(m[3]), float(m[2]), float(m[0])
    inserted = float(m_all[:, 3].sum())
    value = args.steps * B_total / (ms_total * 1e-3)

    # ---- e2e: the same public call + the step's metrics read back to the host, inside the timed region -------------
    # "pipelined" has every step's metrics written to PINNED host memory by the commit kernel itself (update(..., metrics_out=
    # pinned slot)) + an event, and consumes them one step behind (what a loop that logs metrics does); "blocking" calls .cpu() on the metrics after every step.
    e2e_steps = args.steps
    pinned = torch.empty((2, 4), dtype=torch.float32).pin_memory()
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    e2e = {}
    for mode in ("blocking", "pipelined"):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s_i in range(e2e_steps):
            ks = qr.split(key)                                    # host-side key chain, README.md:133
            key, sub = ks[0], ks[1]
            if mode == "blocking":
                rep, state, md = me.update(rep, state, sub, donate=True)
                host_metrics = me._last_metrics.cpu()             # D2H of the step's metrics (16 B) + sync every step
            else:
                # the commit kernel writes the step's metrics straight into the pinned host slot (16 B over PCIe, no copy node
                # between this generation's commit and the next generation's generate); the event orders the host's read
                rep, state, md = me.update(rep, state, sub, donate=True, metrics_out=pinned[s_i & 1])
                evs[s_i & 1].record()
                if s_i > 0:
                    evs[(s_i - 1) & 1].synchronize()
                    host_metrics = pinned[(s_i - 1) & 1].clone()  # step s-1's metrics, on the host, while step s runs
        if mode == "pipelined":
            evs[(e2e_steps - 1) & 1].synchronize()
            host_metrics = pinned[(e2e_steps - 1) & 1].clone()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e[mode] = float(t[0])
    e2e_ms = e2e["pipelined"]
    e2e_value = e2e_steps * B_total / (e2e_ms * 1e-3)
    e2e_blocking_value = e2e_steps * B_total / (e2e["blocking"] * 1e-3)
    if world > 1:
        me.check_errors(rep)                                      # device error flags, all-reduced: every rank raises together

    consistent = True
    if world > 1:
        from qdax_b200 import parallel
        consistent = parallel.all_equal(rep.genotypes) and parallel.all_equal(rep.fitnesses)
    parity = None
    if not args.no_oracle_parity:
        parity = oracle_parity(world, rank, dev, args.exchange, cfg)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines: every kernel against ITS bound; the dominant kernel's entry is also the contract's `roofline` object ---
    hbm_peak, peak_src, _, _ = _peaks()
    ab = algorithmic_bytes(D, Dd)
    rl = kernel_rooflines(cfg, args.config, kern_ms, B, K, world, clocks["sm_mhz"] if clocks else None, added_last)
    if world > 1:
        # north_star: "as a fraction of the HBM / NVLink roofline".  What the reference-faithful exchange (all_gather of every
        # offspring tuple, distributed_map_elites.py:133-141) would move per rank and generation, against what the active one moves.
        nvlink_peak = 900.0                                   # GB/s per direction and GPU (NVLink 5 through NVSwitch, nominal)
        per_off = 4 * D + 4 + 4 * Dd + 4
        ag_bytes = (world - 1) * B * per_off                  # received by every rank
        used = getattr(me, "_exchange", args.exchange)
        ex = {"exchange": used, "bound": "nvlink" if used == "allgather" else "latency (system-scope atomics + winner rows over NVLink)",
              "peak": nvlink_peak, "unit": "GB/s per direction", "peak_source": "nominal NVLink 5 per GPU",
              "allgather_bytes_per_rank_per_generation": ag_bytes, "allgather_ms_at_peak": ag_bytes / (nvlink_peak * 1e9) * 1e3}
        if used == "allgather" and "exchange" in kern_ms:
            ex["achieved"] = ag_bytes / (kern_ms["exchange"] * 1e-3) / 1e9
            ex["frac"] = ex["achieved"] / nvlink_peak
        else:
            # p2p / regen / winners: keys of improving offers (8 B x (R - 1) each), 8 key words + a flag per peer, and the rows of the
            # elected winners owned by other ranks -- a few KB per generation in steady state
            moved = added_last * (world - 1) / world * per_off + (world - 1) * (8 * 8 + 8) + 64 * (world - 1) * 8
            ex["bytes_per_rank_per_generation_estimate"] = moved
            ex["achieved"] = moved / (ms_total / args.steps * 1e-3) / 1e9
            ex["frac"] = ex["achieved"] / nvlink_peak
            ex["note"] = ("the exchange is not bandwidth-bound: it moves ~%.1f KB per rank and generation where the all-gather of the reference would "
                          "move %.0f MB (%.2f ms at the NVLink peak, against a %.3f ms generation)" % (moved / 1e3, ag_bytes / 1e6, ex["allgather_ms_at_peak"], ms_total / args.steps))
        rl["exchange"] = ex
    dom = max((k for k in kern_ms if k in ("generate", "cells")), key=lambda k: kern_ms[k])
    dom_bytes = ab["generate_per_offspring"] * B if dom == "generate" else (4 * Dd + 4) * B + K * Dd * 4
    dom_gbs = dom_bytes / (kern_ms[dom] * 1e-3) / 1e9
    ent = _ncu_entry({"generate": "qdx_generate_kernel", "cells": "qdx_cells_tc_kernel"}[dom], args.config, world)
    roofline = {"kernel": rl[dom]["kernel"], "bound": "hbm", "achieved": dom_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": dom_gbs / hbm_peak,
                "traffic": ent.get("traffic_bytes") if ent else None, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                "avg_launch_ms": kern_ms[dom], "binding_limit": rl[dom]["bound"],
                "binding_frac": rl[dom].get("frac"), "binding_unit": rl[dom].get("unit"),
                "note": "the HBM figure is what the contract asks for; this kernel is bound by '%s' (rooflines.%s): parents are read from L2, "
                        "only the offspring rows reach DRAM" % (rl[dom]["bound"], dom)}
    insert = dict(rl.get("commit", {}))
    if world == 1 and not args.no_insert_probe:
        # the insert kernel's HBM roofline is only meaningful where W * row bytes is large: the headline figure is
        # the SUSTAINED one of the 4 KB-row probe (train of 8 back-to-back launches); the timed workload's own commit sits beside it
        probe = insert_probe(dev, hbm_peak)
        insert = {"kernel": "qdx_commit_stream_kernel", "bound": "hbm", "achieved": probe["train_of_8_achieved"], "peak": hbm_peak, "unit": "GB/s",
                  "frac": probe["train_of_8_frac"], "what": "train of 8 back-to-back launches at 4 KB rows (sustained), of the measured HBM copy peak",
                  "single_launch_frac": probe["frac"], "traffic": probe["traffic"], "large_rows": probe,
                  "headline": {"frac": probe["train_of_8_frac"], "what": "train of 8 back-to-back launches at 4 KB rows (sustained), of measured HBM copy peak"},
                  "timed_workload_commit": rl.get("commit", {})}

    others = None
    if world == 1 and not args.no_other_configs:
        others = {}
        for name in ("c1", "c2", "c4", "c5"):
            if name == args.config:
                continue
            try:
                others[name] = measure_config(name, args, dev, args.steps)
            except Exception as e:                       # the headline line must survive a failure here
                others[name] = {"error": repr(e)}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(cfg, args)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak" if (args.weak and world > 1) else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": _config_dict(args, cfg, B_step=B_total),
        "timed_api": ("MAPElites.update" if world == 1 else "DistributedMAPElites.update") + "(repertoire, emitter_state, key, donate=True), host key chain",
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 16, "ms_per_step": e2e_ms / e2e_steps,
                "api": ("MAPElites.update(repertoire, emitter_state, key, donate=True)" if world == 1 else "DistributedMAPElites.update(...) per rank")
                       + " + every step's metrics written by the commit kernel to pinned host memory (metrics_out=, 16 B over PCIe, + event) and read one step behind",
                "blocking_readback_value": e2e_blocking_value,
                "note": "inputs of a step are the 2-word RNG key (host) and the HBM-resident repertoire (carried state)"},
        "insertions": {"per_s": inserted / (ms_total * 1e-3), "per_step": inserted / args.steps,
                       "cold_start": {"steps": 3, "inserted": cold_inserted, "per_s": cold_inserted / (cold_ms * 1e-3), "ms_per_step": cold_ms / 3,
                                      "what": "the first 3 generations after init (repertoire 1 % full)"},
                       "what": "offspring that entered the repertoire (4th metric of the commit kernel) in the timed region: steady state, repertoire converged"},
        "gpu_launches": launches, "kernel_ms": kern_ms,
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "kernel_ms_source": "instrumented pass of the same K steps right after the timed region (kernel-by-kernel enqueue, one CUDA event after every "
                            "kernel); ms per step there: %.4f" % ms_instr, "roofline": roofline, "rooflines": rl, "insert_roofline": insert,
        "clocks": clocks, "replicas_bit_identical": consistent,
        "oracle_parity": parity["ok"] if parity else None, "oracle_parity_detail": parity,
        "exchange_used": (getattr(me, "_exchange", "none") if world > 1 else "none"), "exchange_fallback": getattr(me, "exchange_fallback", None),
        "final": {"coverage": coverage, "qd_score": qd, "inserted_last_step": added_last},
    }
    if others is not None:
        line["other_configs"] = others
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def insert_probe(dev, hbm_peak):
    """The insert (commit) kernel where its HBM roofline is meaningful (SURVEY.md 8d caveat): BASELINE configs[3] shape --
    K = 50 000 cells, D = 1000 (4 KB rows), Dd = 32 -- 65 536 offspring into an EMPTY repertoire, so ~30 000 winner rows
    (~250 MB algorithmic) move in one launch.  Outside the timed region of the headline; CUDA events around ONE launch after
    an L2-evicting read pass (median of 5), and around a train of 8 back-to-back launches into 8 empty repertoires (the
    sustained figure, which is the headline: a single launch is flattered by rows still dirty in L2 when it ends)."""
    import numpy as np
    import torch

    from qdax_b200 import _native

    K, D, Dd, B, R = 50000, 1000, 32, 65536, 8
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    cent = torch.rand(K, Dd, device=dev, generator=gen)
    gs = [torch.rand(B, D, device=dev, generator=gen) for _ in range(2)]
    d = torch.rand(B, Dd, device=dev, generator=gen)
    f = torch.randn(B, device=dev, generator=gen)
    cells = _native.cells(d, cent, None)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    m = torch.empty(4, device=dev)
    new_rep = lambda: (torch.zeros(K, D, device=dev), torch.empty(K, device=dev), torch.zeros(K, Dd, device=dev), _native.Workspace(K, dev))
    reps = [new_rep()]

    def arm(n):
        for (rg, rf, rd, w) in reps[:n]:
            rf.fill_(float("-inf"))
            _native.offer_cells(cells, f, w, rf)
        flush.fill_(1)
        _ = flush.view(torch.int32).sum()          # evict L2 with a READ pass: no dirty lines left to write back

    single, train = [], []
    for rep in range(6):
        arm(1)
        rg, rf, rd, w = reps[0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _native.commit(w, gs[0], f, d, rg, rf, rd, metrics_out=m)
        e1.record()
        torch.cuda.synchronize()
        single.append(e0.elapsed_time(e1))
    W = float(m[3])
    reps += [new_rep() for _ in range(R - 1)]
    for rep in range(4):
        arm(R)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r, (rg, rf, rd, w) in enumerate(reps):
            _native.commit(w, gs[r & 1], f, d, rg, rf, rd, metrics_out=m)
        e1.record()
        torch.cuda.synchronize()
        train.append(e0.elapsed_time(e1) / R)
    nbytes = K * 20 + W * 2 * (4 * D + 4 * Dd + 4)
    ms1, mst = float(np.median(single[1:])), float(np.median(train[1:]))
    ent = _ncu_entry("qdx_commit_stream_kernel", "c4_cold_start", 1)
    return {"workload": "K=50000 cells, D=1000, Dd=32, 65536 offspring into an empty repertoire (BASELINE configs[3] cold start)",
            "winners": W, "algorithmic_bytes_per_launch": nbytes, "launch_ms": ms1, "achieved": nbytes / (ms1 * 1e-3) / 1e9,
            "peak": hbm_peak, "unit": "GB/s", "frac": nbytes / (ms1 * 1e-3) / 1e9 / hbm_peak,
            "train_of_8_launch_ms": mst, "train_of_8_achieved": nbytes / (mst * 1e-3) / 1e9, "train_of_8_frac": nbytes / (mst * 1e-3) / 1e9 / hbm_peak,
            "traffic": ent.get("traffic_bytes") if ent else None, "traffic_source": ent.get("source") if ent else None,
            "timing": "CUDA events; single launch after an L2-evicting read pass (median of 5) / 8 back-to-back launches"}


def cpu_baseline(cfg, args):
    """Oracle port timed on the host cores on a bounded sample (about 10-30 s of CPU work): whole generations of the FULL
    workload batch, all cores (thread count set explicitly)."""
    from oracle import c_oracle as co
    from oracle import jax_prng as jr

    co.build()
    co.set_threads(host_cores())
    task = cfg["task"]
    B = cfg["B"]
    g, f, d, cent, K = _oracle_state(co, cfg)
    t0 = time.perf_counter()
    g, f, d, key, _, _ = co.map_elites_scan(g, f, d, cent, jr.key(7), 1, B, task)
    t1 = time.perf_counter() - t0
    n = int(min(40, max(2, 12.0 / max(t1, 1e-3))))
    t0 = time.perf_counter()
    g, f, d, key, m, secs = co.map_elites_scan(g, f, d, cent, key, n, B, task)
    dt = time.perf_counter() - t0
    return {"value": n * B / dt, "unit": UNIT, "cores": co.get_threads(), "kind": "port",
            "sample": f"{n} generations of {B} offspring (the full workload batch), K={K} brute-force cells, {dt:.1f} s",
            "stage_seconds": {"emit": secs[0], "score": secs[1], "cells": secs[2], "add+metrics": secs[3]},
            "label": "C/OpenMP restatement of QDax 0.5.1 (not jax[cpu])"}


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) was sent to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # library chatter on fd 1 (e.g. "NCCL version ...") -> stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "regen", "winners", "allgather"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="offspring per generation in the CPU arm (0 = the full workload batch)")
    ap.add_argument("--flush-l2", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-insert-probe", action="store_true", help="skip the large-row insert-kernel measurement (N=1 only)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip c1 / c2 / c4 / c5 (N=1 only)")
    ap.add_argument("--no-oracle-parity", action="store_true")
    ap.add_argument("--weak", action="store_true", help="N > 1: keep the 1-GPU batch per rank (weak scaling; not the BASELINE configuration)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_gpu(args, cfg)


if __name__ == "__main__":
    main()
