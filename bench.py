#!/usr/bin/env python
"""bench.py -- offspring evaluated + inserted per second for the MAP-Elites generation step.

Workload (BASELINE.json configs[2], the configuration the headline metric and the 1e9/s target are quoted on; it
fits one GPU, so the same workload is used at every N): arm 100-DoF, grid 100x100 (K = 10^4), Iso+LineDD
(iso 0.05, line 0.1, clip [0,1]), B_total = 2^20 offspring per generation sharded over N GPUs (strong scaling),
replicated repertoire, per-rank keys split(key, N)[rank].  A step = ONE full generation: select parents ->
variation -> arm scoring -> cell assignment -> per-cell best -> [exchange] -> commit into the repertoire -> QD metrics.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--exchange p2p|regen|winners|allgather] [--config c1|c2|c3]
    python bench.py --impl reference ...      # CPU arm: the oracle port of the reference path on the host cores

One JSON line on stdout (rank 0).  Timing: CUDA events on the launching stream, barrier + synchronize on both
sides, max over ranks; W >= 3 warm-up steps; the 419 MB offspring buffer written and re-read every step exceeds the
126 MB L2 (inputs larger than L2).
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
    "c2": dict(task="rastrigin", D=100, K=10000, B=65536, cvt=True, note="BASELINE configs[1] (CVT 10k, brute-force cells)"),
    "c3": dict(task="arm", D=100, grid=(100, 100), B=1 << 20, cvt=False, note="BASELINE configs[2] (arm 100-DoF, batch 2^20)"),
    "c4": dict(task="sphere", D=1000, K=50000, Dd=32, B=65536, cvt=True,
               note="BASELINE configs[3] (sphere 1000-D, desc = p[:32] (declared extension), 50k centroids, tensor-core cell assignment)"),
}
METRIC = "offspring evaluated+inserted/sec"
UNIT = "offspring/s"


def algorithmic_bytes(D, Dd):
    """g = 4D; stage (a) reads 2g (parents) and writes g; fused (b) adds the write of fitness (4) +
    descriptor (4 Dd); the cell id written for the exchange / debugging adds 4."""
    g = 4 * D
    return {"generate_per_offspring": 3 * g + 4 + 4 * Dd + 4,
            # insert (d): B*8 (fitness+cell, folded into generate on the fused path), K*(8+8+4), W*2*(g+4Dd+4)
            "commit_fixed_per_cell": 8 + 8 + 4, "commit_per_winner": 2 * (g + 4 * Dd + 4)}


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
def run_reference(args, cfg):
    """The reference's CPU path for this workload.  jax / jaxlib cannot be installed in this image (no wheel, no
    network), so this is the oracle PORT (oracle/qdx_oracle.c, OpenMP over all host cores) of exactly what the
    reference executes: UniformSelector x2 + isoline_variation (Threefry normal draws), arm scoring, brute-force
    get_cells_indices over K centroids, segment_max insertion, QD metrics.  Each step is a bounded sample of the
    workload: one generation of B_sample offspring."""
    import numpy as np

    from oracle import c_oracle as co
    from oracle import jax_prng as jr
    from oracle import qdax_numpy as qn

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    co.build()
    D, task = cfg["D"], cfg["task"]
    B = min(cfg["B"], args.cpu_sample)
    cent, K = _centroids_np(cfg)
    Dd = cfg.get("Dd", 2)
    init = co.uniform(jr.split(jr.key(42))[1], 100 * D).reshape(100, D)
    f0, d0 = co.score(task, init, Dd)
    g, f, d, _ = co.add(np.zeros((K, D)), np.full(K, -np.inf), np.zeros((K, Dd)), init, f0, d0, co.cells(d0, cent))
    key = jr.key(7)
    g, f, d, key, _, _ = co.map_elites_scan(g, f, d, cent, key, args.warmup, B, task)
    t0 = time.perf_counter()
    g, f, d, key, m, secs = co.map_elites_scan(g, f, d, cent, key, args.steps, B, task)
    dt = time.perf_counter() - t0
    value = args.steps * B / dt
    cores = co.get_threads()
    sample = f"{args.steps} generations of {B} offspring (workload batch {cfg['B']}), K={K} brute-force cells, after {args.warmup} warm-up"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config_dict(args, cfg, B_step=B),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "stage_seconds": {"emit": secs[0], "score": secs[1], "cells": secs[2], "add+metrics": secs[3]},
                         "label": "C/OpenMP restatement of QDax 0.5.1 (not jax[cpu]: jax is not installable here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "final_coverage": float(m[-1, 2]),
    }
    emit(line)


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


def _config_dict(args, cfg, B_step):
    K = cfg["K"] if cfg["cvt"] else cfg["grid"][0] * cfg["grid"][1]
    return {"workload": f"MAP-Elites {cfg['task']} {cfg['D']}-D, K={K} {'CVT(seeded uniform)' if cfg['cvt'] else 'grid'} cells, "
                        f"batch {cfg['B']} per generation, iso 0.05 / line 0.1 / clip [0,1]; {cfg['note']}",
            "name": args.config, "global_batch": cfg["B"], "batch_per_step": B_step, "genotype_dim": cfg["D"], "cells": K,
            "descriptor_dim": cfg.get("Dd", 2), "parallelism": f"dp{args.gpus} (offspring sharded, repertoire replicated)",
            "exchange": args.exchange if args.gpus > 1 else "none", "l2": "working set > L2 (offspring buffer 400 B x batch)"
            if cfg["B"] * cfg["D"] * 4 > 126e6 else "L2 flushed between steps by a 256 MB write" if args.flush_l2 else "working set fits L2; not flushed"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args, cfg):
    import functools

    import numpy as np
    import torch
    import torch.distributed as dist

    from qdax_b200 import _lib
    from qdax_b200 import random as qr
    from qdax_b200.core.containers.mapelites_repertoire import compute_euclidean_centroids
    from qdax_b200.core.distributed_map_elites import DistributedMAPElites
    from qdax_b200.core.emitters.mutation_operators import isoline_variation
    from qdax_b200.core.emitters.standard_emitters import MixingEmitter
    from qdax_b200.core.map_elites import MAPElites
    from qdax_b200.tasks.arm import arm_scoring_function
    from qdax_b200.tasks.standard_functions import rastrigin_scoring_function, sphere_scoring_function
    from qdax_b200.utils.metrics import default_qd_metrics

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

    D, task, B_total = cfg["D"], cfg["task"], cfg["B"]
    assert B_total % world == 0
    B = B_total // world
    Dd = cfg.get("Dd", 2)
    scoring = {"arm": arm_scoring_function, "rastrigin": rastrigin_scoring_function,
               "sphere": functools.partial(sphere_scoring_function, desc_dim=Dd)}[task]
    emitter = MixingEmitter(lambda x, k: x, functools.partial(isoline_variation, iso_sigma=0.05, line_sigma=0.1, minval=0.0, maxval=1.0), 1.0, B)
    metrics_fn = functools.partial(default_qd_metrics, qd_offset=0.0)
    if world > 1:
        me = DistributedMAPElites(scoring, emitter, metrics_fn, exchange=args.exchange)
    else:
        me = MAPElites(scoring, emitter, metrics_fn)
    if cfg["cvt"]:
        K = cfg["K"]
        cent = torch.from_numpy(_cvt_points(K, Dd)).to(dev)
    else:
        cent = compute_euclidean_centroids(cfg["grid"], 0.0, 1.0, device=dev)
        K = cent.shape[0]
    key = qr.key(42)
    key, subkey = qr.split(key)
    init = qr.uniform(subkey, (100, D), device=dev)          # identical on every rank -> identical replicas
    base = MAPElites(scoring, emitter, metrics_fn)
    rep, state, _ = base.init(init, cent, qr.key(1))
    rank_key = qr.split(qr.key(7), world)[rank]               # examples/distributed_mapelites.ipynb cell 23

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if args.flush_l2 else None

    def step(rep, key, timeline=False):
        """one generation, launch-only (device key chain on 1 GPU; host split + per-rank key on N GPUs)"""
        if flush is not None:
            flush.fill_(1)
        if world > 1:
            ks = qr.split(key)
            key, sub = ks[0], ks[1]
            m = torch.empty(4, dtype=torch.float32, device=dev)
            me._fused_distributed_generation(rep, fcfg, 3, sub, m)
            return key, m
        m = torch.empty(4, dtype=torch.float32, device=dev)
        me._fused_generation(rep, fcfg, 2, None, m, carry)      # scan_update step: carry key advanced on the host
        return key, m

    fcfg = me._fused_config(rep)
    assert fcfg is not None, "bench configuration must take the fused native path"
    rep = rep._clone_state()
    carry = np.array(rank_key, dtype=np.uint32)
    key = rank_key
    for _ in range(max(args.warmup, 3)):
        key, m = step(rep, key)
    barrier()
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record()
    for _ in range(5):                                        # rough ms per generation (after the one-time costs), untimed
        key, m = step(rep, key)
    w1.record()
    barrier()
    warm_ms = torch.tensor([w0.elapsed_time(w1) / 5], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(warm_ms, op=dist.ReduceOp.MAX)        # the same figure, hence the same step count, on every rank

    # ---- timed region: K generations, device timed (events around the whole region only).  The per-kernel timeline is
    # taken in a second, instrumented pass of the same K steps right after it: an event record between two kernels stops the
    # next launch from being staged behind the running one (~6 us per event, five per generation -- measured 0.411 vs
    # 0.444 ms per generation at N = 2), so it must not sit inside the headline loop.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # the recipe's clocks line (-lms 200): started before, killed after the timed region
    # nvidia-smi needs ~0.3 s to initialise and its first query stalls the GPU for ~0.5 ms -- 15 % of a 4 ms timed region (20
    # generations at N = 8).  Keep the GPU under load with untimed generations (the same count on every rank) while it does.
    for _ in range(min(int(600.0 / max(float(warm_ms[0]), 1e-3)) + 1, 20000)):
        key, m = step(rep, key)
    launches0 = _lib.launch_count
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        key, m = step(rep, key)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps      # CPU time to enqueue one generation (launch-only)
    ev1.record()
    barrier()
    launches = _lib.launch_count - launches0
    ms_total = ev0.elapsed_time(ev1)
    # instrumented pass: same loop with a CUDA event after every kernel (on the launching stream)
    me._timeline = []
    barrier()
    iv0, iv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iv0.record()
    for _ in range(args.steps):
        key, m = step(rep, key)
    iv1.record()
    barrier()
    tl, me._timeline = me._timeline, None
    ms_instrumented = iv0.elapsed_time(iv1)
    clocks = sampler.stop() if rank == 0 else None
    per_kernel = {}
    prev = None
    for label, e in tl:
        if label != "begin" and prev is not None:
            per_kernel.setdefault(label, []).append(prev.elapsed_time(e))
        prev = e
    kern_ms = {k: float(np.mean(v)) for k, v in per_kernel.items()}
    added_last = float(m[3])
    coverage = float(m[2])
    qd = float(m[0])
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t[0])
    value = args.steps * B_total / (ms_total * 1e-3)

    # ---- e2e: the public API call a user makes, per step: host key in, metrics read back to the host ----------
    # Two read-back disciplines, both inside the timed region: "pipelined" copies every step's metrics to PINNED host memory
    # with an async D2H + event and consumes them one step behind (what a training loop that logs metrics does), so the host
    # prepares step s+1 while step s runs; "blocking" calls .cpu() on the metrics after every step.
    e2e_steps = args.steps
    rep_e = rep
    hkey = key
    pinned = torch.empty((2, 4), dtype=torch.float32).pin_memory()
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    e2e = {}
    for mode in ("blocking", "pipelined"):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s_i in range(e2e_steps):
            ks = qr.split(hkey)                                   # host-side key chain, README.md:133
            hkey, sub = ks[0], ks[1]
            rep_e, state, md = me.update(rep_e, state, sub, donate=True)
            if mode == "blocking":
                host_metrics = me._last_metrics.cpu()             # D2H of the step's metrics (16 B) + sync every step
            else:
                pinned[s_i & 1].copy_(me._last_metrics, non_blocking=True)
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

    if os.environ.get("QDX_TRACE") and hasattr(_lib.lib(), "qdx_debug_xchg_trace"):      # timing experiment builds only
        import ctypes
        t8 = (ctypes.c_ulonglong * 8)()
        _lib.lib().qdx_debug_xchg_trace(t8, 0)
        n_e, n_p = max(t8[0], 1), max(t8[3], 1)
        sys.stderr.write("[xchg trace rank %d] elect launches %d: wait-for-flags %.2f us, elect kernel (CTA 0) %.2f us; publishes %d: %.2f us each\n"
                         % (rank, t8[0], t8[1] / n_e / 1e3, t8[2] / n_e / 1e3, t8[3], t8[4] / n_p / 1e3))
    consistent = True
    if world > 1:
        from qdax_b200 import parallel
        consistent = parallel.all_equal(rep.genotypes) and parallel.all_equal(rep.fitnesses)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel + the insert kernel ----------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "of measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "of fallback")
    ab = algorithmic_bytes(D, Dd)
    dom = max((k for k in kern_ms if k in ("generate", "cells")), key=lambda k: kern_ms[k])
    cells_kernel = "qdx_cells_tc_kernel (tcgen05 TF32 + exact re-rank)" if Dd >= 8 else "qdx_cells_bf_kernel"
    if dom == "generate":
        dom_bytes = ab["generate_per_offspring"] * B
    else:
        dom_bytes = (4 * Dd + 4) * B + K * Dd * 4
    dom_gbs = dom_bytes / (kern_ms[dom] * 1e-3) / 1e9
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        ent = tr.get("%s@%s@n%d" % ({"generate": "qdx_generate_kernel", "cells": "qdx_cells_tc_kernel"}[dom], args.config, world))
        traffic = ent["traffic_bytes"] if ent else None
    except Exception:
        pass
    roofline = {"kernel": {"generate": "qdx_generate_kernel", "cells": cells_kernel}[dom], "bound": "hbm", "achieved": dom_gbs,
                "peak": hbm_peak, "unit": "GB/s", "frac": dom_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": kern_ms[dom],
                "note": "ALU-bound kernel (one Threefry-2x32-20 block + erfinv per gene, sincos per joint): the HBM fraction is low by construction; see DESIGN.md section 6"}
    if dom == "generate" and args.config == "c3":
        roofline["issue_bound_evidence"] = {"source": "profiles/r1_generate_v4_ncu_summary.txt (ncu --set full of this kernel at this config)",
                                            "warp_instructions_per_offspring_row": 596.6, "issue_active_frac": 0.796, "alu_pipe_frac": 0.672,
                                            "fma_pipe_frac": 0.405, "dram_frac": 0.066}
    W = added_last
    commit_bytes = K * ab["commit_fixed_per_cell"] + W * ab["commit_per_winner"]
    insert = {"kernel": "qdx_commit_stream_kernel", "bound": "hbm", "achieved": commit_bytes / (kern_ms["commit"] * 1e-3) / 1e9, "peak": hbm_peak,
              "unit": "GB/s", "frac": commit_bytes / (kern_ms["commit"] * 1e-3) / 1e9 / hbm_peak, "winners_last_step": W,
              "algorithmic_bytes_per_launch": commit_bytes, "avg_launch_ms": kern_ms["commit"],
              "note": "K=10^4: <= 8.4 MB per launch, launch/latency-bound (SURVEY.md 8d caveat)"}

    if world == 1 and not args.no_insert_probe:
        insert["large_rows"] = insert_probe(dev, hbm_peak)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(cfg, args)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": _config_dict(args, cfg, B_step=B_total),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 16, "ms_per_step": e2e_ms / e2e_steps,
                "api": ("MAPElites.update(repertoire, emitter_state, key, donate=True)" if world == 1 else "DistributedMAPElites.update(...) per rank")
                       + " + every step's metrics copied to pinned host memory (async D2H + event) and read one step behind",
                "blocking_readback_value": e2e_blocking_value,
                "note": "inputs of a step are the 2-word RNG key (host) and the HBM-resident repertoire (carried state)"},
        "gpu_launches": launches, "kernel_ms": kern_ms,
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "kernel_ms_source": "instrumented pass of the same K steps right after the timed region (one CUDA event after every kernel); "
                            "ms per step there: %.4f" % (ms_instrumented / args.steps), "roofline": roofline, "insert_roofline": insert,
        "clocks": clocks, "replicas_bit_identical": consistent,
        "exchange_used": (getattr(me, "_exchange", "none") if world > 1 else "none"), "exchange_fallback": getattr(me, "exchange_fallback", None),
        "final": {"coverage": coverage, "qd_score": qd, "inserted_last_step": added_last},
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def insert_probe(dev, hbm_peak):
    """The insert (commit) kernel where its HBM roofline is meaningful (SURVEY.md 8d caveat): BASELINE configs[3] shape --
    K = 50 000 cells, D = 1000 (4 KB rows), Dd = 32 -- 65 536 offspring into an EMPTY repertoire, so ~30 000 winner rows
    (~250 MB algorithmic) move in one launch.  Outside the timed region of the headline; CUDA events around ONE launch after
    an L2-evicting read pass (median of 5), and around a train of 8 back-to-back launches into 8 empty repertoires."""
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
    return {"workload": "K=50000 cells, D=1000, Dd=32, 65536 offspring into an empty repertoire (BASELINE configs[3] cold start)",
            "winners": W, "algorithmic_bytes_per_launch": nbytes, "launch_ms": ms1, "achieved": nbytes / (ms1 * 1e-3) / 1e9,
            "peak": hbm_peak, "unit": "GB/s", "frac": nbytes / (ms1 * 1e-3) / 1e9 / hbm_peak,
            "train_of_8_launch_ms": mst, "train_of_8_frac": nbytes / (mst * 1e-3) / 1e9 / hbm_peak,
            "timing": "CUDA events; single launch after an L2-evicting read pass (median of 5) / 8 back-to-back launches"}


def cpu_baseline(cfg, args):
    """Oracle port timed on the host cores on a bounded sample (about 10-30 s of CPU work)."""
    import numpy as np

    from oracle import c_oracle as co
    from oracle import jax_prng as jr

    co.build()
    D, task = cfg["D"], cfg["task"]
    B = min(cfg["B"], args.cpu_sample)
    cent, K = _centroids_np(cfg)
    Dd = cfg.get("Dd", 2)
    init = co.uniform(jr.split(jr.key(42))[1], 100 * D).reshape(100, D)
    f0, d0 = co.score(task, init, Dd)
    g, f, d, _ = co.add(np.zeros((K, D)), np.full(K, -np.inf), np.zeros((K, Dd)), init, f0, d0, co.cells(d0, cent))
    t0 = time.perf_counter()
    g, f, d, key, _, _ = co.map_elites_scan(g, f, d, cent, jr.key(7), 1, B, task)
    t1 = time.perf_counter() - t0
    n = int(min(60, max(2, 12.0 / max(t1, 1e-3))))
    t0 = time.perf_counter()
    g, f, d, key, m, secs = co.map_elites_scan(g, f, d, cent, key, n, B, task)
    dt = time.perf_counter() - t0
    return {"value": n * B / dt, "unit": UNIT, "cores": co.get_threads(), "kind": "port",
            "sample": f"{n} generations of {B} offspring (1/{cfg['B'] // B} of the batch), K={K} brute-force cells, {dt:.1f} s",
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
    ap.add_argument("--cpu-sample", type=int, default=1 << 16, help="offspring per generation in the CPU arm")
    ap.add_argument("--flush-l2", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-insert-probe", action="store_true", help="skip the large-row insert-kernel measurement (N=1 only)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_gpu(args, cfg)


if __name__ == "__main__":
    main()
