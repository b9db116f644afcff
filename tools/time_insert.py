"""Insert (commit) kernel where its HBM roofline is meaningful (SURVEY 8d caveat): config-4 cold start --
K = 50 000 cells, D = 1000 (4 KB genotype rows), Dd = 32, B = 65 536 offspring into an EMPTY repertoire, so W ~ 30 000
winner rows (~250 MB of algorithmic traffic) move in one launch.  Also the c3 shape for contrast.

Two timings per shape, both CUDA events on the launching stream:
  * single: ONE launch after an L2-evicting read pass (includes the event pair's and the launch's fixed cost);
  * train:  R launches back to back, each into its own empty repertoire / workspace (offers done before the timed
            region) and reading one of two offspring buffers alternately (each larger than L2) -- the average launch
            duration over a timed region, the figure bench.py's roofline contract asks for.
--trace (library built with -DQDX_COMMIT_TRACE=1): per-phase globaltimer stamps of the single launch."""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from qdax_b200 import _native, _lib
dev = torch.device("cuda:0")
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
trace = "--trace" in sys.argv and hasattr(_lib.lib(), "qdx_debug_commit_trace")
R = 8
out = {}
shapes = {"c4_cold_start": (50000, 1000, 32, 65536), "c4_cold_start_b262144": (50000, 1000, 32, 262144),
          "c3_cold_start": (10000, 100, 2, 1 << 20)}
for name, (K, D, Dd, B) in shapes.items():
    rng = np.random.default_rng(0)
    cent = torch.from_numpy(rng.random((K, Dd)).astype(np.float32)).to(dev)
    gs = [torch.rand(B, D, device=dev) for _ in range(2)]
    g = gs[0]; d = torch.rand(B, Dd, device=dev); f = torch.randn(B, device=dev)
    cells = _native.cells(d, cent, None)
    ws = _native.Workspace(K, dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    times, tr = [], None
    for rep in range(6):
        rep_g = torch.zeros(K, D, device=dev); rep_f = torch.full((K,), float("-inf"), device=dev); rep_d = torch.zeros(K, Dd, device=dev)
        m = torch.empty(4, device=dev)
        _native.offer_cells(cells, f, ws, rep_f)
        flush.fill_(1); _ = flush.view(torch.int32).sum()        # L2 flush: evict with a READ pass so no dirty lines are left to write back
        if trace:
            _lib.lib().qdx_debug_commit_trace(None, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _native.commit(ws, g, f, d, rep_g, rep_f, rep_d, metrics_out=m)
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        if trace:
            t = (C.c_ulonglong * 8)(); _lib.lib().qdx_debug_commit_trace(t, 0); t = list(t)
            tr = {"phase1_first_done_us": (t[1] - t[0]) / 1e3, "phase1_last_done_us": (t[2] - t[0]) / 1e3, "barrier_last_out_us": (t[3] - t[0]) / 1e3,
                  "stream_first_done_us": (t[4] - t[0]) / 1e3, "stream_last_done_us": (t[5] - t[0]) / 1e3, "end_us": (t[6] - t[0]) / 1e3,
                  "metrics_written_us": (t[7] - t[0]) / 1e3}
    W = float(m[3])
    ms = float(np.median(times[1:]))
    bytes_ = K * 20 + W * 2 * (4 * D + 4 * Dd + 4)
    # ---- train of R launches
    del rep_g
    reps = [(torch.zeros(K, D, device=dev), torch.full((K,), float("-inf"), device=dev), torch.zeros(K, Dd, device=dev), _native.Workspace(K, dev))
            for _ in range(R)]
    train = []
    for it in range(4):
        for (rg, rf, rd, w) in reps:
            rf.fill_(float("-inf")); _native.offer_cells(cells, f, w, rf)
        flush.fill_(1); _ = flush.view(torch.int32).sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r, (rg, rf, rd, w) in enumerate(reps):
            _native.commit(w, gs[r & 1], f, d, rg, rf, rd, metrics_out=m)
        e1.record(); torch.cuda.synchronize()
        train.append(e0.elapsed_time(e1) / R)
    ms_t = float(np.median(train[1:]))
    out[name] = {"K": K, "D": D, "B": B, "winners": W, "commit_ms": ms, "algorithmic_bytes": bytes_, "GBps": bytes_ / (ms * 1e-3) / 1e9,
                 "frac_of_measured_hbm_peak": bytes_ / (ms * 1e-3) / 1e9 / peak, "coverage": float(m[2]),
                 "train_launches": R, "train_ms_per_launch": ms_t, "train_GBps": bytes_ / (ms_t * 1e-3) / 1e9,
                 "train_frac_of_measured_hbm_peak": bytes_ / (ms_t * 1e-3) / 1e9 / peak}
    if tr:
        out[name]["trace"] = tr
    del reps, gs, g
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
