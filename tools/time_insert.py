"""Insert (commit) kernel where its HBM roofline is meaningful (SURVEY 8d caveat): config-4 cold start --
K = 50 000 cells, D = 1000 (4 KB genotype rows), Dd = 32, B = 65 536 offspring into an EMPTY repertoire, so W ~ 36 000
winner rows (~290 MB of algorithmic traffic) move in one launch.  Also the steady-state c3 shape for contrast."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from qdax_b200 import _native
dev = torch.device("cuda:0")
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0
out = {}
for name, (K, D, Dd, B) in {"c4_cold_start": (50000, 1000, 32, 65536), "c4_cold_start_b262144": (50000, 1000, 32, 262144),
                            "c3_cold_start": (10000, 100, 2, 1 << 20)}.items():
    rng = np.random.default_rng(0)
    cent = torch.from_numpy(rng.random((K, Dd)).astype(np.float32)).to(dev)
    g = torch.rand(B, D, device=dev); d = torch.rand(B, Dd, device=dev); f = torch.randn(B, device=dev)
    cells = _native.cells(d, cent, None)
    ws = _native.Workspace(K, dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    times = []
    for rep in range(6):
        rep_g = torch.zeros(K, D, device=dev); rep_f = torch.full((K,), float("-inf"), device=dev); rep_d = torch.zeros(K, Dd, device=dev)
        m = torch.empty(4, device=dev)
        _native.offer_cells(cells, f, ws, rep_f)
        flush.fill_(1); _ = flush.view(torch.int32).sum()        # L2 flush: evict with a READ pass so no dirty lines are left to write back
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _native.commit(ws, g, f, d, rep_g, rep_f, rep_d, metrics_out=m)
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    W = float(m[3])
    ms = float(np.median(times[1:]))
    bytes_ = K * 20 + W * 2 * (4 * D + 4 * Dd + 4)
    out[name] = {"K": K, "D": D, "B": B, "winners": W, "commit_ms": ms, "algorithmic_bytes": bytes_, "GBps": bytes_ / (ms * 1e-3) / 1e9,
                 "frac_of_measured_hbm_peak": bytes_ / (ms * 1e-3) / 1e9 / peak, "coverage": float(m[2])}
print(json.dumps(out, indent=1))
