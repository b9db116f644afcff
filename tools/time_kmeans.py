"""compute_cvt_centroids(backend="gpu") at the BASELINE configs[3] tessellation: K = 50 000 centroids in 32-D."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from qdax_b200 import random as qr
from qdax_b200.core.containers.mapelites_repertoire import lloyd_cvt_centroids
dev = torch.device("cuda:0")
out = {}
for name, (N, K, Dd, iters) in {"c4_k50000_d32": (1 << 20, 50000, 32, 20), "c2_k10000_d2": (1 << 20, 10000, 2, 20)}.items():
    x = qr.uniform(qr.key(0), (N, Dd), device=dev)
    lloyd_cvt_centroids(x, K, 1); torch.cuda.synchronize()
    t0 = time.perf_counter()
    cent, ran = lloyd_cvt_centroids(x, K, iters)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[name] = {"samples": N, "centroids": K, "desc_dim": Dd, "iterations": ran, "seconds": dt, "ms_per_iteration": 1e3 * dt / ran}
print(json.dumps(out, indent=1))
