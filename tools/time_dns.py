"""BASELINE config 5: Dominated Novelty Search, rastrigin 100-D, population 100 000 (valid), batch 1024, k = 3:
time one DominatedNoveltyRepertoire.add on the GPU (CUDA events) and check it against the oracle at reduced N."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from qdax_b200 import _native, random as qr
from qdax_b200.tasks.standard_functions import rastrigin_scoring_function
dev = torch.device("cuda:0")
P, B, D, k = 100000, 1024, 100, 3
g = qr.uniform(qr.key(2), (P + B, D), device=dev)
f, d, _ = rastrigin_scoring_function(g)
pg, pf, pd = g[:P].contiguous(), f[:P].contiguous(), d[:P].contiguous()
bg, bf, bd = g[P:].contiguous(), f[P:].contiguous(), d[P:].contiguous()
for _ in range(2): out = _native.dns_add(pg, pf, pd, bg, bf, bd, k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): out = _native.dns_add(pg, pf, pd, bg, bf, bd, k)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
N = P + B
res = {"config": "c5 DNS rastrigin 100-D, population 100000, batch 1024, k=3", "ms_per_add": ms, "candidates": N, "pairs": N * N,
       "pairs_per_s": N * N / (ms * 1e-3), "offspring_per_s": B / (ms * 1e-3)}
if "--check" in sys.argv:
    from oracle import c_oracle as co
    n = 6000
    t0 = time.perf_counter()
    G, F, Dn, meta, surv = co.dns_add(pg[:n].cpu().numpy(), pf[:n].cpu().numpy(), pd[:n].cpu().numpy(), bg.cpu().numpy(), bf.cpu().numpy(), bd.cpu().numpy(), k)
    res["cpu_port_s_at_N%d" % (n + B)] = time.perf_counter() - t0
    o = _native.dns_add(pg[:n].contiguous(), pf[:n].contiguous(), pd[:n].contiguous(), bg, bf, bd, k)
    res["bit_exact_vs_oracle_at_reduced_N"] = bool(np.array_equal(o[4].cpu().numpy(), surv) and np.array_equal(o[0].cpu().numpy(), G))
print(json.dumps(res))
