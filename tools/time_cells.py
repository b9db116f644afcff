"""Time the cell-assignment kernels (CUDA events, 10 reps after 3 warm-up): tcgen05 path vs FP32 brute force."""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from qdax_b200 import _native
dev = torch.device("cuda:0")
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
tc_only = "tc_only" in sys.argv      # skip the (slow) brute-force reference: timing experiments of the tensor-core kernel
for (B, K, Dd) in [(65536, 50000, 32)] if tc_only else [(65536, 50000, 32), (65536, 10000, 2), (1 << 20, 10000, 2)]:
    rng = np.random.default_rng(0)
    cent = torch.from_numpy(rng.random((K, Dd)).astype(np.float32)).to(dev)
    desc = torch.from_numpy(rng.random((B, Dd)).astype(np.float32)).to(dev)
    o = torch.empty(B, dtype=torch.int32, device=dev)
    r = {"bruteforce_ms": float("nan") if tc_only else timeit(lambda: _native.cells(desc, cent, None, out=o, allow_tc=False), 3 if Dd > 4 else 10)}
    ref = o.clone()
    if Dd >= 8:
        r["tc_ms"] = timeit(lambda: _native.cells_tc(desc, cent, out=o))
        r["equal"] = bool(torch.equal(o, ref))
        scratch_probe = None
    r["pairs_per_s_bf"] = B * K / (r["bruteforce_ms"] * 1e-3)
    if "tc_ms" in r:
        r["pairs_per_s_tc"] = B * K / (r["tc_ms"] * 1e-3)
        r["tflops_tc"] = 2.0 * B * K * 32 / (r["tc_ms"] * 1e-3) / 1e12
    out[f"B{B}_K{K}_Dd{Dd}"] = r
print(json.dumps(out, indent=1))
