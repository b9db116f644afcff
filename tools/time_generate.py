"""The generate kernel alone (select + isoline + arm scoring + grid cell + offer) at the per-GPU batch sizes of the 1 / 2 / 4 / 8
GPU runs of config c3: average duration of back-to-back launches (CUDA events), and -- with a library built with
-DQDX_GEN_TRACE=1 -- where the time of ONE launch goes (first / last CTA start, first / last CTA end, globaltimer)."""
import ctypes as C, functools, json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from qdax_b200 import _lib, _native, random as qr
from qdax_b200.core.containers.mapelites_repertoire import compute_euclidean_centroids
from qdax_b200.core.emitters.mutation_operators import isoline_variation
from qdax_b200.core.emitters.standard_emitters import MixingEmitter
from qdax_b200.core.map_elites import MAPElites
from qdax_b200.tasks.arm import arm_scoring_function
from qdax_b200.utils.metrics import default_qd_metrics

dev = torch.device("cuda:0")
D = 100
em = MixingEmitter(lambda x, k: x, functools.partial(isoline_variation, iso_sigma=0.05, line_sigma=0.1, minval=0.0, maxval=1.0), 1.0, 1 << 17)
me = MAPElites(arm_scoring_function, em, functools.partial(default_qd_metrics, qd_offset=0.0))
cent = compute_euclidean_centroids((100, 100), 0.0, 1.0, device=dev)
rep, state, _ = me.init(qr.uniform(qr.key(1), (100, D), device=dev), cent, qr.key(2))
(rep, state, key), _ = me.scan((rep, state, qr.key(3)), 30, donate=False)          # steady state repertoire
ws = rep._workspace(); rep_f = rep.fitnesses.reshape(-1); grid = rep._grid()
_native.ensure_selection(rep_f, ws)
trace = hasattr(_lib.lib(), "qdx_debug_gen_trace")
out = {}
for B in [1 << 17, 1 << 18, 1 << 19, 1 << 20] + [int(a) for a in sys.argv[1:]]:
    g = torch.empty(B, D, device=dev); f = torch.empty(B, device=dev); d = torch.empty(B, 2, device=dev); c = torch.empty(B, dtype=torch.int32, device=dev)
    keys = _native.host_generation_keys(_native.KEYMODE_UPDATE, qr.key(5))
    run = lambda: _native.generate(rep.genotypes, rep_f, cent, ws, B, 0.05, 0.1, 0.0, 1.0, "arm", 2, grid, True, 0, True, g, f, d, c, gen_keys=keys)
    for _ in range(3): run()
    torch.cuda.synchronize()
    R = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(R): run()
    e1.record(); torch.cuda.synchronize()
    res = {"ms_back_to_back": e0.elapsed_time(e1) / R, "ideal_from_2^20_ms": None}
    singles = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); run(); b.record(); torch.cuda.synchronize(); singles.append(a.elapsed_time(b))
    res["ms_single_launch"] = float(np.median(singles))
    if trace:
        _lib.lib().qdx_debug_gen_trace(None, 1)
        run(); t = (C.c_ulonglong * 4)(); _lib.lib().qdx_debug_gen_trace(t, 0); t = list(t)
        res["trace_us"] = {"last_cta_start": (t[1] - t[0]) / 1e3, "first_cta_end": (t[2] - t[0]) / 1e3, "last_cta_end": (t[3] - t[0]) / 1e3}
    ws_keys = ws.keytab(); ws_keys.zero_()
    out[str(B)] = res
base = out[str(1 << 20)]["ms_back_to_back"]
for k, v in out.items(): v["ideal_from_2^20_ms"] = base * int(k) / (1 << 20)
print(json.dumps(out))
