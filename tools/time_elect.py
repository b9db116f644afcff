"""Single-GPU timing harness for the multi-GPU tail of a generation (elect -> commit mode 2) at the c3 shape:
the key table of one steady-state generation is produced by generate(offer), then qdx_elect_winners is timed in a loop
(it does not consume the table), then qdx_commit(mode 2)."""
import functools, json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from qdax_b200 import _native, random as qr
from qdax_b200.core.containers.mapelites_repertoire import compute_euclidean_centroids
from qdax_b200.core.emitters.mutation_operators import isoline_variation
from qdax_b200.core.emitters.standard_emitters import MixingEmitter
from qdax_b200.core.map_elites import MAPElites
from qdax_b200.tasks.arm import arm_scoring_function
from qdax_b200.utils.metrics import default_qd_metrics

dev = torch.device("cuda:0")
B, D, Dd = int(sys.argv[1]) if len(sys.argv) > 1 else 131072, 100, 2
em = MixingEmitter(lambda x, k: x, functools.partial(isoline_variation, iso_sigma=0.05, line_sigma=0.1, minval=0.0, maxval=1.0), 1.0, B)
me = MAPElites(arm_scoring_function, em, functools.partial(default_qd_metrics, qd_offset=0.0))
cent = compute_euclidean_centroids((100, 100), 0.0, 1.0, device=dev)
rep, st, _ = me.init(qr.uniform(qr.key(1), (100, D), device=dev), cent, qr.key(2))
(rep, st, key), _ = me.scan((rep, st, qr.key(3)), 30, donate=True)          # steady state
K = cent.shape[0]
ws = rep._workspace(); rep_f = rep.fitnesses.reshape(-1)
buf = me._offspring_buffers(B, D, Dd, dev)
_native.select_prepare(rep_f, ws, _native.KEYMODE_UPDATE, qr.key(9), rank_slot=0)
_native.generate(rep.genotypes, rep_f, rep.centroids, ws, B, 0.05, 0.1, 0.0, 1.0, "arm", Dd, rep._grid(), True, 0, True, buf["g"], buf["f"], buf["d"], buf["c"])
stage = torch.zeros(K, D + Dd + 1, device=dev)
flat = stage.reshape(-1); sg = flat[:K * D].view(K, D); sd = flat[K * D:K * (D + Dd)].view(K, Dd); sf = flat[K * (D + Dd):].view(K)
elected = int((ws.keytab() != 0).sum())
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_elect = timeit(lambda: _native.elect_winners(ws, rep.genotypes, "arm", Dd, B, 1, 0.05, 0.1, 0.0, 1.0, True, sg, sf, sd))
t_regen = timeit(lambda: _native.call("qdx_regenerate_winners", ws.ptr, __import__("ctypes").c_int64(K), __import__("ctypes").c_int64(D), __import__("ctypes").c_int64(B), __import__("ctypes").c_int32(1), _native._ptr(rep.genotypes), __import__("ctypes").c_float(0.05), __import__("ctypes").c_float(0.1), __import__("ctypes").c_int32(1), __import__("ctypes").c_float(0.0), __import__("ctypes").c_int32(1), __import__("ctypes").c_float(1.0), __import__("ctypes").c_int32(1), _native._ptr(sg), _native._stream()))
# check against the rows the owner produced
kt = ws.keytab().clone()
idx = (~kt & 0x7FFFFFFF).long()
cells = torch.nonzero(kt != 0).reshape(-1)
ok = bool(torch.equal(sg[cells], buf["g"][idx[cells]]) and torch.equal(sf[cells], buf["f"][idx[cells]]) and torch.equal(sd[cells], buf["d"][idx[cells]]))
m = torch.empty(4, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); _native.commit(ws, sg, sf, sd, rep.genotypes, rep_f, rep.descriptors, metrics_out=m, mode=2); e1.record(); torch.cuda.synchronize()
print(json.dumps({"B_dev": B, "elected_cells": elected, "elect_ms": t_elect, "regen_only_ms": t_regen, "commit_mode2_ms": e0.elapsed_time(e1), "regenerated_rows_bit_identical": ok}))
