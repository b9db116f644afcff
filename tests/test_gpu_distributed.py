"""Multi-GPU parity (one process per GPU, NCCL): DistributedMAPElites with both exchanges yields, on every rank, the
repertoire the oracle computes for the concatenated batch (global offspring index = rank * B_dev + i), bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import functools, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["QDX_ROOT"])
from oracle import c_oracle as co, jax_prng as jr, qdax_numpy as qn
from qdax_b200 import parallel, random as qr
from qdax_b200.core.containers.mapelites_repertoire import compute_euclidean_centroids
from qdax_b200.core.distributed_map_elites import DistributedMAPElites
from qdax_b200.core.emitters.mutation_operators import isoline_variation
from qdax_b200.core.emitters.standard_emitters import MixingEmitter
from qdax_b200.tasks.arm import arm_scoring_function
from qdax_b200.tasks.standard_functions import rastrigin_scoring_function
from qdax_b200.utils.metrics import default_qd_metrics

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
N = lambda t: t.detach().cpu().numpy()

def run(task, scoring, cent_t, B_dev, D, iters, exchange, donate=False):
    em = MixingEmitter(lambda x, k: x, functools.partial(isoline_variation, iso_sigma=0.05, line_sigma=0.1, minval=0.0, maxval=1.0), 1.0, B_dev)
    me = DistributedMAPElites(scoring, em, functools.partial(default_qd_metrics, qd_offset=0.0), exchange=exchange)
    init_all = qr.uniform(jr.key(11), (world * 16, D), device=dev)
    init_fn = me.get_distributed_init_fn(cent_t)
    rep, state, _ = init_fn(init_all[rank * 16:(rank + 1) * 16].contiguous(), jr.key(0))   # sharded init, gathered inside
    keys = qr.split(jr.key(5), world)                                              # notebook cell 23
    key = keys[rank]
    cent = N(cent_t)
    K = cent.shape[0]
    f0, d0 = co.score(task, N(init_all))
    g, f, d, _ = co.add(np.zeros((K, D)), np.full(K, -np.inf), np.zeros((K, 2)), N(init_all), f0, d0, co.cells(d0, cent))
    assert np.array_equal(N(rep.genotypes), g) and np.array_equal(N(rep.fitnesses).ravel(), f)
    okeys = [np.array(k) for k in keys]
    for it in range(iters):
        ks = qr.split(key); key, sub = ks[0], ks[1]
        rep, state, m = me.update(rep, state, sub, donate=donate)
        subs = []
        for r in range(world):
            s2 = jr.split(okeys[r]); okeys[r] = s2[0]; subs.append(s2[1])
        g, f, d, *_ = co.distributed_update(g, f, d, cent, np.stack(subs), B_dev, task)
        assert np.array_equal(N(rep.fitnesses).ravel(), f), (task, exchange, it, "fitnesses")
        assert np.array_equal(N(rep.genotypes), g) and np.array_equal(N(rep.descriptors), d), (task, exchange, it)
        ref = co.metrics(f, 0.0)
        assert np.allclose([float(m["qd_score"]), float(m["max_fitness"]), float(m["coverage"])], ref, rtol=1e-5)
    assert parallel.all_equal(rep.genotypes) and parallel.all_equal(rep.fitnesses)
    # scan() API == the loop above
    return rep

for exchange in ("allgather", "winners", "regen", "p2p"):
    grid = compute_euclidean_centroids((16, 16), 0.0, 1.0, device=dev)
    run("arm", arm_scoring_function, grid, 300, 20, 4, exchange)
    cvt = torch.from_numpy(np.random.default_rng(0).random((500, 2)).astype(np.float32)).to(dev)
    run("rastrigin", rastrigin_scoring_function, cvt, 257, 100, 3, exchange)
# soak: the bench path (donated repertoire, one C call per generation, peer-memory exchange with offspring blocks) for 40 generations,
# checked against the oracle after every generation; a shard size that is not a multiple of the tile height
grid = compute_euclidean_centroids((24, 24), 0.0, 1.0, device=dev)
run("arm", arm_scoring_function, grid, 1000 + 8 * 3, 40, 40, "p2p", donate=True)
dist.barrier()
if rank == 0:
    print("DISTRIBUTED_OK", world)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_map_elites_nccl(tmp_path, world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} CUDA devices")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, QDX_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29517 + world), str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DISTRIBUTED_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
