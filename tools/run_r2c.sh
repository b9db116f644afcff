#!/bin/bash
# headline loop without per-kernel events: c3 at N=1 and N=2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-insert-probe > gpurun_out/r2c_n1.json 2> gpurun_out/r2c_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 --exchange p2p > gpurun_out/r2c_n2.json 2> gpurun_out/r2c_n2.err
python - <<'PY'
import json
for n in ["n1", "n2"]:
    try:
        d=json.load(open(f"gpurun_out/r2c_{n}.json")); print(n, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "host enqueue %.4f"%d["host_enqueue_ms_per_step"], d["kernel_ms_source"][-28:], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], "blocking %.4g"%d["e2e"]["blocking_readback_value"], d["replicas_bit_identical"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r2c_{n}.err").read()[-1500:])
PY
