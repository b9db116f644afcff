#!/bin/bash
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e_n2.json 2> gpurun_out/r2e_n2.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2e_n2.json")); print("n2", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], d["clocks"], d["replicas_bit_identical"], d["gpu_launches"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2e_n2.err").read()[-1500:])
PY
