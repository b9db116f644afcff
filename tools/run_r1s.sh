#!/bin/bash
# 2-GPU check after the commit-kernel rewrite: distributed parity tests (all four exchanges) + c3 at N=2
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1s_tests.log; cat gpurun_out/r1s_tests.log
for ex in p2p regen; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 --exchange $ex > gpurun_out/r1s_n2_$ex.json 2> gpurun_out/r1s_n2_$ex.err
done
python - <<'PY'
import json
for n in ["n2_p2p", "n2_regen"]:
    try:
        d=json.load(open(f"gpurun_out/r1s_{n}.json")); print(n, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], d["replicas_bit_identical"], d.get("exchange_used"), d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r1s_{n}.err").read()[-1200:])
PY
