#!/bin/bash
# 1 -> 8 GPU scaling series of bench.py (same launch line the driver uses); tools/scale_run.sh [tag] [exchange]
TAG=${1:-scale}; EX=${2:-p2p}
timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/${TAG}_disttests.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 --exchange $EX > gpurun_out/${TAG}_n$n.json 2> gpurun_out/${TAG}_n$n.err
done
for ex in regen allgather; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29699 bench.py --gpus 8 --steps 20 --warmup 5 --exchange $ex > gpurun_out/${TAG}_n8_$ex.json 2> gpurun_out/${TAG}_n8_$ex.err
done
cat gpurun_out/${TAG}_disttests.log
python - <<PY
import json
for n in ["n1","n2","n4","n8","n8_regen","n8_allgather"]:
    try:
        d=json.load(open(f"gpurun_out/${TAG}_{n}.json")); print(n, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], d["replicas_bit_identical"], d.get("exchange_used"), d.get("exchange_fallback"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/${TAG}_{n}.err").read()[-800:])
PY
