#!/bin/bash
# 1 -> 8 GPU scaling series of bench.py (same launch line the driver uses)
EX=${1:-regen}
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 --exchange $EX > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29699 bench.py --gpus 8 --steps 20 --warmup 5 --exchange allgather > gpurun_out/scale_n8_allgather.json 2> gpurun_out/scale_n8_allgather.err
python - <<'PY'
import json
for n in ["n1","n2","n4","n8","n8_allgather"]:
    try:
        d=json.load(open(f"gpurun_out/scale_{n}.json")); print(n, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], d["replicas_bit_identical"], d["clocks"])
    except Exception as e: print(n, "ERR", e)
PY
