#!/bin/bash
# session-2 check: GPU tests (incl. 2-GPU NCCL + p2p parity), 1-GPU bench, 2-GPU bench with both exchanges
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1b_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1b_n1.json 2> gpurun_out/r1b_n1.err
python bench.py --config c1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r1b_c1.json 2> gpurun_out/r1b_c1.err
for ex in regen p2p; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 --exchange $ex > gpurun_out/r1b_n2_$ex.json 2> gpurun_out/r1b_n2_$ex.err
done
cat gpurun_out/r1b_tests.log
python - <<'PY'
import json
for n in ["n1","c1","n2_regen","n2_p2p"]:
    try:
        d=json.load(open(f"gpurun_out/r1b_{n}.json")); print(n, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], d["replicas_bit_identical"], d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r1b_{n}.err").read()[-1500:])
PY
