#!/bin/bash
# trimmed 2 -> 8 GPU series on one 8 x B200 box: 2- and 4-GPU parity tests, c3 with the p2p exchange at N = 2, 4, 8, the
# reference-faithful all-gather at N = 8
TAG=${1:-scale3}
timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/${TAG}_disttests.log
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 --exchange p2p > gpurun_out/${TAG}_n$n.json 2> gpurun_out/${TAG}_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29699 bench.py --gpus 8 --steps 20 --warmup 5 --exchange allgather > gpurun_out/${TAG}_n8_allgather.json 2> gpurun_out/${TAG}_n8_allgather.err
cat gpurun_out/${TAG}_disttests.log
python - <<PY
import json
for n in ["n2","n4","n8","n8_allgather"]:
    try:
        d=json.load(open(f"gpurun_out/${TAG}_{n}.json")); print(n, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "host %.3f"%d["host_enqueue_ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], "blocking %.4g"%d["e2e"]["blocking_readback_value"], d["replicas_bit_identical"], d.get("exchange_used"), d.get("exchange_fallback"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/${TAG}_{n}.err").read()[-800:])
PY
