#!/bin/bash
# balanced elect kernel: 2-GPU parity tests, c3 at N=2 (normal + exchange-trace build)
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 --exchange p2p > gpurun_out/r2b_n2.json 2> gpurun_out/r2b_n2.err
QDX_TRACE=1 QDX_LIB_PATH=$PWD/qdax_b200/libqdx_xtrace.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 --exchange p2p > gpurun_out/r2b_n2t.json 2> gpurun_out/r2b_n2t.err
grep "xchg trace" gpurun_out/r2b_n2t.err
python - <<'PY'
import json
for n in ["n2", "n2t"]:
    try:
        d=json.load(open(f"gpurun_out/r2b_{n}.json")); print(n, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], "blocking %.4g"%d["e2e"]["blocking_readback_value"], d["replicas_bit_identical"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r2b_{n}.err").read()[-1500:])
PY
