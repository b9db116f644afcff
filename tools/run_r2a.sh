#!/bin/bash
# where the multi-GPU tail goes: c3 at N=2 with the exchange-trace build
QDX_TRACE=1 QDX_LIB_PATH=$PWD/qdax_b200/libqdx_xtrace.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 --exchange p2p > gpurun_out/r2a_n2.json 2> gpurun_out/r2a_n2.err
grep "xchg trace" gpurun_out/r2a_n2.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2a_n2.json")); print("n2", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], "blocking %.4g"%d["e2e"]["blocking_readback_value"], d["replicas_bit_identical"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2a_n2.err").read()[-1500:])
PY
