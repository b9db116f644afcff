#!/bin/bash
# pytree genotypes (8f rank 2): GPU parity tests + regression check of the headline bench
timeout 900 python -m pytest tests/test_gpu_pytree.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r1t_pytree.log; cat gpurun_out/r1t_pytree.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1t_tests.log; cat gpurun_out/r1t_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1t_n1.json 2> gpurun_out/r1t_n1.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r1t_n1.json")); print("n1", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], d["final"])
except Exception as e:
    print("n1 ERR", e); print(open("gpurun_out/r1t_n1.err").read()[-1200:])
PY
