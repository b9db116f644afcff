#!/bin/bash
# GPU k-means (8f rank 4): parity tests + timing; bench line with the live large-row insert probe
timeout 900 python -m pytest tests/test_kmeans.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r1u_kmeans.log; cat gpurun_out/r1u_kmeans.log
timeout 600 python tools/time_kmeans.py > gpurun_out/r1u_kmeans.json 2> gpurun_out/r1u_kmeans.err; cat gpurun_out/r1u_kmeans.json; tail -3 gpurun_out/r1u_kmeans.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1u_n1.json 2> gpurun_out/r1u_n1.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r1u_n1.json")); print("n1", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], d["insert_roofline"].get("large_rows"))
except Exception as e:
    print("n1 ERR", e); print(open("gpurun_out/r1u_n1.err").read()[-1500:])
PY
