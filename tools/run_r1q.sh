#!/bin/bash
# streaming commit with dynamic job batches: parity tests + timing of the batch-size variants + per-phase trace
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1q_tests.log; cat gpurun_out/r1q_tests.log
for v in base trace g22 g84 g41; do
  lib=$PWD/qdax_b200/libqdx_$v.so; [ "$v" = base ] && lib=$PWD/qdax_b200/libqdx.so
  QDX_LIB_PATH=$lib timeout 300 python tools/time_insert.py --trace > gpurun_out/r1q_insert_$v.json 2> gpurun_out/r1q_insert_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r1q_insert_$v.json"))
    for k,x in d.items(): print("$v", k, "single %.4f ms %.3f | train %.4f ms %.3f"%(x["commit_ms"], x["frac_of_measured_hbm_peak"], x["train_ms_per_launch"], x["train_frac_of_measured_hbm_peak"]), x.get("trace"))
except Exception as e:
    print("$v ERR", e); print(open("gpurun_out/r1q_insert_$v.err").read()[-1500:])
PY
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1q_n1.json 2> gpurun_out/r1q_n1.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r1q_n1.json")); print("n1", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], d["final"])
except Exception as e:
    print("n1 ERR", e); print(open("gpurun_out/r1q_n1.err").read()[-1200:])
PY
