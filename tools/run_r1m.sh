#!/bin/bash
# insert-kernel experiments: base / per-phase trace / ordinary launch / descriptors copied in phase 2
for v in base trace nocoop descp2 descp2trace; do
  lib=$PWD/qdax_b200/libqdx_$v.so; [ "$v" = base ] && lib=$PWD/qdax_b200/libqdx.so
  QDX_LIB_PATH=$lib timeout 300 python tools/time_insert.py --trace > gpurun_out/r1m_insert_$v.json 2> gpurun_out/r1m_insert_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r1m_insert_$v.json"))
    for k,x in d.items(): print("$v", k, "single %.4f ms %.3f | train %.4f ms %.3f"%(x["commit_ms"], x["frac_of_measured_hbm_peak"], x["train_ms_per_launch"], x["train_frac_of_measured_hbm_peak"]), x.get("trace"))
except Exception as e:
    print("$v ERR", e); print(open("gpurun_out/r1m_insert_$v.err").read()[-1500:])
PY
done
