#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1d_tests.log
python bench.py > gpurun_out/r1d_n1.json 2> gpurun_out/r1d_n1.err
python bench.py --config c1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r1d_c1.json 2> gpurun_out/r1d_c1.err
python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1d_c2.json 2> gpurun_out/r1d_c2.err
python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_c4.json 2> gpurun_out/r1d_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1d_launches_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_ncu_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1d_launches_c2.csv python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qdx_generate -s 4 -c 1 -o gpurun_out/r1d_prof_generate -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_ncu_full.log 2>&1
cat gpurun_out/r1d_tests.log
python - <<'PY'
import json
for n in ["n1","c1","c2","c4"]:
    try:
        d=json.load(open(f"gpurun_out/r1d_{n}.json")); print(n, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], "e2e_ms %.4f"%d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"]["sm_mhz"], d.get("cpu_baseline",{}).get("value"))
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r1d_{n}.err").read()[-1500:])
PY
