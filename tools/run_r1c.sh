#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1c_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1c_n1.json 2> gpurun_out/r1c_n1.err
QDX_LIB_PATH=$PWD/qdax_b200/libqdx_serial.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1c_n1_serial.json 2> gpurun_out/r1c_n1_serial.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1c_n1b.json 2> gpurun_out/r1c_n1b.err
python bench.py --config c1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r1c_c1.json 2> gpurun_out/r1c_c1.err
python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1c_c2.json 2> gpurun_out/r1c_c2.err
python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_c4.json 2> gpurun_out/r1c_c4.err
cat gpurun_out/r1c_tests.log
python - <<'PY'
import json
for n in ["n1","n1_serial","n1b","c1","c2","c4"]:
    try:
        d=json.load(open(f"gpurun_out/r1c_{n}.json")); print(n, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], "e2e_ms %.4f"%d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r1c_{n}.err").read()[-1500:])
PY
