#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r1g_tests.log
python tools/time_insert.py > gpurun_out/r1g_insert.json 2> gpurun_out/r1g_insert.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1g_n1.json 2> gpurun_out/r1g_n1.err
python bench.py --config c1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r1g_c1.json 2> gpurun_out/r1g_c1.err
python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1g_c2.json 2> gpurun_out/r1g_c2.err
python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_c4.json 2> gpurun_out/r1g_c4.err
cat gpurun_out/r1g_tests.log; cat gpurun_out/r1g_insert.json | python -c "import json,sys; d=json.load(sys.stdin); [print(k, round(v['commit_ms'],4), round(v['frac_of_measured_hbm_peak'],3), v['winners']) for k,v in d.items()]" || tail -5 gpurun_out/r1g_insert.err
python - <<'PY'
import json
for n in ["n1","c1","c2","c4"]:
    try:
        d=json.load(open(f"gpurun_out/r1g_{n}.json")); print(n, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], d["final"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r1g_{n}.err").read()[-1200:])
PY
