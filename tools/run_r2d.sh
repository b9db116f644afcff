#!/bin/bash
# round-1 final on one GPU: whole GPU test suite, the driver's default bench line (both arms), configs c1 / c2 / c4
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2d_tests.log; cat gpurun_out/r2d_tests.log
python bench.py > gpurun_out/r2d_n1.json 2> gpurun_out/r2d_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_ref.json 2> gpurun_out/r2d_ref.err
python bench.py --config c1 --steps 50 --warmup 5 --no-cpu-baseline --no-insert-probe > gpurun_out/r2d_c1.json 2> gpurun_out/r2d_c1.err
python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline --no-insert-probe > gpurun_out/r2d_c2.json 2> gpurun_out/r2d_c2.err
python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline --no-insert-probe > gpurun_out/r2d_c4.json 2> gpurun_out/r2d_c4.err
python - <<'PY'
import json
for n in ["n1","c1","c2","c4"]:
    try:
        d=json.load(open(f"gpurun_out/r2d_{n}.json")); print(n, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], "blocking %.4g"%d["e2e"]["blocking_readback_value"], "roofline %.3f"%d["roofline"]["frac"], (d["insert_roofline"].get("large_rows") or {}).get("frac"), d.get("cpu_baseline",{}).get("value"))
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r2d_{n}.err").read()[-1200:])
print(open("gpurun_out/r2d_ref.json").read()[:600])
PY
