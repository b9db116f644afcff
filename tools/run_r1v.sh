#!/bin/bash
timeout 600 python -m pytest tests/test_kmeans.py -m gpu -x -q 2>&1 | tail -4
bash tools/sanitize.sh r1v
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1v_n1.json 2> gpurun_out/r1v_n1.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r1v_n1.json")); print("n1", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "probe", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["insert_roofline"]["large_rows"].items() if k in ("launch_ms","frac","train_of_8_launch_ms","train_of_8_frac")})
except Exception as e:
    print("n1 ERR", e); print(open("gpurun_out/r1v_n1.err").read()[-1500:])
PY
