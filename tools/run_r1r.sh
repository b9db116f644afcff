#!/bin/bash
# final round-1 numbers for the streaming commit kernel: timing, one ncu --set full capture, bench lines of every config, launch list
python tools/time_insert.py > gpurun_out/r1r_insert.json 2> gpurun_out/r1r_insert.err
cat gpurun_out/r1r_insert.json | python -c "import json,sys; d=json.load(sys.stdin); [print(k, 'single %.4f ms %.3f | train %.4f ms %.3f'%(v['commit_ms'], v['frac_of_measured_hbm_peak'], v['train_ms_per_launch'], v['train_frac_of_measured_hbm_peak']), v['winners']) for k,v in d.items()]" || tail -5 gpurun_out/r1r_insert.err
ncu --set full --clock-control none --import-source on -k regex:qdx_commit -s 2 -c 1 -o gpurun_out/r1r_prof_commit -f python tools/time_insert.py > gpurun_out/r1r_ncu.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r1r_n1.json 2> gpurun_out/r1r_n1.err
python bench.py --config c1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r1r_c1.json 2> gpurun_out/r1r_c1.err
python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1r_c2.json 2> gpurun_out/r1r_c2.err
python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1r_c4.json 2> gpurun_out/r1r_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1r_launches_c3.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r1r_ncu_c3.log 2>&1
python - <<'PY'
import json
for n in ["n1","c1","c2","c4"]:
    try:
        d=json.load(open(f"gpurun_out/r1r_{n}.json")); print(n, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], d["final"], d.get("cpu_baseline"))
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r1r_{n}.err").read()[-1200:])
PY
