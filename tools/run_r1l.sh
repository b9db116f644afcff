#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1l_tests.log
python tools/time_insert.py > gpurun_out/r1l_insert.json 2> gpurun_out/r1l_insert.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1l_n1.json 2> gpurun_out/r1l_n1.err
for ex in p2p; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 --exchange $ex > gpurun_out/r1l_n2_$ex.json 2> gpurun_out/r1l_n2_$ex.err
done
cat gpurun_out/r1l_tests.log; cat gpurun_out/r1l_insert.json | python -c "import json,sys; d=json.load(sys.stdin); [print(k, round(v['commit_ms'],4), round(v['frac_of_measured_hbm_peak'],3), v['winners']) for k,v in d.items()]" || tail -5 gpurun_out/r1l_insert.err
python - <<'PY'
import json
for n in ["n1","n2_p2p"]:
    try:
        d=json.load(open(f"gpurun_out/r1l_{n}.json")); print(n, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "e2e %.4g"%d["e2e"]["value"], d["replicas_bit_identical"], d.get("exchange_used"), d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r1l_{n}.err").read()[-1200:])
PY
