for v in base rot u2 rotu2; do
  QDX_LIB_PATH=$PWD/qdax_b200/libqdx_$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print('$v', d['value'], d['ms_per_step'], d['kernel_ms'])" || tail -3 gpurun_out/ab_$v.err
done
QDX_LIB_PATH=$PWD/qdax_b200/libqdx_rotu2.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
