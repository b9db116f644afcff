#!/bin/bash
# A/B of libqdx variants with the same ABI: tools/run_ab.sh <tag> <config> <variant...>
tag=$1; cfg=$2; shift 2
for v in "$@"; do
  lib=$PWD/qdax_b200/libqdx_$v.so; [ "$v" = base ] && lib=$PWD/qdax_b200/libqdx.so
  QDX_LIB_PATH=$lib python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_${cfg}_$v.json 2> gpurun_out/${tag}_${cfg}_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_${cfg}_$v.json')); print('$cfg','$v', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], {k:round(x,4) for k,x in d['kernel_ms'].items()}, d['final'])" || tail -3 gpurun_out/${tag}_${cfg}_$v.err
done
