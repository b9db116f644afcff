#!/bin/bash
# compute-sanitizer over the hot-path kernels at small shapes: memcheck (all), racecheck (shared-memory hazards of the
# streaming commit ring / generate tiles) and initcheck.  usage: tools/sanitize.sh <tag>      (needs a GPU; ~6 min)
#
# initcheck and the bulk-copy engine: rows written by cp.async.bulk shared->global (generate phase 3, streaming commit)
# are not recorded as initialised by the tool, so every later ordinary read of them is reported.  The second initcheck
# run below uses a library built with -DQDX_GEN_PLAIN_STORE=1 -DQDX_COMMIT_FORCE_GENERIC=1 (same kernels, same rows,
# ordinary 128-bit stores): if the reports disappear there, they were artefacts of the tool, not reads of garbage.
tag=${1:-san}
mkdir -p gpurun_out
SEL='(add_injected_random and 600-100-1040) or (add_injected_random and 9000-14400-512) or dns_golden_and_random or (cells_tensor_core_path and 1024-8-300) or (cells_tensor_core_path and 2000-16-1000) or reference_add_kat or add_golden_injected or (add_injected_random and 300-64-1-3) or (add_injected_random and 500-100-6-4) or golden_c1mini or mixing_emitter_fused_emit or add_with_extra_scores'
SMALL='ask_tell_generic_path or (mixing_emitter_fused_emit and 64-20-256) or reference_add_kat or golden_c1mini or (add_injected_random and 300-64-1-3) or (cells_tensor_core_path and 1024-8-300) or (dns_golden_and_random) or mels_reference_kat'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_mels.py tests/test_gpu_pytree.py -m gpu -x -q -k "$SEL or mels or pytree" > gpurun_out/${tag}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_mels.py -m gpu -x -q -k "(add_injected_random and 600-100-1040) or dns_golden_and_random or (cells_tensor_core_path and 1024-8-300) or reference_add_kat or (add_injected_random and 300-64-1-3) or golden_c1mini or mels_reference_kat" > gpurun_out/${tag}_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${tag}_racecheck.log | tail -3
for lib in default plain; do
  L=$PWD/qdax_b200/libqdx.so; [ $lib = plain ] && L=$PWD/qdax_b200/libqdx_plain.so
  [ -f $L ] || { echo "initcheck[$lib]: $L missing (tools/build_variant.sh plain -DQDX_GEN_PLAIN_STORE=1 -DQDX_COMMIT_FORCE_GENERIC=1)"; continue; }
  QDX_LIB_PATH=$L timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_mels.py -m gpu -q -k "$SMALL" > gpurun_out/${tag}_initcheck_$lib.log 2>&1
  echo "initcheck[$lib] rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_initcheck_$lib.log | tail -3
  grep -A3 "Uninitialized" gpurun_out/${tag}_initcheck_$lib.log | grep -oE "at [a-zA-Z_0-9]+|in [a-zA-Z_0-9<>, ]+\(" | sort | uniq -c | sort -rn | head -8
done
