#!/bin/bash
# compute-sanitizer over the kernels touched in this round at small shapes: memcheck (all), racecheck (shared-memory hazards of
# the streaming commit ring / generate tiles).  usage: tools/sanitize.sh <tag>
tag=${1:-san}
SEL='reference_add_kat or add_golden_injected or (add_injected_random and 300-64-1-3) or (add_injected_random and 500-100-6-4) or golden_c1mini or mixing_emitter_fused_emit or add_with_extra_scores'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_mels.py tests/test_gpu_pytree.py -m gpu -x -q -k "$SEL or mels or pytree" > gpurun_out/${tag}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_mels.py -m gpu -x -q -k "reference_add_kat or (add_injected_random and 300-64-1-3) or golden_c1mini or mels_reference_kat" > gpurun_out/${tag}_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${tag}_racecheck.log | tail -3
