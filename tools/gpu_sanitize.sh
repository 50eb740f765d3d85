#!/bin/bash
# GPU-box visit: compute-sanitizer over the small-CAS parity tests (race / memory checks).
#   memcheck  -- out-of-bounds / misaligned accesses of every kernel the parity tests launch
#   racecheck -- shared-memory hazards of the kernels that stage tiles in shared memory (win_kernel, tile kernels)
# The sanitizer slows kernels down 10-100 x, so only tests on CAS <= (8,8) are selected; the first test of each
# run also proves the native library (not a fallback) is what runs.
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh r2a'
tag=${1:-r2san}
out=gpurun_out
mkdir -p $out
SEL='native_library or idx2det or synthetic_ups_states or wavefunction_states or propagate_state_generic or tups_against_oracle or generic_generators or rdms_against_reference or sigma_against_oracle or window_sweeps or window_gradient or per_string'
for tool in memcheck racecheck; do
  timeout 700 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > $out/${tag}_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a $out/${tag}_${tool}.log
  grep -E "ERROR SUMMARY|passed|failed|Race|Invalid" $out/${tag}_${tool}.log | tail -8
done
