#!/bin/bash
# GPU-box visit: compute-sanitizer over small-CAS parity tests (race / memory checks).
#   memcheck  -- out-of-bounds / misaligned accesses of every kernel the selected tests launch (window, brick, generic, gather,
#                panel, DMMA, symmetric Gram, batched, re-shard kernels)
#   racecheck -- shared-memory hazards of the kernels that stage tiles in shared memory (win_kernel, win_grad_kernel, DMMA pipelines,
#                the bulk-copy re-shard kernel)
# The sanitizer slows kernels down 10-100 x, so only tests on small CAS are selected.
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh r2san'
tag=${1:-r2san}
out=gpurun_out
mkdir -p $out
SEL_MEM='native_library or idx2det or synthetic_ups_states or propagate_state_generic or tups_against_oracle or generic_generators or rdms_against_reference or sigma_against_oracle or window_sweeps_equal_single_brick_launches[8 or window_gradient_sweep[8 or per_string or averaged or variants or reshard_rows'
SEL_RACE='tups_against_oracle[8 or sigma_against_oracle[6 or rdms_against_reference or window_sweeps_equal_single_brick_launches[8 or window_gradient_sweep[8 or averaged or reshard_rows'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL_MEM" > $out/${tag}_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a $out/${tag}_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $out/${tag}_memcheck.log | tail -6
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL_RACE" > $out/${tag}_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a $out/${tag}_racecheck.log
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|Race|hazard" $out/${tag}_racecheck.log | tail -8
