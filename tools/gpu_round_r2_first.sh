#!/bin/bash
# First GPU visit of the next round (2 GPUs): run what was written without GPU time at the end of round 1.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_round_r2_first.sh r2a'
# 1. the opt-in tests (sq_sigma_dist + the shift-rule theta gradient on sharded vectors), 2. the new per-string parity test,
# 3. the whole GPU suite.  When 1 is green, drop the SQ_RUN_UNVERIFIED gate in tests/test_gpu_distributed.py.
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
SQ_RUN_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q -k "sigma and 2" > $out/${tag}_unverified.log 2>&1
echo "unverified rc=$?"; tail -15 $out/${tag}_unverified.log
SQ_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rdm_triangle" > $out/${tag}_rdm_tri.log 2>&1
echo "rdm_tri rc=$?"; tail -5 $out/${tag}_rdm_tri.log
SQ_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "table_free" > $out/${tag}_etab_alu.log 2>&1
echo "etab alu rc=$?"; tail -5 $out/${tag}_etab_alu.log
timeout 300 python tools/ab_option.py 16 etab smem alu > $out/${tag}_ab_etab_alu.txt 2>&1; tail -12 $out/${tag}_ab_etab_alu.txt
timeout 300 python tools/ab_option.py 16 rdm_tri 0 1 > $out/${tag}_ab_rdm_tri.txt 2>&1; tail -12 $out/${tag}_ab_rdm_tri.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "per_string" > $out/${tag}_strings.log 2>&1
echo "per-string rc=$?"; tail -5 $out/${tag}_strings.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=6 --durations=10 > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -20 $out/${tag}_tests.log
