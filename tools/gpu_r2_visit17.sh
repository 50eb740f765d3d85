#!/bin/bash
# Round-2 visit 17 (1 GPU): quad_grad_kernel (two commuting bricks per launch in the theta-gradient sweep): parity, A/B at CAS(16,16).
out=gpurun_out; mkdir -p $out; tag=r2q
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_saups.py -m gpu -x -q -k "gradient or wavefunction or rotosolve or config2 or saups or optimis" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -5 $out/${tag}_tests.log
timeout 600 python tools/ab_grad.py 16 16 "5:4:0,40,3,16,2" > $out/${tag}_ab_quadgrad.txt 2>&1; cat $out/${tag}_ab_quadgrad.txt
