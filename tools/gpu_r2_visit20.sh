#!/bin/bash
# Round-2 visit 20 (1 GPU): quad_grad_kernel with value-major partial sums and one reduction launch; residency 2 / 3.
out=gpurun_out; mkdir -p $out; tag=r2t
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gradient" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -3 $out/${tag}_tests.log
{ timeout 600 python tools/ab_grad.py 16 16 "5:4:0,40,3,16,2" | head -3; echo "== SQ_QGRAD_MINB=3"; SQ_QGRAD_MINB=3 timeout 600 python tools/ab_grad.py 16 16 "5:4:0,40,3,16,2" | head -3; } > $out/${tag}_ab_quadgrad.txt 2>&1; cat $out/${tag}_ab_quadgrad.txt
