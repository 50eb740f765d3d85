#!/bin/bash
# Round-2 visit 18 (1 GPU): quad_grad_kernel with the rolled step loop (9 600 instead of 29 800 instructions): parity, A/B.
out=gpurun_out; mkdir -p $out; tag=r2r
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gradient" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -3 $out/${tag}_tests.log
{ timeout 600 python tools/ab_grad.py 16 16 "5:4:0,40,3,16,2"; echo "== SQ_QGRAD_MINB=1"; SQ_QGRAD_MINB=1 timeout 600 python tools/ab_grad.py 16 16 "5:4:0,40,3,16,2"; } > $out/${tag}_ab_quadgrad.txt 2>&1; cat $out/${tag}_ab_quadgrad.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quad_grad_kernel -s 4 -c 1 -f -o $out/${tag}_quadgrad python tools/ab_grad.py 16 1 "5:4:0,40,3,16,2" > $out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
