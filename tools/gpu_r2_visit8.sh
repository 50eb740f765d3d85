#!/bin/bash
# Round-2 eighth visit (1 GPU): rewritten window gradient kernel -- parity, A/B against one brick per launch.
out=gpurun_out; mkdir -p $out; tag=r2h
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gradient" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -6 $out/${tag}_tests.log
timeout 600 python tools/ab_grad.py 16 16 > $out/${tag}_ab_grad.txt 2>&1; tail -14 $out/${tag}_ab_grad.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:win_grad_kernel -s 6 -c 1 -f -o $out/${tag}_wingrad python tools/ab_grad.py 16 2 "5:4:0,40,3,16,2" > /dev/null 2>&1
ls -la $out/${tag}_wingrad.ncu-rep
