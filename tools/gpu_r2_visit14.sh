#!/bin/bash
# Round-2 visit 14 (1 GPU): win3_kernel with uniform block code (orientation in the matrix variants) and tabulated copies.
out=gpurun_out; mkdir -p $out; tag=r2n
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window_sweeps or cas16 or tups or averaged" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -5 $out/${tag}_tests.log
{
for v in "SQ_WIN3=0" "SQ_WIN3=1" "SQ_WIN3=1 SQ_WIN3_THREADS=384" "SQ_WIN3=1 SQ_WIN3_RANGE=12" "SQ_WIN3=1 SQ_WIN3_TMAX=400"; do
  echo "== $v"; env $v timeout 300 python tools/win_scan.py --reps 5 1 2>&1 | tail -1
done
echo "== SQ_WIN3=1 per launch"; SQ_WIN3=1 SQ_LAUNCH_TIMING=1 timeout 300 python tools/win_scan.py --reps 1 1 2>&1 | tail -28
} > $out/${tag}_ab_win3.txt 2>&1
cat $out/${tag}_ab_win3.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:win3_kernel -s 30 -c 2 -f -o $out/${tag}_win3 \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-extras --no-cpu-baseline > $out/${tag}_win3.log 2>&1
echo "ncu rc=$?"
