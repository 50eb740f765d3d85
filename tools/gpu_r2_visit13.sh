#!/bin/bash
# Round-2 visit 13 (1 GPU): why win3_kernel is not faster -- per-launch timing of both kernels, ncu --set full with source counters.
out=gpurun_out; mkdir -p $out; tag=r2m
{
for v in "SQ_WIN3=0" "SQ_WIN3=1"; do
  echo "== $v"; env $v SQ_LAUNCH_TIMING=1 timeout 300 python tools/win_scan.py --reps 1 1 2>&1 | tail -28
done
for v in "SQ_WIN3=1 SQ_WIN3_TMAX=800" "SQ_WIN3=1 SQ_WIN3_TMAX=640" "SQ_WIN3=1 SQ_WIN3_TMAX=300"; do
  echo "== $v"; env $v timeout 300 python tools/win_scan.py --reps 5 1 2>&1 | tail -1
done
} > $out/${tag}_timing.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:win3_kernel -s 30 -c 2 -f -o $out/${tag}_win3 \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-extras --no-cpu-baseline > $out/${tag}_win3.log 2>&1
echo "ncu rc=$?"
ls -la $out/${tag}*
