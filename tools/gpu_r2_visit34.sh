#!/bin/bash
# Round-2 visit 34 (1 GPU): ncu --set full of the blocked half-build kernels (one launch each) at CAS(14,14)
out=gpurun_out; mkdir -p $out; tag=r3m
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"blk_kernel" -s 2 -c 4 -o $out/${tag}_blk -f python tools/ab_sigma_spinsym.py 14 > $out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $out/${tag}_ncu.log
