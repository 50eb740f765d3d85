#!/bin/bash
# Round-2 fourth visit (1 GPU): ncu --set full captures of win_kernel and the DMMA kernels (source-level counters).
out=gpurun_out; mkdir -p $out; tag=r2d
timeout 600 ncu --set full --clock-control none --import-source on -k regex:win_kernel -s 30 -c 2 -f -o $out/${tag}_win \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-extras --no-cpu-baseline > $out/${tag}_win.log 2>&1
echo "win rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sigma_dmma -s 4 -c 1 -f -o $out/${tag}_sigma \
   python tools/bench_energy.py 16 1 > $out/${tag}_sigma.log 2>&1
echo "sigma rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_dmma -s 4 -c 1 -f -o $out/${tag}_gram \
   python tools/bench_energy.py 16 1 > $out/${tag}_gram.log 2>&1
echo "gram rc=$?"
ls -la $out/*.ncu-rep
