#!/bin/bash
# Round-2 visit 35 (1 GPU): blocked half build with per-CTA action tables: parity, A/B at CAS(16,16), ncu --set full of the two kernels
out=gpurun_out; mkdir -p $out; tag=r3o
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spin_flip or sigma or rdm or config2 or fused_energy" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -5 $out/${tag}_tests.log | cut -c1-250
timeout 600 python tools/ab_sigma_spinsym.py > $out/${tag}_ab_spinsym.txt 2>&1; cat $out/${tag}_ab_spinsym.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"blk_kernel" -s 2 -c 3 -o $out/${tag}_blk -f python tools/ab_sigma_spinsym.py 14 > $out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
