#!/bin/bash
# Round-2 visit 28 (1 GPU): sigma of spin-flip symmetric vectors from the upper triangle: parity, timing at CAS(16,16).
out=gpurun_out; mkdir -p $out; tag=r3b
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spin_flip or sigma or rdm or config2 or fused_energy or backwards or wavefunction" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -12 $out/${tag}_tests.log | cut -c1-250
timeout 600 python tools/ab_sigma_spinsym.py > $out/${tag}_ab_spinsym.txt 2>&1; cat $out/${tag}_ab_spinsym.txt
