#!/bin/bash
# Round-2 visit 26 (1 GPU): mixed-circuit test of the backwards gradient sweep.
out=gpurun_out; mkdir -p $out; tag=r2z
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "backwards or fused_energy or gradient" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -15 $out/${tag}_tests.log | cut -c1-200
