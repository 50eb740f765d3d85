#!/bin/bash
# Round-2 visit 47 (1 GPU): WaveFunctionUPS optimisation callables on the shared sigma + backwards sweep: wave-function tests, new entry test
out=gpurun_out; mkdir -p $out; tag=r4c
timeout 900 python -m pytest tests -m gpu -x -q -k "backward or wavefunction or optimisation or rotosolve or tups_energy or linear_response or fused_energy or state_averaged or ucc" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -4 $out/${tag}_tests.log | cut -c1-250
