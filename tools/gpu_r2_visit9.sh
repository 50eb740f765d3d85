#!/bin/bash
# Round-2 ninth visit (1 GPU): unified brick loop of win_kernel -- parity (window tests, CAS(16,16) test), bench.
out=gpurun_out; mkdir -p $out; tag=r2i
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window or cas16 or tups or synthetic or wavefunction_states or averaged" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -4 $out/${tag}_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; head -c 300 $out/${tag}_bench.json; echo
