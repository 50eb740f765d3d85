#!/bin/bash
# Round-2 visit 36 (1 GPU): win_kernel with the brick matrices staged in shared memory (128-bit broadcast loads instead of indexed
# constant loads): parity of the window tests, bench line without the extras
out=gpurun_out; mkdir -p $out; tag=r3p
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window or tups or synthetic or wavefunction_states or size_independent or state_averaged or cas16 or config3" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -4 $out/${tag}_tests.log | cut -c1-250
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"; cut -c1-400 $out/${tag}_bench.json
SQ_LAUNCH_TIMING=1 timeout 300 python tools/win_scan.py 2>&1 | tail -32 > $out/${tag}_timing.txt; tail -30 $out/${tag}_timing.txt
