#!/bin/bash
# Round-2 visit 16 (1 GPU): full single-GPU suite after the backwards gradient sweep of sq_ups_energy_grad + bench with extras.
out=gpurun_out; mkdir -p $out; tag=r2p
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -5 $out/${tag}_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; head -c 600 $out/${tag}_bench.json; echo
