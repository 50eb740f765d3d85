#!/bin/bash
# Round-2 sixth visit (1 GPU): fused sigma kernel, batched window kernel -- parity suite, A/B timings, bench line.
out=gpurun_out; mkdir -p $out; tag=r2f
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -12 $out/${tag}_tests.log
timeout 300 python tools/ab_option.py 16 sigma_fused 0 1 > $out/${tag}_ab_sigma_fused.txt 2>&1; tail -10 $out/${tag}_ab_sigma_fused.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 3000 $out/${tag}_bench.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sigma_fused -s 1 -c 1 -f -o $out/${tag}_sigma_fused python tools/ab_option.py 16 sigma_fused 1 1 > /dev/null 2>&1
ls -la $out/${tag}_sigma_fused.ncu-rep
