#!/bin/bash
# Round-2 visit 46 (1 GPU): final state -- full single-GPU suite, smoke(), bench line with extras, reference arm, ncu launch list of the bench command.
out=gpurun_out; mkdir -p $out; tag=r4a
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -3 $out/${tag}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/${tag}_smoke.log
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; head -c 400 $out/${tag}_bench.json; echo
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "ref rc=$?"; head -c 400 $out/${tag}_bench_reference.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-extras --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1; echo "ncu rc=$?"
