#!/bin/bash
# GPU-box visit: full parity suite, the bench line, and one ncu --set full capture of the fp64 GEMMs (DMMA evidence).
tag=${1:-r1c}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=6 --durations=10 > $out/${tag}_tests.log 2>&1
echo "tests rc=$?" | tee -a $out/${tag}_tests.log
tail -25 $out/${tag}_tests.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"; cat $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
timeout 600 ncu --set full --clock-control none -k regex:'Kernel2' -s 2 -c 4 -o $out/${tag}_prof_dgemm -f \
    python tools/bench_energy.py 16 1 > $out/${tag}_prof_dgemm.log 2>&1
echo "ncu dgemm rc=$?"
