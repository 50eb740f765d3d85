#!/bin/bash
# One GPU-box visit: parity tests, the bench line, and the ncu evidence for the energy / RDM / gradient kernels.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r1b}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > $out/${tag}_tests.log 2>&1
echo "tests rc=$?" | tee -a $out/${tag}_tests.log
tail -15 $out/${tag}_tests.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"; cat $out/${tag}_bench.json
# every launch of the energy path at CAS(16,16): shares of D-panel build / DGEMM / scatter / gradient kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_energy_launches.csv \
    python tools/bench_energy.py 16 2 > $out/${tag}_energy_launches.log 2>&1
echo "ncu launches rc=$?"
# full captures: sigma/RDM panel kernels + the fp64 GEMM, and the fused gradient kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'build_D|scatter_E|gemm|GEMM' -c 6 -o $out/${tag}_prof_sigma -f \
    python tools/bench_energy.py 16 1 > $out/${tag}_prof_sigma.log 2>&1
echo "ncu sigma rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tile_grad' -s 4 -c 3 -o $out/${tag}_prof_grad -f \
    python tools/bench_energy.py 16 1 > $out/${tag}_prof_grad.log 2>&1
echo "ncu grad rc=$?"
ls -la $out | tail -12
