#!/bin/bash
# Round-2 visit 12 (1 GPU): first run of win3_kernel (orbital-triple register blocks, merged tiles): parity, A/B against win_kernel.
out=gpurun_out; mkdir -p $out; tag=r2l
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window_sweeps or cas16 or tups or synthetic or wavefunction_states or averaged" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -5 $out/${tag}_tests.log
{
for v in "SQ_WIN3=0" "SQ_WIN3=1" "SQ_WIN3=1 SQ_WIN3_THREADS=384" "SQ_WIN3=1 SQ_WIN3_RANGE=3" "SQ_WIN3=1 SQ_WIN3_RANGE=12" "SQ_WIN3=1 SQ_WIN3_TMAX=400" "SQ_WIN3=1 SQ_WIN3_THREADS=384 SQ_WIN3_TMAX=800"; do
  echo "== $v"; env $v timeout 300 python tools/win_scan.py --reps 5 1 2>&1 | tail -2
done
} > $out/${tag}_ab_win3.txt 2>&1
cat $out/${tag}_ab_win3.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; head -c 300 $out/${tag}_bench.json; echo
