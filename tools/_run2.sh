timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "window or tups_against or size_independent or synthetic" 2>&1 | tail -3
timeout 200 python tools/win_scan.py "0" "1" "6:5:4,72,3,12,3" 2>&1 | tail -3
SQ_LAUNCH_TIMING=1 timeout 100 python tools/win_scan.py --layers 4 --reps 1 "1" 2>&1 | tail -10
timeout 300 python tools/bench_energy.py 16 2 2>&1 | tail -8
