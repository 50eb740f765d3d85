timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/win_scan.py "0" "1" "6:5:4,72,4,16,3" "6:5:4,72,3,16,3" "6:5:4,72,4,12,3" 2>&1 | tail -5
SQ_WIN="6:5:4,72,4,16,3" SQ_LAUNCH_TIMING=1 timeout 100 python tools/win_scan.py --layers 4 --reps 1 "6:5:4,72,4,16,3" 2>&1 | tail -12
