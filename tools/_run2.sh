python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "window or tups_against" 2>&1 | tail -3
SQ_LAUNCH_TIMING=1 python tools/win_scan.py --layers 4 --reps 1 "1" 2>&1 | tail -7
python tools/win_scan.py "1" "8:6:4,16,12,100,6,8,3" "8:0:0,16,12,100,6,16,3" "8:0:0,16,12,100,6,10,3"  "8:0:0,16,12,100,6,7,3" 2>&1 | tail -5
