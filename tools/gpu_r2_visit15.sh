#!/bin/bash
# Round-2 visit 15 (1 GPU): launch-planner configurations (windows with short suffix runs cost 0.3-0.4 ms more per sweep).
out=gpurun_out; mkdir -p $out; tag=r2o
timeout 900 python tools/win_scan.py --reps 5 "1" "6:5:4,72,4,16,3" "6:5:4,72,5,16,3" "6:5:4,72,6,16,3" "6:5:0,72,5,16,3" "6:0:0,72,5,16,3" "6:5:4,72,5,12,3" "6:5:4,72,4,12,3" > $out/${tag}_scan.txt 2>&1
cat $out/${tag}_scan.txt
SQ_WIN="6:5:4,72,5,16,3" SQ_LAUNCH_TIMING=1 timeout 300 python tools/win_scan.py --reps 1 "6:5:4,72,5,16,3" 2>&1 | tail -32 > $out/${tag}_timing_suffix5.txt
cat $out/${tag}_timing_suffix5.txt
