ncu --set full --clock-control none --import-source on -k regex:win_kernel -s 2 -c 3 -o gpurun_out/prof_win4 python bench.py --layers 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_win4.log 2>&1
tail -3 gpurun_out/prof_win4.log
