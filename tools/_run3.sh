set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r1_win.json 2> gpurun_out/bench_r1_win.err; tail -2 gpurun_out/bench_r1_win.err; cat gpurun_out/bench_r1_win.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1_win.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_r1_win.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:win_kernel -s 30 -c 2 -o gpurun_out/prof_r1_win python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_r1_win.log 2>&1
tail -2 gpurun_out/prof_r1_win.log
