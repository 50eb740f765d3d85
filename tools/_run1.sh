set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "window or tups_against or size_independent or synthetic" 2>&1 | tail -15
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bw1.json 2> gpurun_out/bw1.err; tail -3 gpurun_out/bw1.err; cat gpurun_out/bw1.json
