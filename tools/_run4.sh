set -x
timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_shard_2_win.json 2> gpurun_out/bench_shard_2_win.err; tail -3 gpurun_out/bench_shard_2_win.err; cat gpurun_out/bench_shard_2_win.json
