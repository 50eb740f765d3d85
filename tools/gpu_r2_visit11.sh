#!/bin/bash
# Round-2 eleventh visit (8 GPUs): sharded parity at world 8, strong-scaling bench at 8 and 4 GPUs, CAS(20,20) energy + gradient.
out=gpurun_out; mkdir -p $out; tag=r2k
timeout 400 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -k "ups and 8" > $out/${tag}_dist8.log 2>&1
echo "dist8 rc=$?"; tail -12 $out/${tag}_dist8.log | cut -c1-300
B="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $B --nproc-per-node 8 --master-port 29741 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench8.log 2>&1
echo "bench8 rc=$?"; tail -1 $out/${tag}_bench8.log | cut -c1-2500
timeout 200 $B --nproc-per-node 4 --master-port 29742 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_bench4.log 2>&1
echo "bench4 rc=$?"; tail -1 $out/${tag}_bench4.log | cut -c1-400
timeout 420 $B --nproc-per-node 8 --master-port 29743 tools/sharded_check.py 20 2 grad > $out/${tag}_cas20.log 2>&1
echo "cas20 rc=$?"; tail -4 $out/${tag}_cas20.log
