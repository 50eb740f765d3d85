#!/bin/bash
# Round-2 tenth visit (2 GPUs): sharded energy + theta gradient through the re-sharding phases; CAS(18,18) timing; bench N=2.
out=gpurun_out; mkdir -p $out; tag=r2j
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -k "2" > $out/${tag}_dist.log 2>&1
echo "dist rc=$?"; tail -30 $out/${tag}_dist.log | cut -c1-300
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $B --master-port 29731 tools/sharded_check.py 18 1 sigma grad > $out/${tag}_cas18.log 2>&1
echo "cas18 rc=$?"; tail -4 $out/${tag}_cas18.log
timeout 300 $B --master-port 29732 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_bench2.log 2>&1
echo "bench rc=$?"; tail -1 $out/${tag}_bench2.log | cut -c1-300
