#!/bin/bash
# Round-2 visit 25 (8 GPUs): final code -- parity worlds 8 (state) and 4 (sigma + gradient routes incl. the backwards sweep),
# CAS(20,20) energy + theta gradient without the adjoint pass, bench --gpus 8.
out=gpurun_out; mkdir -p $out; tag=r2y
timeout 500 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -k "(ups and 8) or (sigma and 4)" > $out/${tag}_dist.log 2>&1
echo "dist rc=$?"; tail -6 $out/${tag}_dist.log | cut -c1-300
B="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 420 $B --nproc-per-node 8 --master-port 29751 tools/sharded_check.py 20 2 grad > $out/${tag}_cas20.log 2>&1
echo "cas20 rc=$?"; tail -3 $out/${tag}_cas20.log | cut -c1-600
timeout 200 $B --nproc-per-node 8 --master-port 29752 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_bench8.log 2>&1
echo "bench8 rc=$?"; tail -1 $out/${tag}_bench8.log | cut -c1-600
