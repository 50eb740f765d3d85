#!/bin/bash
# Round-2 visit 31 (8 GPUs): half sigma build of spin-flip symmetric sharded vectors: parity world 4, CAS(20,20) energy + theta gradient.
out=gpurun_out; mkdir -p $out; tag=r3e
timeout 400 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -k "sigma and 4" > $out/${tag}_dist.log 2>&1
echo "dist rc=$?"; tail -4 $out/${tag}_dist.log | cut -c1-300
B="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 420 $B --nproc-per-node 8 --master-port 29781 tools/sharded_check.py 20 2 grad > $out/${tag}_cas20.log 2>&1
echo "cas20 rc=$?"; tail -3 $out/${tag}_cas20.log | cut -c1-600
