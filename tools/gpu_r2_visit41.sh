#!/bin/bash
# Round-2 visit 41 (2 GPUs): the multi-GPU parity worlds against the single-GPU engine after the blocked sigma panels became its default
out=gpurun_out; mkdir -p $out; tag=r3u
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -k "2" > $out/${tag}_dist.log 2>&1
echo "dist rc=$?"; tail -5 $out/${tag}_dist.log | cut -c1-300
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $B --master-port 29791 tests/dist_sigma_worker.py > $out/${tag}_worker.log 2>&1
echo "worker rc=$?"; grep -v "^\*\|OMP_NUM\|^$\|NCCL" $out/${tag}_worker.log | tail -12 | cut -c1-220
