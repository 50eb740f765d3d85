#!/bin/bash
# Round-2 visit 30 (2 GPUs): half sigma build of spin-flip symmetric SHARDED vectors: parity world 2, timing at CAS(16,16) / (18,18).
out=gpurun_out; mkdir -p $out; tag=r3d
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -k "sigma and 2" > $out/${tag}_dist.log 2>&1
echo "dist rc=$?"; tail -25 $out/${tag}_dist.log | cut -c1-250
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
{
for v in 1 0; do
  echo "== SQ_SPINSYM_SHARDED=$v"
  SQ_SPINSYM_SHARDED=$v timeout 400 $B --master-port 2977$v tools/sharded_check.py 18 1 grad 2>&1 | grep -E "energy \+ theta|Error|error" | tail -2
done
} > $out/${tag}_ab_spinsym_sharded.txt 2>&1
cat $out/${tag}_ab_spinsym_sharded.txt
