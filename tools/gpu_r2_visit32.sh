#!/bin/bash
# Round-2 visit 32 (2 GPUs): sharded RDMs through the S / A Gram route with peer gathers, half band for spin-flip symmetric vectors.
out=gpurun_out; mkdir -p $out; tag=r3f
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $B --master-port 29791 tests/dist_sigma_worker.py > $out/${tag}_worker.log 2>&1
echo "worker rc=$?"; grep -v "^\*\|OMP_NUM\|^$\|NCCL" $out/${tag}_worker.log | tail -22 | cut -c1-220
{
for v in 1 0; do
  echo "== SQ_SPINSYM_SHARDED=$v"
  SQ_SPINSYM_SHARDED=$v timeout 300 $B --master-port 2979$((2+v)) tools/sharded_check.py 16 2 rdm 2>&1 | grep -iE "rdm|energy|Error|error" | tail -3
done
} > $out/${tag}_ab_rdm_sharded.txt 2>&1
cat $out/${tag}_ab_rdm_sharded.txt | cut -c1-300
