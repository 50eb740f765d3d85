#!/bin/bash
# Round-2 visit 22 (2 GPUs): sharded energy + theta gradient with the BACKWARDS sweep (no adjoint pass): parity worlds, A/B at CAS(16,16) and (18,18).
out=gpurun_out; mkdir -p $out; tag=r2v
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -k "2" > $out/${tag}_dist.log 2>&1
echo "dist rc=$?"; tail -5 $out/${tag}_dist.log | cut -c1-300
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
{
for v in 1 0; do
  echo "== SQ_BACKWARD_SWEEP=$v"
  SQ_BACKWARD_SWEEP=$v timeout 300 $B --master-port 2974$v tools/sharded_check.py 16 4 grad 2>&1 | grep -E "energy \+ theta|Error|error" | tail -2
  SQ_BACKWARD_SWEEP=$v timeout 400 $B --master-port 2975$v tools/sharded_check.py 18 1 grad 2>&1 | grep -E "energy \+ theta|Error|error" | tail -2
done
} > $out/${tag}_ab_backward.txt 2>&1
cat $out/${tag}_ab_backward.txt
timeout 300 $B --master-port 29762 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_bench2.log 2>&1
echo "bench rc=$?"; tail -1 $out/${tag}_bench2.log | cut -c1-400
