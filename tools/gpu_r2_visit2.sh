#!/bin/bash
# Round-2 second visit (2 GPUs): first run of the re-sharding route (sq_reshard_rows, constrained layout-B spaces, phase driver).
out=gpurun_out; mkdir -p $out; tag=r2b
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -k "2" > $out/${tag}_dist.log 2>&1
echo "dist rc=$?"; tail -25 $out/${tag}_dist.log
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $B --master-port 29721 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $out/${tag}_bench2_reshard.log 2>&1
echo "bench reshard rc=$?"; tail -2 $out/${tag}_bench2_reshard.log
SQ_RESHARD_KERNEL=lsu timeout 300 $B --master-port 29722 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_bench2_reshard_lsu.log 2>&1
echo "bench reshard lsu rc=$?"; tail -1 $out/${tag}_bench2_reshard_lsu.log | cut -c1-1500
SQ_RESHARD=0 timeout 300 $B --master-port 29723 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_bench2_peer.log 2>&1
echo "bench peer rc=$?"; tail -1 $out/${tag}_bench2_peer.log | cut -c1-600
timeout 300 $B --master-port 29724 tools/sharded_check.py 18 2 > $out/${tag}_cas18.log 2>&1
echo "cas18 rc=$?"; tail -3 $out/${tag}_cas18.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cas16_bench_config" > $out/${tag}_cas16.log 2>&1
echo "cas16 test rc=$?"; tail -5 $out/${tag}_cas16.log
