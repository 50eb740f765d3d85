#!/bin/bash
# Round-2 first visit (2 GPUs): first runs of the code written without GPU time + sharded baseline numbers.
out=gpurun_out; mkdir -p $out; tag=r2a
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
SQ_RUN_UNVERIFIED=1 timeout 500 python -m pytest tests/test_gpu_distributed.py -m gpu -q -k "sigma and 2" > $out/${tag}_unverified.log 2>&1
echo "unverified rc=$?"; tail -25 $out/${tag}_unverified.log
SQ_RUN_UNVERIFIED=1 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rdm_triangle or table_free" > $out/${tag}_optin.log 2>&1
echo "optin rc=$?"; tail -8 $out/${tag}_optin.log
timeout 200 python tools/ab_option.py 16 etab smem alu > $out/${tag}_ab_etab_alu.txt 2>&1; tail -6 $out/${tag}_ab_etab_alu.txt
timeout 200 python tools/ab_option.py 16 rdm_tri 0 1 > $out/${tag}_ab_rdm_tri.txt 2>&1; tail -6 $out/${tag}_ab_rdm_tri.txt
timeout 600 python -m pytest tests -m gpu -q --maxfail=6 --durations=8 > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -20 $out/${tag}_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 --mode sharded --no-cpu-baseline --no-e2e --no-extras > $out/${tag}_bench_sharded2.log 2>&1
echo "bench sharded rc=$?"; tail -3 $out/${tag}_bench_sharded2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 tools/sharded_check.py 18 1 energy sigma > $out/${tag}_cas18.log 2>&1
echo "cas18 rc=$?"; tail -12 $out/${tag}_cas18.log
