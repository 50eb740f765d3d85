nvidia-smi --query-gpu=index,memory.total --format=csv | head -9
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 tools/sharded_check.py 20 1 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -12 | tee gpurun_out/cas20_8gpu.log
