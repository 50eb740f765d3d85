timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 tools/sharded_check.py 18 2 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -8
