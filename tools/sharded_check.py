"""torchrun tool: size-independent checks of the alpha-sharded engine at a CAS no reference reaches
(norm conservation, U^dagger U = 1 round trip back to the HF determinant) plus timing.

    python -m torch.distributed.run --nproc-per-node 8 tools/sharded_check.py 20 2
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    from slowquant_b200.distributed import ShardedSpace, construct_ups_state_sharded, dot_sharded
    from slowquant_b200.util import UpsStructure

    ne = n // 2
    t0 = time.perf_counter()
    sp = ShardedSpace(0, n, 0, ne, ne, device=local_rank)
    lay = UpsStructure()
    lay.create_tiled(n, {"n_layers": L, "do_tups": True})
    th = np.random.default_rng(1234).uniform(-np.pi, np.pi, lay.n_params)
    st = sp.alloc_state()
    st.set_determinant(0)
    construct_ups_state_sharded(st, th, lay)   # includes table build
    torch.cuda.synchronize()
    dist.barrier()
    t_setup = time.perf_counter() - t0
    norm = dot_sharded(st, st) ** 0.5
    # timed: L more layers on the now dense vector
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    construct_ups_state_sharded(st, th, lay)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    norm2 = dot_sharded(st, st) ** 0.5
    # undo both applications: must return to the HF determinant
    construct_ups_state_sharded(st, th, lay, dagger=True)
    construct_ups_state_sharded(st, th, lay, dagger=True)
    torch.cuda.synchronize()
    loc = st.local
    if rank == 0 and loc.numel() > 0:
        first = float(loc[0])
        loc[0] -= 1.0
    err = float(torch.max(torch.abs(loc))) if loc.numel() else 0.0
    e = torch.tensor([err], dtype=torch.float64, device="cuda")
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    mem = torch.tensor([float(8 * sp.local_len) / 1e9], dtype=torch.float64, device="cuda")
    dist.all_reduce(mem, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(
            f"CAS({n},{n}) N_det={sp.ci_info.num_det} world={world} shard<={mem.item():.2f} GB  L={L}: "
            f"{dt*1e3:.1f} ms -> {L/dt:.2f} layers/s;  norm {norm:.15f} / {norm2:.15f};  "
            f"max|U^dU psi - psi| = {e.item():.2e};  setup+first apply {t_setup:.1f} s",
            flush=True,
        )
    st.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
