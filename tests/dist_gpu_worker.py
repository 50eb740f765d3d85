"""torchrun worker: alpha-sharded tUPS/QNP state construction on N GPUs checked against the single-GPU
engine (each rank recomputes the full vector on its own GPU and compares its shard; the single-GPU engine is itself
compared with the oracle and the reference's goldens in tests/test_gpu_parity.py).  Every route runs: the re-sharding
phases with the bulk-copy and the vector load/store re-shard kernel, and the peer-memory exchange route; plus a circuit with
generic operators between the bricks and an operator that is local in neither row layout.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/dist_gpu_worker.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    from slowquant_b200 import operator_state_algebra as osa
    from slowquant_b200.ci_spaces import get_indexing
    from slowquant_b200.distributed import ShardedSpace, construct_ups_state_sharded, dot_sharded, energy_sharded, rdm12_sharded
    from slowquant_b200.util import UpsStructure
    from slowquant_b200 import _lib

    def _sub(lay, k0, k1):
        part = UpsStructure()
        part.excitation_operator_type = list(lay.excitation_operator_type[k0:k1])
        part.excitation_indices = list(lay.excitation_indices[k0:k1])
        part.n_params = k1 - k0
        return part

    worst = 0.0
    cases = [(8, 4, 4, 2, False), (9, 4, 5, 2, True), (10, 5, 5, 2, False), (12, 6, 6, 1, False)]
    for n, na, nb, L, qnp in cases:
        sp = ShardedSpace(0, n, 0, na, nb, device=local_rank)
        info = get_indexing(0, n, 0, na, nb, device=local_rank)
        lay = UpsStructure()
        lay.create_tiled(n, {"n_layers": L, "do_qnp": True} if qnp else {"n_layers": L, "do_tups": True})
        rng = np.random.default_rng(1000 + n)       # same stream on every rank
        th = rng.uniform(-np.pi, np.pi, lay.n_params)
        full = rng.normal(size=info.num_det)
        full /= np.linalg.norm(full)
        ref = osa.construct_ups_state(full, info, th, lay)
        ref_d = osa.construct_ups_state(full, info, th, lay, dagger=True)
        nbs = info.num_beta_strings
        lo, hi = sp.row_begin * nbs, sp.row_end * nbs
        st = sp.alloc_state()
        err = 0.0
        lib = _lib.load()
        for route, kernel in ((True, b"tma"), (True, b"lsu"), (False, b"tma")):
            lib.sq_set_option(b"reshard", kernel)
            st.set_from_full(full)
            construct_ups_state_sharded(st, th, lay, reshard=route)
            torch.cuda.synchronize()
            if hi > lo:
                err = max(err, float(np.max(np.abs(st.local.cpu().numpy() - ref[lo:hi]))))
            # a sub-range of the circuit (propagate_unitary-style), forward and adjoint
            k0, k1 = 3, lay.n_params - 2
            for dg in (False, True):
                part = osa.construct_ups_state(full, info, th[k0:k1], _sub(lay, k0, k1), dagger=dg)
                st.set_from_full(full)
                construct_ups_state_sharded(st, th, lay, dagger=dg, first=k0, last=k1, reshard=route)
                torch.cuda.synchronize()
                if hi > lo:
                    err = max(err, float(np.max(np.abs(st.local.cpu().numpy() - part[lo:hi]))))
        lib.sq_set_option(b"reshard", b"tma")
        st.set_from_full(full)
        construct_ups_state_sharded(st, th, lay)
        nrm = dot_sharded(st, st)
        st.set_from_full(full)
        construct_ups_state_sharded(st, th, lay, dagger=True)
        torch.cuda.synchronize()
        err_d = float(np.max(np.abs(st.local.cpu().numpy() - ref_d[lo:hi]))) if hi > lo else 0.0
        # HF determinant start
        st.set_determinant(0)
        hf = np.zeros(info.num_det)
        hf[0] = 1.0
        ref_hf = osa.construct_ups_state(hf, info, th, lay)
        construct_ups_state_sharded(st, th, lay)
        torch.cuda.synchronize()
        err_hf = float(np.max(np.abs(st.local.cpu().numpy() - ref_hf[lo:hi]))) if hi > lo else 0.0
        # RDMs and energy of the sharded vector (and a transition RDM between two sharded vectors) vs the single-GPU engine
        d1_ref, d2_ref = osa.reduced_density_matrices(ref_hf, ref_hf, info)
        d1, d2 = rdm12_sharded(st, st)
        err_rdm = max(float(np.max(np.abs(d1 - d1_ref))), float(np.max(np.abs(d2 - d2_ref))))
        h_syn = rng.normal(size=(n, n))
        g_syn = 0.1 * rng.normal(size=(n, n, n, n))
        e_ref = float(np.sum(h_syn * d1_ref) + 0.5 * np.sum(g_syn * d2_ref))
        err_rdm = max(err_rdm, abs(energy_sharded(st, h_syn, g_syn) - e_ref))
        st2 = sp.alloc_state()
        st2.set_from_full(full)
        t1_ref, t2_ref = osa.reduced_density_matrices(full, ref_hf, info)
        t1, t2 = rdm12_sharded(st2, st)
        err_rdm = max(err_rdm, float(np.max(np.abs(t1 - t1_ref))), float(np.max(np.abs(t2 - t2_ref))))
        st2.close()
        st.close()
        e = torch.tensor([err, err_d, err_hf, abs(nrm - 1.0), err_rdm], dtype=torch.float64, device="cuda")
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"CAS({na + nb},{n}) L={L} world={world}: max|diff| fwd {e[0]:.2e} dagger {e[1]:.2e} hf {e[2]:.2e} "
                  f"|norm-1| {e[3]:.2e} rdm/energy {e[4]:.2e}", flush=True)
        worst = max(worst, float(e.max()))
    # generic operators between the bricks: spin-orbital single / double, an sa_double, and a pair double between the first and
    # the last orbital, which neither row layout can run locally ("X" phase: tiles rotated in place over peer memory)
    n, na, nb = 8, 4, 4
    sp = ShardedSpace(0, n, 0, na, nb, device=local_rank)
    info = get_indexing(0, n, 0, na, nb, device=local_rank)
    lay = UpsStructure()
    lay.create_tiled(n, {"n_layers": 2, "do_tups": True})
    kk = max(world - 1, 0).bit_length()
    free = [o for o in range(n) if o >= kk]                    # alpha moves on the first log2(world) orbitals would cross GPUs
    extra_t = ["single", "double", "sa_double_1", "double", "sa_single"]
    extra_i = [(2 * free[0], 2 * free[2]), (2 * free[0], 2 * free[1] + 1, 2 * free[2], 2 * free[3] + 1),
               (free[0], free[1], free[2], free[3]), (0, 1, 2 * (n - 1), 2 * (n - 1) + 1), (0, n - 1)]
    lay.excitation_operator_type[7:7] = extra_t
    lay.excitation_indices[7:7] = extra_i
    lay.n_params = len(lay.excitation_operator_type)
    rng = np.random.default_rng(77)
    th = rng.uniform(-np.pi, np.pi, lay.n_params)
    full = rng.normal(size=info.num_det)
    full /= np.linalg.norm(full)
    nbs = info.num_beta_strings
    lo, hi = sp.row_begin * nbs, sp.row_end * nbs
    st = sp.alloc_state()
    err = 0.0
    for dg in (False, True):
        ref = osa.construct_ups_state(full, info, th, lay, dagger=dg)
        for route in (True, False):
            st.set_from_full(full)
            construct_ups_state_sharded(st, th, lay, dagger=dg, reshard=route)
            torch.cuda.synchronize()
            if hi > lo:
                err = max(err, float(np.max(np.abs(st.local.cpu().numpy() - ref[lo:hi]))))
    st.close()
    e = torch.tensor([err], dtype=torch.float64, device="cuda")
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"CAS({na + nb},{n}) mixed circuit world={world}: max|diff| {e[0]:.2e}", flush=True)
    worst = max(worst, float(e.max()))
    dist.barrier()
    dist.destroy_process_group()
    if worst > 1e-12:
        print("DIST_CHECK_FAILED", worst, flush=True)
        sys.exit(1)
    if rank == 0:
        print("DIST_CHECK_OK", flush=True)


if __name__ == "__main__":
    main()
