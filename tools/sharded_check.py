"""torchrun tool: size-independent checks of the alpha-sharded engine at a CAS no reference reaches
(norm conservation, U^dagger U = 1 round trip back to the HF determinant) plus timing.

    python -m torch.distributed.run --nproc-per-node 8 tools/sharded_check.py 20 2
    python -m torch.distributed.run --nproc-per-node 2 tools/sharded_check.py 18 2 energy

With a third argument "energy": also <H> of the sharded vector from its RDMs (synthetic symmetric integrals), checked at
the Hartree-Fock determinant against the closed form 2 sum_i h_ii + sum_ij (2 g_iijj - g_ijji), with Tr Gamma1 = N_e and
sum_pq Gamma2[ppqq] = N_e (N_e - 1) on the correlated state.  A fourth argument "sigma" adds <H> through the sharded sigma
vector (sq_sigma_dist), which must agree with the RDM route.
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
    from slowquant_b200.distributed import ShardedSpace, construct_ups_state_sharded, dot_sharded, rdm12_sharded
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
    tokens = set(sys.argv[3:])
    if "energy" in tokens:
        tokens |= {"rdm", "hf"}
    if "energy-nohf" in tokens:
        tokens |= {"rdm"}
    energy = None
    if tokens & {"rdm", "sigma", "grad"}:
        rng = np.random.default_rng(2024)
        A = rng.normal(size=(n, n))
        h = A + A.T
        B = 0.1 * rng.normal(size=(n, n, n, n))
        g = B + B.transpose(1, 0, 2, 3)
        g = g + g.transpose(0, 1, 3, 2)
        g = g + g.transpose(2, 3, 0, 1)
    if "rdm" in tokens:
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        d1, d2 = rdm12_sharded(st, st)
        t_rdm = time.perf_counter() - t0
        energy = float(np.sum(h * d1) + 0.5 * np.sum(g * d2))
        hf_msg = ""
        if "hf" in tokens:     # second vector + second RDM pass
            hf = sp.alloc_state()
            hf.set_determinant(0)
            h1, h2 = rdm12_sharded(hf, hf)
            hf.close()
            e_hf = float(np.sum(h * h1) + 0.5 * np.sum(g * h2))
            occ = range(ne)
            e_hf_exact = sum(2 * h[i, i] for i in occ) + sum(2 * g[i, i, j, j] - g[i, j, j, i] for i in occ for j in occ)
            hf_msg = f";  E_HF {e_hf:.12f} vs closed form {e_hf_exact:.12f} (diff {e_hf - e_hf_exact:.2e})"
        if rank == 0:
            sym = float(np.max(np.abs(d1 - d1.T)))
            print(
                f"CAS({n},{n}) world={world} sharded 1-/2-RDM: {t_rdm:.2f} s;  E = {energy:.12f};  Tr G1 = {np.trace(d1):.12f} "
                f"(N_e = {2 * ne});  sum G2[ppqq] = {np.einsum('ppqq->', d2):.10f} (N_e (N_e - 1) = {2 * ne * (2 * ne - 1)});  "
                f"max|G1 - G1^T| = {sym:.1e}" + hf_msg,
                flush=True,
            )
    if "sigma" in tokens:
        # <H> once more through H|psi> (sq_sigma_dist: peer gathers + NVLink atomics) -- must equal the RDM route
        from slowquant_b200.distributed import energy_sharded_sigma

        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        e_sig = energy_sharded_sigma(st, h, g)
        t_sig = time.perf_counter() - t0
        if rank == 0:
            cmp_msg = f" (E_sigma - E_RDM = {e_sig - energy:.2e})" if energy is not None else ""
            print(f"CAS({n},{n}) world={world} sharded sigma + dot: {t_sig:.2f} s;  E_sigma = {e_sig:.12f}" + cmp_msg, flush=True)
        if energy is None:
            energy = e_sig
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
    if "grad" in tokens:
        # energy + theta gradient of U(theta)|HF> (ups_wavefunction.py:1019-1142) on the sharded vector: state, sigma, adjoint and the
        # gradient loop as local phases between re-shards.  The HF reference is given by its determinant index (no extra vector).
        from slowquant_b200.distributed import energy_and_theta_gradient_sharded

        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        tm = {}
        e_g, grad = energy_and_theta_gradient_sharded(0, th, lay, h, g, space=sp, timings=tm)
        torch.cuda.synchronize()
        dist.barrier()
        t_g = time.perf_counter() - t0
        if rank == 0:
            parts = ", ".join(f"{k[:-2]} {v:.2f} s" for k, v in tm.items())
            print(f"CAS({n},{n}) world={world} energy + theta gradient of U|HF> ({lay.n_params} parameters): {t_g:.2f} s ({parts});  "
                  f"E = {e_g:.12f};  |grad| = {np.linalg.norm(grad):.10f}  max|grad| = {np.max(np.abs(grad)):.10f}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
