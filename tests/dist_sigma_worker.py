"""torchrun worker: H|psi> of an alpha-sharded vector (sq_sigma_dist: NVLink peer gathers + system-scope atomics into the
owners' shards) against the single-GPU sigma kernel, and <psi|H|psi> through it against the RDM route.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/dist_sigma_worker.py
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
    from slowquant_b200.distributed import (
        ShardedSpace, energy_and_theta_gradient_sharded, energy_sharded, energy_sharded_sigma, sigma_sharded,
    )
    from slowquant_b200.util import UpsStructure
    from slowquant_b200.operators import hamiltonian_0i_0a

    worst = 0.0
    for n, na, nb, symmetric in [(6, 3, 3, False), (6, 3, 3, True), (8, 4, 4, True), (9, 4, 5, False), (10, 5, 5, True)]:
        sp = ShardedSpace(0, n, 0, na, nb, device=local_rank)
        info = get_indexing(0, n, 0, na, nb, device=local_rank)
        rng = np.random.default_rng(2000 + n)       # same stream on every rank
        full = rng.normal(size=info.num_det)
        full /= np.linalg.norm(full)
        h = rng.normal(size=(n, n))
        g = 0.1 * rng.normal(size=(n, n, n, n))
        if symmetric:                                # real-orbital symmetry; the other cases assume nothing about g
            h = h + h.T
            g = g + g.transpose(1, 0, 2, 3)
            g = g + g.transpose(0, 1, 3, 2)
            g = g + g.transpose(2, 3, 0, 1)
        e_core = 0.37
        ref = osa.propagate_state([hamiltonian_0i_0a(h, g, 0, n)], full, info) + e_core * full
        nbs = info.num_beta_strings
        lo, hi = sp.row_begin * nbs, sp.row_end * nbs
        st = sp.alloc_state()
        st.set_from_full(full)
        sig = sigma_sharded(st, h, g, e_core)
        scale = float(np.max(np.abs(ref)))
        err = float(np.max(np.abs(sig.local.cpu().numpy() - ref[lo:hi]))) / scale if hi > lo else 0.0
        sig.close()
        e_sig = energy_sharded_sigma(st, h, g, e_core)
        e_ref = float(full @ ref)
        e_rdm = energy_sharded(st, h, g, e_core)
        err_e = max(abs(e_sig - e_ref), abs(e_sig - e_rdm)) / max(1.0, abs(e_ref))
        st.close()
        # energy + theta gradient of U(theta)|HF> on the sharded vector (shift-rule composition) vs the fused single-GPU call
        err_g = 0.0
        if n <= 9:
            lay = UpsStructure()
            lay.create_tiled(n, {"n_layers": 1, "do_tups": True})
            th = rng.uniform(-np.pi, np.pi, lay.n_params)
            hf = np.zeros(info.num_det)
            hf[0] = 1.0
            e1, g1 = osa.ups_energy_and_gradient(hf, info, th, lay, hamiltonian_0i_0a(h, g, 0, n))
            ref_st = sp.alloc_state()
            ref_st.set_determinant(0)
            e2, g2 = energy_and_theta_gradient_sharded(ref_st, th, lay, h, g, 0.0)                       # re-sharding phases (default)
            e5, g5 = energy_and_theta_gradient_sharded(ref_st, th, lay, h, g, 0.0, reshard=False)        # layout A only: fused local stretches
            e3, g3 = energy_and_theta_gradient_sharded(ref_st, th, lay, h, g, 0.0, fused_local=False)    # shift rule everywhere
            e4, g4 = energy_and_theta_gradient_sharded(ref_st, th, lay, h, g, 0.0, peer_gradient=True)   # exchange bricks on peer memory
            ref_st.close()
            err_g = max(abs(e1 - e2), abs(e1 - e3), abs(e1 - e4), abs(e1 - e5), float(np.max(np.abs(g1 - g2))), float(np.max(np.abs(g1 - g3))),
                        float(np.max(np.abs(g1 - g4))), float(np.max(np.abs(g1 - g5)))) / max(1.0, abs(e1))
            if rank == 0:
                print(f"    theta gradient: re-sharded {np.max(np.abs(g1 - g2)):.2e}, fused-local {np.max(np.abs(g1 - g5)):.2e}, shift rule {np.max(np.abs(g1 - g3)):.2e}, "
                      f"peer kernel {np.max(np.abs(g1 - g4)):.2e}", flush=True)
        # a spin-flip symmetric vector (a tUPS state on the closed-shell reference; lambda = +1 / -1 for even / odd pair counts):
        # the half build of the sharded sigma (kept cyclic band + mirror pass) against the full build and the single-GPU sigma
        err_s = 0.0
        if na == nb and symmetric:
            import slowquant_b200.distributed as D

            lay_s = UpsStructure()
            lay_s.create_tiled(n, {"n_layers": 2, "do_tups": True})
            th_s = rng.uniform(-np.pi, np.pi, lay_s.n_params)
            hf = np.zeros(info.num_det)
            hf[0] = 1.0
            psi = osa.construct_ups_state(hf, info, th_s.tolist(), lay_s)
            osa._lib.load().sq_set_option(b"sigma_spinsym", b"0")           # the single-GPU reference: full build
            ref_s = osa.propagate_state([hamiltonian_0i_0a(h, g, 0, n)], psi, info) + e_core * psi
            osa._lib.load().sq_set_option(b"sigma_spinsym", b"1")
            st_s = sp.alloc_state()
            st_s.set_from_full(psi)
            outs, rdms = [], []
            for flag in (True, False):
                D._SPINSYM_SHARDED = flag
                sg = sigma_sharded(st_s, h, g, e_core)
                outs.append(sg.local.cpu().numpy().copy())
                sg.close()
                rdms.append(D.rdm12_sharded(st_s, st_s))      # S / A Gram route on the shards; half band when the flag is on
            D._SPINSYM_SHARDED = True
            st_s.close()
            r1, r2 = osa.reduced_density_matrices(psi, psi, info)
            err_r = max(float(np.max(np.abs(rdms[0][0] - r1))), float(np.max(np.abs(rdms[0][1] - r2))),
                        float(np.max(np.abs(rdms[1][0] - r1))), float(np.max(np.abs(rdms[1][1] - r2))))
            err_e = max(err_e, err_r)
            if rank == 0:
                print(f"    RDMs of the symmetric state on the shards (half band / full) vs single GPU: {err_r:.2e}", flush=True)
            sc = float(np.max(np.abs(ref_s)))
            if hi > lo:
                err_s = max(float(np.max(np.abs(outs[0] - ref_s[lo:hi]))), float(np.max(np.abs(outs[1] - ref_s[lo:hi]))),
                            float(np.max(np.abs(outs[0] - outs[1])))) / sc
            if rank == 0:
                print(f"    spin-flip symmetric state: half build vs single GPU {np.max(np.abs(outs[0] - ref_s[lo:hi])) / sc:.2e}, "
                      f"half vs full sharded build {np.max(np.abs(outs[0] - outs[1])) / sc:.2e}", flush=True)
        e = torch.tensor([max(err, err_s), max(err_e, err_g)], dtype=torch.float64, device="cuda")
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"CAS({na + nb},{n}) world={world} symmetric={symmetric}: sigma rel. max|diff| {e[0]:.2e}, energy / theta gradient {e[1]:.2e}", flush=True)
        worst = max(worst, float(e.max()))
    dist.barrier()
    dist.destroy_process_group()
    if worst > 1e-11:
        print("DIST_SIGMA_FAILED", worst, flush=True)
        sys.exit(1)
    if rank == 0:
        print("DIST_SIGMA_OK", flush=True)


if __name__ == "__main__":
    main()
