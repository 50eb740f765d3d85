/*
 * sqsv.h — C ABI of the B200-native state-vector engine (libsqsv.so).
 *
 * This is the drop-in boundary for the state-vector hot path of SlowQuant
 * (reference: slowquant/unitary_coupled_cluster/{ci_spaces,operator_state_algebra,
 * density_matrix}.py and the rdm/energy/gradient methods of ups_wavefunction.py).
 * The reference has no FFI layer; its boundary is the Python function surface.
 * Every entry point below names the reference function it replaces (file:line,
 * relative to the reference root); the Python shim in slowquant_b200/ binds these
 * with ctypes and keeps the reference's call signatures.
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross this boundary;
 *   - pointers named *_dev are CUDA device pointers on the space's device,
 *     pointers named *_host are host pointers; the caller owns all of them;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returns an int status (SQ_OK == 0); nothing throws;
 *     sq_last_error() returns a thread-local message for the last failure;
 *   - the CI vector is the row-major matrix C[Ia][Ib] of fp64 amplitudes,
 *     Ia = alpha-string index, Ib = beta-string index, idx = Ia*Nb + Ib, which is
 *     exactly the determinant ordering of get_indexing (ci_spaces.py:76-116);
 *   - spin-orbital index = 2*spatial + (0 alpha | 1 beta), as in the reference.
 */
#ifndef SQSV_H
#define SQSV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQ_OK 0
#define SQ_ERR_INVALID 1     /* bad argument                                  -> ValueError */
#define SQ_ERR_CUDA 2        /* CUDA runtime failure                          -> RuntimeError */
#define SQ_ERR_OUTSIDE 3     /* operator string leaves the CI space           -> KeyError
                                (operator_state_algebra.py:131-135, do_unsafe=False) */
#define SQ_ERR_UNSUPPORTED 4 /* feature not available in this build           -> NotImplementedError */
#define SQ_ERR_NOMEM 5

/* excitation-operator type codes; mirror UpsStructure.excitation_operator_type
 * (util.py:642-1073) */
#define SQ_EXC_SA_SINGLE 0
#define SQ_EXC_SINGLE 1
#define SQ_EXC_DOUBLE 2
#define SQ_EXC_TRIPLE 3
#define SQ_EXC_QUADRUPLE 4
#define SQ_EXC_QUINTUPLE 5
#define SQ_EXC_SEXTUPLE 6
#define SQ_EXC_SA_DOUBLE_1 7
#define SQ_EXC_SA_DOUBLE_2 8
#define SQ_EXC_SA_DOUBLE_3 9
#define SQ_EXC_SA_DOUBLE_4 10
#define SQ_EXC_SA_DOUBLE_5 11

typedef struct sq_space sq_space;   /* CI space: string lists + device tables (CI_Info, ci_spaces.py:9-53) */
typedef struct sq_layout sq_layout; /* compiled ansatz layout (UpsStructure, util.py:642)                 */

const char* sq_last_error(void);
int sq_version(void);

/* ---- CI space (replaces get_indexing, ci_spaces.py:76-116) ------------------------------------ */

/* Build alpha/beta string lists in itertools.combinations order, rank tables and device copies.
 * [row_begin,row_end) is the range of alpha strings (rows of C) resident on this device; pass
 * 0,-1 for the whole vector.  device = CUDA ordinal; device = -1 builds a host-only space (integer
 * tables, export/rank functions; no kernels). */
int sq_space_create(int n_orb, int n_alpha, int n_beta, int device, int64_t row_begin,
                    int64_t row_end, sq_space** out);
/* The same with a CONSTRAINED alpha list: only the strings with (mask & alpha_cmask) == alpha_cpat (bit o = orbital o), in
 * the order they have in the full list; all of them are local rows.  This is the second row layout of a re-sharded vector
 * (rows grouped by the occupation of the last log2(G) orbitals, see sq_reshard_rows); operators that move an alpha electron
 * on a constrained orbital cannot run in such a space (SQ_ERR_UNSUPPORTED).  No reference counterpart (single-process). */
int sq_space_create_constrained(int n_orb, int n_alpha, int n_beta, int device, uint32_t alpha_cmask,
                                uint32_t alpha_cpat, sq_space** out);
int sq_space_destroy(sq_space* sp);
int64_t sq_space_num_det(const sq_space* sp);
int64_t sq_space_num_strings(const sq_space* sp, int spin /*0 alpha, 1 beta*/);
int64_t sq_space_local_rows(const sq_space* sp);
/* host copy of the occupation bitmask of every string (bit o = spatial orbital o) */
int sq_space_export_strings(const sq_space* sp, int spin, uint32_t* out_host);
/* idx2det[first .. first+count) as the reference's determinant integers
 * (interleaved a0 b0 a1 b1 ..., orbital 0 most significant; ci_spaces.py:99-107) */
int sq_space_export_idx2det(const sq_space* sp, int64_t first, int64_t count, int64_t* out_host);
/* det2idx for n determinants; -1 where the determinant is not in the space (ci_spaces.py:106) */
int sq_space_det2idx(const sq_space* sp, int64_t n, const int64_t* dets_host, int64_t* idx_host);

/* ---- ansatz layout (mirrors UpsStructure lists, util.py:645-651) ---------------------------- */

/* exc_type[k] is an SQ_EXC_* code; the indices of operator k are
 * idx_flat[idx_offsets[k] .. idx_offsets[k+1]) exactly as stored in
 * UpsStructure.excitation_indices (spatial (i,a)/(i,j,a,b) for sa_*, spin-orbital tuples otherwise). */
int sq_layout_create(sq_space* sp, int n_ops, const int32_t* exc_type, const int32_t* idx_offsets,
                     const int32_t* idx_flat, sq_layout** out);
/* sa_double_1..5 generators are sums of products of E_pq; the symbolic normal ordering stays on the
 * host (fermionic_operator.py), which attaches the normal-ordered strings of T_k = G - G^dagger here.
 * String s consists of ops_flat[op_offsets[s] .. op_offsets[s+1]), each entry = 2*spin_orbital + dagger,
 * in label order (creators first; fermionic_operator.py:27-104). */
int sq_layout_attach_generator(sq_layout* lay, int k, int n_strings, const int32_t* ops_flat,
                               const int32_t* op_offsets, const double* coeffs);
int sq_layout_destroy(sq_layout* lay);
int sq_layout_num_ops(const sq_layout* lay);
/* number of kernel launches sq_ups_apply would issue for ops [first,last) (all |theta| > 1e-28) */
int sq_layout_num_launches(const sq_layout* lay, int first, int last);
/* amplitudes those launches read and write; algorithmic HBM bytes = 16 * this (roofline accounting) */
int64_t sq_layout_touched_amplitudes(const sq_layout* lay, int first, int last);
/* launch plan of ops [first,last): out6 = {launches, window sweeps, bricks inside window sweeps, quad launches,
 * single-brick launches, other launches}.  (No reference counterpart: the reference applies one operator per pass.) */
int sq_layout_plan_stats(const sq_layout* lay, int first, int last, int64_t* out6);
/* the same for an explicit operator list (one phase of the re-sharding driver): out8 = the six counters above, then the number
 * of kernels (without gauge sweeps) and the amplitudes those kernels read and write. */
int sq_layout_plan_stats_list(const sq_layout* lay, int n_list, const int32_t* op_list, int64_t* out8);
/* the plan itself (planner tests): operators of [first,last) in execution order (dagger != 0: the reversed circuit) and the
 * launch each one rides in; thetas_host may be NULL (all active; |theta| < 1e-28 is skipped as in sq_ups_apply). */
int sq_layout_plan_export(const sq_layout* lay, const double* thetas_host, int first, int last, int dagger,
                          int32_t* ops_out, int32_t* launch_out, int cap, int* n_out);
/* run-time switches (A/B comparisons, tests).  name "wingrad": "1" routes sq_ups_grad_sweep through the window kernel
 * (default "0": one brick per launch).  name "win": launch planner of sq_ups_apply, value "0" (window sweeps off),
 * "1" (defaults) or "w1:w2:w3,smem_kb,min_suffix,max_bricks,min_bricks".  name "etab": E_pq table of the sigma / RDM
 * panel kernels in shared memory ("smem", default), constant memory ("const"), or no table at all ("alu": the record of E_pq is
 * computed from (p,q) with uniform integer instructions).  name "pipeline": "1" (default) overlaps
 * the gather, DGEMM and scatter of neighbouring sigma / RDM panels on internal streams, "0" runs one panel at a time.  name "rows": "0" (default) determinant-per-thread gather /
 * scatter kernels, "1" row-per-CTA kernels with the row staged in shared memory (measured slower; "rows_cfg" =
 * "threads,chunks" sets their geometry).  name "panel": determinants per
 * panel for spaces that build their panels afterwards ("0": about 1 GiB per panel).  name "rdm_tri": "1" builds the symmetric
 * Gram matrix of sq_rdm12 with bra == ket from three half-size DGEMMs (3/4 of the flops), "0" (default) from one DGEMM.
 * name "sigma_spinsym": sq_sigma / sq_rdm12 of a spin-flip symmetric vector (c[B,A] = lambda (-1)^popc(A & B) c[A,B], measured on
 * every call) from the determinants above the diagonal: "1" (default; 32 x 32 blocks of determinants, beta partners through their
 * mirrors), "tri" (determinant-per-thread kernels of the half build), "0" (always the full build).  name "quadgrad": "1" (default)
 * two commuting bricks per launch of the theta-gradient sweep, "0" one.  name "win3": "1" window sweeps with register blocks over
 * orbital triples (measured slower; default "0").  name "rdm_sym": "1" (default) two symmetric Gram matrices of the S / A generators
 * for bra == ket.  name "sgemm_cta" / "sgemm_wm": CTAs per SM ("2") and row parts per CTA ("2") of the sigma DMMA kernel.
 * The switches are process-global (not per space): set them before the calls they affect, from one thread. */
int sq_set_option(const char* name, const char* value);

/* ---- unitary product state (construct_ups_state, operator_state_algebra.py:963-1412;
 *      propagate_unitary, :1867-2309) ---------------------------------------------------------- */

/* In place: state <- U_{last-1} ... U_{first} state  (dagger=0), or the adjoint product
 * (reversed order, theta -> -theta; :990-1001) when dagger=1.  thetas_host has one entry per layout
 * operator; operators with |theta| < 1e-28 are skipped (:998). */
int sq_ups_apply(sq_space* sp, sq_layout* lay, const double* thetas_host, int first, int last,
                 int dagger, double* state_dev, void* stream);
/* The same circuit on a batch of n_states vectors states_dev + s * state_stride (the *_SA twins: construct_ups_state_SA,
 * :1415-1864; propagate_unitary_SA, :2312-2754; the common tail of RotoSolve's shifted states, ups_wavefunction.py:1183-1187).
 * Window and gauge sweeps process the whole batch in one launch (tables and work lists staged once per CTA for all states). */
int sq_ups_apply_batch(sq_space* sp, sq_layout* lay, const double* thetas_host, int first, int last, int dagger,
                       double* states_dev, int n_states, int64_t state_stride, void* stream);
/* The same product over an explicit operator list: layout operators op_list[0..n_list) in the given EXECUTION order (dagger
 * != 0 only negates the angles).  The caller vouches that this order is equivalent to the circuit order, i.e. operators that
 * changed places commute (disjoint orbitals) -- the re-sharding driver runs the part of a circuit that is executable in the
 * current row layout this way.  gauge_flags bit 0: the vector is in the window kernel's sign-free gauge on entry, bit 1: it
 * is in that gauge on return; 0 = the reference's sign convention on both sides, as sq_ups_apply.  An empty list with
 * gauge_flags = 1 just takes the vector out of the gauge. */
int sq_ups_apply_list(sq_space* sp, sq_layout* lay, const double* thetas_host, int n_list, const int32_t* op_list,
                      int dagger, int gauge_flags, double* state_dev, void* stream);
/* out <- T_k in   (get_grad_action, :2757-2865; sa_single uses Ta + Tb) */
int sq_grad_action(sq_space* sp, sq_layout* lay, int k, const double* in_dev, double* out_dev,
                   void* stream);
/* Fused reverse sweep of ups_wavefunction.py:1114-1138 for operators [first,last):
 *   grad_host[k] = 2 <bra| T_k |ket>;  bra <- U_k bra;  ket <- U_k ket   (k ascending). */
int sq_ups_grad_sweep(sq_space* sp, sq_layout* lay, const double* thetas_host, int first, int last,
                      double* bra_dev, double* ket_dev, double* grad_host, void* stream);

/* The same sweep over an explicit operator list in execution order (one phase of the re-sharding driver: operators that changed
 * places commute, which leaves every <bra|T_k|ket> unchanged); grad_host[i] belongs to operator op_list[i]. */
int sq_ups_grad_sweep_list(sq_space* sp, sq_layout* lay, const double* thetas_host, int n_list, const int32_t* op_list,
                           double* bra_dev, double* ket_dev, double* grad_host, void* stream);
/* The same sweep run BACKWARDS through the circuit: op_list in the execution order of the adjoint circuit; for every operator
 * 2 <bra|T_k|ket> is taken first, then both vectors <- U_k^dagger.  Started from (H psi, psi) this is the gradient of
 * ups_wavefunction.py:1114-1138 without the adjoint pass (T_k commutes with its own rotation). */
int sq_ups_grad_sweep_list_rev(sq_space* sp, sq_layout* lay, const double* thetas_host, int n_list, const int32_t* op_list,
                               double* bra_dev, double* ket_dev, double* grad_host, void* stream);

/* Energy and theta gradient of a unitary product state in one call (_calc_energy_optimization /
 * _calc_gradient_optimization, ups_wavefunction.py:1019-1142): psi = U(theta) ref, *energy_host = <psi|H|psi> with H given by
 * the folded integrals of sq_sigma, grad_host[k] = dE/dtheta_k (skipped when grad_host is NULL) by the sweep of :1114-1138 run
 * backwards through the circuit from (H psi, psi): g_k = 2 <bra|T_k|ket>, then both vectors <- U_k^dagger (T_k commutes with
 * its own rotation, so these are the reference's numbers without the adjoint pass U^dagger H psi).
 * work_ket_dev / work_bra_dev: two vectors of the state's length, distinct from ref_dev; on return work_ket_dev holds
 * psi when grad_host is NULL and psi swept back to the reference state otherwise. */
int sq_ups_energy_grad(sq_space* sp, sq_layout* lay, const double* thetas_host, double e_core,
                       const double* h_act_host, const double* g_act_host, const double* ref_dev,
                       double* work_ket_dev, double* work_bra_dev, double* energy_host, double* grad_host, void* stream);

/* ---- alpha-sharded vectors: one process per GPU, shards peer-mapped over NVLink ----------------
 * (no counterpart in the reference, which is single-process).  The vector is split by rows
 * (alpha strings); operators whose row pairs stay on one device run unchanged, the others rotate their
 * tiles in place through peer pointers (the two owners of a pair split its columns).  The caller separates
 * "exchange" operators from their neighbours with a device-wide barrier (sq_layout_needs_exchange). */

/* contiguous row ranges grouped by the occupation of the first log2(world) orbitals; row_starts_out[world+1] */
int sq_partition_prefix(int n_orb, int n_alpha, int world, int64_t* row_starts_out);
/* declare the partition of a space created with [row_starts[rank], row_starts[rank+1]) */
int sq_space_set_partition(sq_space* sp, int world, int rank, const int64_t* row_starts);
/* plain cudaMalloc'ed shard (IPC-exportable), and CUDA IPC plumbing: 64-byte handles */
int sq_dist_alloc(int device, int64_t n_doubles, double** out);
int sq_dist_free(double* ptr);
int sq_ipc_export(const double* ptr, unsigned char* handle64);
int sq_ipc_import(int device, const unsigned char* handle64, double** out);
int sq_ipc_close(double* ptr);
/* work-list statistics of operator k on this rank: {orbital-pair id or -1, local row pairs, local inert
 * rows, cross-device row pairs this rank works on (half the columns each), cross_global, touched amplitudes} */
int sq_layout_op_stats(const sq_layout* lay, int k, int64_t* out6);
/* 1 if operators [first,last) contain a row pair that spans two devices (identical on every rank) */
int sq_layout_needs_exchange(const sq_layout* lay, int first, int last);
/* sq_ups_apply on a sharded vector: shard_ptrs_host[r] = device pointer of rank r's shard as mapped in THIS
 * process (own shard for r == rank).  Only orbital-pair operators (sa_single, pair double) may exchange. */
int sq_ups_apply_dist(sq_space* sp, sq_layout* lay, const double* thetas_host, int first, int last,
                      int dagger, double* const* shard_ptrs_host, void* stream);
/* Re-shard (the all-to-all exchange step of the sharded engine): the n_rows local rows of src_dev (row r =
 * NB doubles) move to their owners in the other row layout, row r to local row dst_row_dev[r] of rank dst_rank_dev[r],
 * written straight into dst_ptrs_host[rank] (that rank's destination buffer as mapped into this process) over NVLink by
 * the bulk-copy engine.  The tables are device arrays built by the caller from the string lists.  Device-wide barrier
 * after it (all rows landed) before any rank computes on the destination buffers. */
int sq_reshard_rows(int device, int64_t n_rows, int64_t NB, const double* src_dev, const int32_t* dst_rank_dev,
                    const int32_t* dst_row_dev, double* const* dst_ptrs_host, int world, void* stream);
/* 1 if operator k cannot run in this space (constrained alpha list, sq_space_create_constrained), else 0 */
int sq_layout_op_blocked(const sq_layout* lay, int k);
/* 1-/2-RDM contributions of THIS rank's rows of alpha-sharded vectors (same definitions as sq_rdm12; the caller sums the
 * ranks' results, e.g. with an all-reduce).  *_ptrs_host[r] = base pointer of rank r's shard as mapped into this process;
 * alpha partners on other ranks are read in place over NVLink.  All ranks must have finished writing their shards. */
int sq_rdm12_dist(sq_space* sp, const double* const* bra_ptrs_host, const double* const* ket_ptrs_host,
                  double* rdm1_host, double* rdm2_host, void* stream);
/* sq_rdm12_dist for a spin-flip symmetric vector (bra == ket; lambda = +-1 from sq_spinsym_measure_dist and a MAX all-reduce): the
 * panels hold the kept half of this rank's rows only, off-diagonal columns weighted by sqrt(2).  lambda = 0: sq_rdm12_dist. */
int sq_rdm12_dist_sym(sq_space* sp, const double* const* bra_ptrs_host, const double* const* ket_ptrs_host, double lambda,
                      double* rdm1_host, double* rdm2_host, void* stream);

/* H|in> of an alpha-sharded vector (the string path of energy_elec, ups_wavefunction.py:770-784, and the sigma vector
 * behind the theta gradient, :1091-1112), accumulated: every rank treats the determinants of ITS rows as sources, reads
 * their alpha partners in place over NVLink and adds the images that belong to another rank's rows into that rank's
 * out shard with system-scope atomics through the peer mapping.  Protocol: every rank sets out = e_core * in on its
 * shard, barrier, every rank calls sq_sigma_dist, stream synchronise, barrier.  *_ptrs_host[r] = base pointer of rank
 * r's shard as mapped into this process; in and out must not alias.  No symmetry of g is assumed. */
int sq_sigma_dist(sq_space* sp, const double* h_act_host, const double* g_act_host, const double* const* in_ptrs_host,
                  double* const* out_ptrs_host, void* stream);
/* Spin-flip symmetric sharded vectors (c[B,A] = lambda (-1)^popc(A & B) c[A,B]; every tUPS state on a closed-shell reference):
 * sq_spinsym_measure_dist -> this rank's {max|c|, max|c[B,A] - phi c[A,B]|, max|c[B,A] + phi c[A,B]|} over its rows (take the MAX over
 * the ranks; lambda = +1 / -1 if the second / third is <= 1e-12 of the first);  sq_sigma_dist_sym with lambda = +-1 builds
 * out += (H - e_core) in on the kept half of every rank's rows only (*used_half = 1; 0: the full build ran, e.g. integrals without
 * the real-orbital symmetry);  after a device-wide barrier sq_spinsym_mirror_dist writes the other half from its mirrors, and
 * a second barrier ends the build.  Same energies as sq_sigma_dist (reference ups_wavefunction.py:770-784) at half the work. */
int sq_spinsym_measure_dist(sq_space* sp, const double* const* in_ptrs_host, double* res3_host, void* stream);
int sq_sigma_dist_sym(sq_space* sp, const double* h_act_host, const double* g_act_host, const double* const* in_ptrs_host,
                      double* const* out_ptrs_host, double lambda, int* used_half, void* stream);
int sq_spinsym_mirror_dist(sq_space* sp, double* const* out_ptrs_host, double lambda, void* stream);

/* The theta-gradient loop (ups_wavefunction.py:1114-1138) over operators [first,last) of an alpha-sharded (bra, ket) pair
 * whose row pairs may live on two GPUs (sa_single / pair-double operators): g_k = 2 <bra|T_k|ket>, then both vectors <- U_k,
 * one fused launch per brick with cross-device tiles rotated in place through the peer mappings.  grad_host receives THIS
 * rank's partial sums (add the ranks' results); device-wide barrier before and after. */
int sq_ups_grad_sweep_dist(sq_space* sp, sq_layout* lay, const double* thetas_host, int first, int last,
                           double* const* bra_ptrs_host, double* const* ket_ptrs_host, double* grad_host, void* stream);

/* ---- generic operator application (apply_operator_serial/threaded, :53-219; propagate_state
 *      inner loop, :596-628) -------------------------------------------------------------------- */

/* out (+)= sum_s coeffs[s] * string_s |in>.  Strings as in sq_layout_attach_generator.  in_dev and
 * out_dev must not alias.  skip_outside != 0 reproduces do_unsafe=True (strings that leave the space
 * are skipped); otherwise such a string returns SQ_ERR_OUTSIDE. */
int sq_apply_strings(sq_space* sp, int n_strings, const int32_t* ops_flat, const int32_t* op_offsets,
                     const double* coeffs, const double* in_dev, double* out_dev, int accumulate,
                     int skip_outside, void* stream);

/* ---- BLAS-1 on CI vectors -------------------------------------------------------------------- */
int sq_dot(sq_space* sp, const double* a_dev, const double* b_dev, double* out_host, void* stream);
int sq_axpy(sq_space* sp, double alpha, const double* x_dev, double* y_dev, void* stream);
int sq_scale_copy(sq_space* sp, double alpha, const double* x_dev, double* y_dev, void* stream);

/* ---- Hamiltonian and reduced density matrices ------------------------------------------------ */

/* sigma <- H|in> with the folded active-space Hamiltonian
 *   H = e_core + sum_pq h_act[p][q] E_pq + 1/2 sum_pqrs g_act[p][q][r][s] e_pqrs
 * (hamiltonian_0i_0a folded, operators.py:476-529 + fermionic_operator.py:379-471);
 * h_act_host [n][n], g_act_host [n][n][n][n] in chemists' notation. */
int sq_sigma(sq_space* sp, double e_core, const double* h_act_host, const double* g_act_host,
             const double* in_dev, double* out_dev, void* stream);
/* rdm1[p][q] = <bra|E_pq|ket>, rdm2[p][q][r][s] = <bra|E_pq E_rs|ket> - delta_qr rdm1[p][s]
 * (ups_wavefunction.py:409-476); rdm2_host may be NULL. */
int sq_rdm12(sq_space* sp, const double* bra_dev, const double* ket_dev, double* rdm1_host,
             double* rdm2_host, void* stream);

/* ---- introspection (host only; works on a space created with device = -1) -------------------- */
/* closed-form action of one ladder string on the determinant with occupation masks (A,B):
 * valid = passes the screens of operator_state_algebra.py:118-123, (tgtA,tgtB) = flipped masks,
 * sign = phase of :127-134. */
int sq_debug_string_action(const sq_space* sp, const int32_t* ops, int n_ops, uint32_t A, uint32_t B,
                           int* valid, uint32_t* tgtA, uint32_t* tgtB, int* sign);
/* number of (p, q, spin) for which the table-free record of E_pq (sq_set_option("etab", "alu")) differs from the table the
 * sigma / RDM panel kernels use (built from the closed-form string action above); must be 0. */
int sq_debug_etab_closed_form(const sq_space* sp, int* n_mismatch);
/* TEST INFRASTRUCTURE, not a compute path: executes the launch plan of operators [first,last) on a HOST vector, every window
 * sweep through a host emulation of win3_kernel that shares its tables, step grouping and block algebra.  Lets the CPU tests
 * check that host logic against the oracle on spaces created with device = -1.  SQ_ERR_UNSUPPORTED if the plan holds anything
 * but window sweeps.  No function of slowquant_b200/ calls it. */
int sq_debug_win3_emulate(sq_space* sp, sq_layout* layout, const double* thetas_host, int first, int last, int dagger,
                          double* host_state);

/* ---- instrumentation ------------------------------------------------------------------------- */
/* number of kernels this library has launched since load (all spaces) */
int64_t sq_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SQSV_H */
