// sigma-build H|c> and 1-/2-RDMs from the folded integrals, without ever materialising the
// 2n^2 + 2C(n,2)^2 + n^4 operator strings the reference loops over (operators.py:476-529 applied by
// operator_state_algebra.py:596-628; ups_wavefunction.py:409-476 for the RDMs).
//
// Both are built on panels of the single-replacement matrix
//     D[rs][J] = <J| E_rs |c>,   E_rs = a+_{r,alpha} a_{s,alpha} + a+_{r,beta} a_{s,beta}
// (Knowles-Handy resolution): for a panel of W determinants D is an n^2 x W dense fp64 matrix, so
//     <bra|E_pq E_rs|ket> = sum_J D^bra[qp][J] D^ket[rs][J]            (fp64 tensor-core GEMM, cuBLAS DGEMM)
//     sigma[I] = e_core c[I] + sum_pq sum_J <I|E_pq|J> ( k_pq c[J] + 1/2 sum_rs g_pqrs D[rs][J] )
// with k_pq = h_pq - 1/2 sum_r g_prrq.  These dense contractions are the only place this engine uses
// tensor cores (DMMA through the library GEMM); the gather that builds D and the scatter that applies
// E_pq are HBM/L2-bound integer-address kernels.

#include <cmath>
#include <cstring>

#include "sqsv_internal.h"

struct ERec {                 // one spin component of E_pq acting on a determinant
  uint32_t tocc, temp;        // target screen: p occupied, q empty unless p == q      (gather form)
  uint32_t occ, emp;          // source screen: q occupied, p empty unless p == q      (scatter form)
  uint32_t flip;              // bits p and q of the same-spin string (0 when p == q)
  uint32_t parS, parO;        // parity masks on the same-spin / other-spin SOURCE strings
  int32_t s0;                 // +-1
};

struct HamWork {
  double* d_gsym = nullptr;   // G_SS and G_AA (two 144 x 144 matrices) of the symmetric 2-RDM route
  double* d_gram = nullptr;   // split-K partial sums of the 2-RDM Gram matrix (sqsv_dmma.cu)
  size_t gram_doubles = 0;
  int n_sm = 0;
  ERec* d_etab = nullptr;     // [n*n][2]  (alpha, beta)
  std::vector<ERec> h_etab;
  uint32_t* d_tabG = nullptr; // [n*n][NB] beta gather table of the row kernels (build_D_rows_kernel)
  uint32_t* d_tabS = nullptr; // [n*n][NB] beta scatter table (scatter_E_rows_kernel)
  // per-string partner tables of the table-driven panel kernels (sq_set_option("etab", "tab")): entry = partner string index << 1 |
  // sign bit of the same-spin factor (s0 folded in), -1 where E_pq does not act; gather (target) form and scatter (source) form
  int32_t *d_pgA = nullptr, *d_pgB = nullptr, *d_psA = nullptr, *d_psB = nullptr;   // alpha: [NA][n*n], beta: [n*n][NB]
  uint32_t* d_parO = nullptr;  // [2][n*n] other-spin parity masks: alpha operators (applied to the beta string), beta operators
  unsigned long long* d_symres = nullptr;   // 3 words of spinsym_check_kernel
  double* d_D[4] = {nullptr, nullptr, nullptr, nullptr};   // ket panels (two in flight), bra panels (two in flight)
  double* d_F[2] = {nullptr, nullptr};
  // panel pipeline: gather, GEMM and scatter of neighbouring panels overlap on three internal streams
  cudaStream_t s_build = nullptr, s_gemm = nullptr, s_scat = nullptr;
  cudaEvent_t ev_start = nullptr, ev_built[2] = {nullptr, nullptr}, ev_gemm[2] = {nullptr, nullptr}, ev_scat[2] = {nullptr, nullptr};
  int64_t W = 0;
  double* d_small = nullptr;  // n^4 + 2 n^2 doubles: G2 accumulator / integral matrices
  int* d_frow = nullptr;      // [n^2] row of F for every (p,q)
};

static std::map<const sq_space*, HamWork*> g_work;

// The E_pq table is read uniformly by every thread (same slot at the same time).  Default: a shared-memory copy staged
// per CTA (four broadcast LDS.128 per slot and thread).  A per-device constant-memory copy (uniform LDC loads) is kept
// behind sq_set_option("etab", "const") as the measured alternative: at CAS(16,16) it is SLOWER for the sigma build
// (557 ms against 521 ms; RDMs unchanged), i.e. the LSU pressure of the panel kernels (profiles/
// r1_energy_kernels_ncu_summary.csv: data pipe 62-77 %) comes from the scattered beta gathers, not from the table reads.
// One constant table per device at a time: it is rebound when another space runs (after draining the device).
#define SQ_ETAB_CONST_ORBS 24
__constant__ ERec c_etab[2 * SQ_ETAB_CONST_ORBS * SQ_ETAB_CONST_ORBS];
static std::map<int, const sq_space*> g_etab_owner;   // device -> space whose table sits in c_etab
static int g_etab_const = 0;                          // sq_set_option("etab", "const") selects the constant table

void sq_hamiltonian_set_etab_mode(int use_const) { g_etab_const = use_const ? 1 : 0; }
// sq_set_option("etab", "alu"): no table at all.  The (p,q) loops of the panel kernels are uniform across the CTA, so the record
// of E_pq -- two single-bit masks, two interval masks, a sign -- is a handful of uniform-datapath integer instructions instead of
// four 16-byte shared-memory loads per (p,q) and warp (the panel kernels are LSU-bound: 62-77 % LSU data pipe,
// profiles/r1_energy_kernels_ncu_summary.csv).  Closed form of sq_make_string_action for the label [a+_p a_q] of one spin (checked
// entry by entry against the table on the host: sq_debug_etab_closed_form, tests/test_host_logic.py).  Candidate for the next GPU visit,
// off by default until measured.
static int g_etab_alu = 0;
void sq_hamiltonian_set_etab_alu(int on) { g_etab_alu = on ? 1 : 0; }
// sq_set_option("etab", "tab"): per-STRING partner tables instead of per-(p,q) records: what the record-driven kernels recompute for
// every determinant and (p,q) -- two mask tests, the flipped string, a popcount and a rank look-up per spin -- depends on one string
// only (12 870 alpha strings serve 165 M determinants at CAS(16,16)), so it is tabulated once per space (13 MB per table at n = 16);
// a thread then needs one (row-uniform) table load for the alpha partner, one coalesced load for the beta partner and one popcount
// for the other-spin sign.  Single-device spaces whose tables stay below 64 MB each; otherwise the record kernels run.
static int g_etab_tab = 0;
void sq_hamiltonian_set_etab_tab(int on) { g_etab_tab = on ? 1 : 0; }
__host__ __device__ __forceinline__ ERec erec_closed(int p, int q, int spin) {
  const uint32_t bp = 1u << p, bq = 1u << q;
  ERec r;
  r.tocc = bp;                  // target: p occupied ...
  r.temp = bq & ~bp;            // ... q empty unless p == q
  r.occ = bq;                   // source: q occupied ...
  r.emp = bp & ~bq;             // ... p empty unless p == q
  r.flip = bp ^ bq;
  r.parS = (bp - 1u) ^ (bq - 1u);                                       // same-spin orbitals in [min, max)
  r.parO = spin ? (((bp << 1) - 1u) ^ ((bq << 1) - 1u)) : r.parS;       // other spin: (min, max] for beta, [min, max) for alpha
  r.s0 = q < p ? -1 : 1;
  return r;
}

// sq_set_option("pipeline", "0"): one panel at a time on the caller's stream (the pre-pipeline behaviour, for A/B runs)
static int g_panel_pipeline = 1;
void sq_hamiltonian_set_pipeline(int on) { g_panel_pipeline = on ? 1 : 0; }
// sq_set_option("rdm_tri", ...): kept as an accepted no-op -- the hand-written Gram kernel always computes only the upper-triangular
// tiles when bra == ket (round 1 measured the three-half-block cuBLAS variant of this at 631 ms against 729 ms at CAS(16,16)).
void sq_hamiltonian_set_rdm_tri(int) {}
// sq_set_option("panel", "<determinants>"): panel width of spaces that have not built their panels yet (tests use it to get
// several panels at small CAS); "0" restores the 1 GiB default
static int64_t g_panel_width = 0;
void sq_hamiltonian_set_panel_width(long long w) { g_panel_width = w > 0 ? (int64_t)w : 0; }

static void free_work(HamWork* w) {
  if (!w) return;
  cudaFree(w->d_gram);
  cudaFree(w->d_gsym);
  cudaFree(w->d_etab);
  cudaFree(w->d_tabG);
  cudaFree(w->d_tabS);
  cudaFree(w->d_pgA); cudaFree(w->d_pgB); cudaFree(w->d_psA); cudaFree(w->d_psB); cudaFree(w->d_parO);
  cudaFree(w->d_symres);
  for (double* p : w->d_D) cudaFree(p);
  for (double* p : w->d_F) cudaFree(p);
  if (w->s_build) cudaStreamDestroy(w->s_build);
  if (w->s_gemm) cudaStreamDestroy(w->s_gemm);
  if (w->s_scat) cudaStreamDestroy(w->s_scat);
  if (w->ev_start) cudaEventDestroy(w->ev_start);
  for (int b = 0; b < 2; ++b) {
    if (w->ev_built[b]) cudaEventDestroy(w->ev_built[b]);
    if (w->ev_gemm[b]) cudaEventDestroy(w->ev_gemm[b]);
    if (w->ev_scat[b]) cudaEventDestroy(w->ev_scat[b]);
  }
  cudaFree(w->d_small);
  cudaFree(w->d_frow);
  delete w;
}

void sq_hamiltonian_release(const sq_space* sp) {
  for (auto& kv : g_etab_owner)
    if (kv.second == sp) kv.second = nullptr;
  auto it = g_work.find(sp);
  if (it != g_work.end()) {
    free_work(it->second);
    g_work.erase(it);
  }
}

static int g_rows_kernels = 0;   // sq_set_option("rows", "1"): row-per-CTA panel kernels (see "row kernels" below)
static bool rows_kernels_fit(const sq_space* sp);
static bool partner_tables_fit(const sq_space* sp);
static int build_partner_tables(sq_space* sp, HamWork* w);
static int build_beta_tables(sq_space* sp, HamWork* w);

static int get_work(sq_space* sp, bool need_second_D, bool need_F, HamWork** out) {
  HamWork* w = nullptr;
  auto it = g_work.find(sp);
  if (it != g_work.end()) {
    w = it->second;
  } else {
    w = new HamWork();
    g_work[sp] = w;
  }
  const int n = sp->n_orb, n2 = n * n;
  if (!w->n_sm) {
    SQ_CUDA(cudaDeviceGetAttribute(&w->n_sm, cudaDevAttrMultiProcessorCount, sp->device));
    if (w->n_sm < 1) w->n_sm = 1;
  }
  if (!w->d_etab) {
    std::vector<ERec> tab(2 * (size_t)n2);
    for (int p = 0; p < n; ++p)
      for (int q = 0; q < n; ++q)
        for (int spin = 0; spin < 2; ++spin) {
          int32_t label[2] = {2 * (2 * p + spin) + 1, 2 * (2 * q + spin)};
          StringAction a;
          SQ_CHECK(sq_make_string_action(sp, label, 2, &a));
          ERec r;
          if (spin == 0)
            r = {a.toccA, a.tempA, a.occA, a.empA, a.flipA, a.parA, a.parB, a.s0};
          else
            r = {a.toccB, a.tempB, a.occB, a.empB, a.flipB, a.parB, a.parA, a.s0};
          tab[2 * ((size_t)p * n + q) + spin] = r;
        }
    SQ_CUDA(cudaMalloc(&w->d_etab, sizeof(ERec) * tab.size()));
    SQ_CUDA(cudaMemcpy(w->d_etab, tab.data(), sizeof(ERec) * tab.size(), cudaMemcpyHostToDevice));
    w->h_etab = tab;
  }
  if (g_rows_kernels && rows_kernels_fit(sp) && !w->d_tabG) SQ_CHECK(build_beta_tables(sp, w));
  if (g_etab_tab && partner_tables_fit(sp) && !w->d_pgA) SQ_CHECK(build_partner_tables(sp, w));
  if (!w->W) {
    // panel width: about 1 GiB per n^2 x W matrix, multiple of 256 determinants
    int64_t Wmax = g_panel_width > 0 ? g_panel_width : ((int64_t)1 << 27) / n2;
    Wmax = (Wmax / 256) * 256;
    if (Wmax < 256) Wmax = 256;
    int64_t len = sp->local_len();
    // blocked half build (build_blk_kernel: one CTA per 1024 determinants, four CTAs per SM): whole waves of CTAs per panel
    const int64_t unit = (int64_t)w->n_sm * 4;
    if (g_panel_width <= 0 && sp->n_alpha == sp->n_beta && sp->NA == sp->NB && len >= 8 * Wmax && Wmax / 1024 >= unit / 2)
      Wmax = std::max<int64_t>(1, (Wmax / 1024 + unit / 2) / unit) * unit * 1024;
    w->W = len < Wmax ? ((len + 255) / 256) * 256 : Wmax;
    if (w->W < 256) w->W = 256;
  }
  const size_t pbytes = sizeof(double) * (size_t)n2 * (size_t)w->W;
  const int64_t n_panels = (sp->local_len() + w->W - 1) / w->W;
  const int depth = (g_panel_pipeline && n_panels > 1) ? 2 : 1;   // panels in flight
  for (int b = 0; b < depth; ++b) {
    if (!w->d_D[b]) SQ_CUDA(cudaMalloc(&w->d_D[b], pbytes));
    if (need_second_D && !w->d_D[2 + b]) SQ_CUDA(cudaMalloc(&w->d_D[2 + b], pbytes));
    if (need_F && !w->d_F[b]) SQ_CUDA(cudaMalloc(&w->d_F[b], pbytes));
  }
  if (!w->s_build) {
    SQ_CUDA(cudaStreamCreateWithFlags(&w->s_build, cudaStreamNonBlocking));
    SQ_CUDA(cudaStreamCreateWithFlags(&w->s_gemm, cudaStreamNonBlocking));
    SQ_CUDA(cudaStreamCreateWithFlags(&w->s_scat, cudaStreamNonBlocking));
    SQ_CUDA(cudaEventCreateWithFlags(&w->ev_start, cudaEventDisableTiming));
    for (int b = 0; b < 2; ++b) {
      SQ_CUDA(cudaEventCreateWithFlags(&w->ev_built[b], cudaEventDisableTiming));
      SQ_CUDA(cudaEventCreateWithFlags(&w->ev_gemm[b], cudaEventDisableTiming));
      SQ_CUDA(cudaEventCreateWithFlags(&w->ev_scat[b], cudaEventDisableTiming));
    }
  }
  if (!w->d_small) SQ_CUDA(cudaMalloc(&w->d_small, sizeof(double) * ((size_t)n2 * (n2 + 1) + 2 * (size_t)n2)));   // integral matrix with an even leading dimension + k
  if (!w->d_frow) SQ_CUDA(cudaMalloc(&w->d_frow, sizeof(int) * (size_t)n2));
  *out = w;
  return SQ_OK;
}

// D[slot][t] = <J_t| E_rs |in>, slot = r*n + s, J_t = determinant j0 + t of the local vector (gather form)
// table access: CONST = the per-device constant copy (uniform loads), otherwise a shared-memory copy staged per CTA
template <bool CONST>
__device__ __forceinline__ const ERec* stage_etab(const ERec* __restrict__ etab, int n2) {
  if (CONST) return c_etab;
  extern __shared__ ERec sm[];
  for (int w = threadIdx.x; w < 2 * n2 * (int)(sizeof(ERec) / 4); w += 256)
    reinterpret_cast<uint32_t*>(sm)[w] = reinterpret_cast<const uint32_t*>(etab)[w];
  __syncthreads();
  return sm;
}

template <bool CONST>
__global__ void __launch_bounds__(256)
build_D_kernel(const double* __restrict__ IN, double* __restrict__ D, int64_t W, int64_t j0, int64_t len,
               const ERec* __restrict__ etab, int n2, const uint32_t* __restrict__ strA,
               const uint32_t* __restrict__ strB, const int32_t* __restrict__ rankA,
               const int32_t* __restrict__ rankB, int64_t NB, int64_t row_begin) {
  const ERec* sm = stage_etab<CONST>(etab, n2);
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= W) return;
  const int64_t j = j0 + t;
  if (j >= len) {
    for (int slot = 0; slot < n2; ++slot) D[(int64_t)slot * W + t] = 0.0;
    return;
  }
  const int64_t ia_loc = j / NB, ib = j - ia_loc * NB;
  const uint32_t a = __ldg(strA + row_begin + ia_loc), b = __ldg(strB + ib);
  for (int slot = 0; slot < n2; ++slot) {
    double v = 0.0;
    const ERec ra = sm[2 * slot], rb = sm[2 * slot + 1];
    if ((a & ra.tocc) == ra.tocc && (a & ra.temp) == 0u) {
      const uint32_t sa = a ^ ra.flip;
      const int par = (__popc(sa & ra.parS) + __popc(b & ra.parO)) & 1;
      const double x = IN[((int64_t)__ldg(rankA + sa) - row_begin) * NB + ib];
      v += (par ? -ra.s0 : ra.s0) * x;
    }
    if ((b & rb.tocc) == rb.tocc && (b & rb.temp) == 0u) {
      const uint32_t sb = b ^ rb.flip;
      const int par = (__popc(sb & rb.parS) + __popc(a & rb.parO)) & 1;
      const double x = IN[ia_loc * NB + __ldg(rankB + sb)];
      v += (par ? -rb.s0 : rb.s0) * x;
    }
    D[(int64_t)slot * W + t] = v;
  }
}

// Alpha-sharded vector: the beta partner of a determinant lies in its own (local) row, the alpha partner in another row
// that may belong to another GPU; it is read in place through the peer-mapped shard of its owner (NVLink), the same
// mechanism the brick kernel uses for exchange operators.  rows [row_starts[r], row_starts[r+1]) live on rank r.
struct PeerView {
  const double* p[SQ_MAX_WORLD];
  int64_t row_starts[SQ_MAX_WORLD + 1];
  int world;
};
__global__ void __launch_bounds__(256)
build_D_peer_kernel(PeerView pv, const double* __restrict__ IN, double* __restrict__ D, int64_t W, int64_t j0, int64_t len,
                    const ERec* __restrict__ etab, int n2, const uint32_t* __restrict__ strA,
                    const uint32_t* __restrict__ strB, const int32_t* __restrict__ rankA,
                    const int32_t* __restrict__ rankB, int64_t NB, int64_t row_begin) {
  const ERec* sm = stage_etab<false>(etab, n2);
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= W) return;
  const int64_t j = j0 + t;
  if (j >= len) {
    for (int slot = 0; slot < n2; ++slot) D[(int64_t)slot * W + t] = 0.0;
    return;
  }
  const int64_t ia_loc = j / NB, ib = j - ia_loc * NB;
  const uint32_t a = __ldg(strA + row_begin + ia_loc), b = __ldg(strB + ib);
  for (int slot = 0; slot < n2; ++slot) {
    double v = 0.0;
    const ERec ra = sm[2 * slot], rb = sm[2 * slot + 1];
    if ((a & ra.tocc) == ra.tocc && (a & ra.temp) == 0u) {
      const uint32_t sa = a ^ ra.flip;
      const int par = (__popc(sa & ra.parS) + __popc(b & ra.parO)) & 1;
      const int64_t gr = __ldg(rankA + sa);          // global row of the alpha partner
      int o = 0;
      while (o + 1 < pv.world && gr >= pv.row_starts[o + 1]) ++o;
      const double x = pv.p[o][(gr - pv.row_starts[o]) * NB + ib];
      v += (par ? -ra.s0 : ra.s0) * x;
    }
    if ((b & rb.tocc) == rb.tocc && (b & rb.temp) == 0u) {
      const uint32_t sb = b ^ rb.flip;
      const int par = (__popc(sb & rb.parS) + __popc(a & rb.parO)) & 1;
      const double x = IN[ia_loc * NB + __ldg(rankB + sb)];
      v += (par ? -rb.s0 : rb.s0) * x;
    }
    D[(int64_t)slot * W + t] = v;
  }
}

// Symmetrised panel for integrals with g_pqrs = g_pqsr: Dsym[slot(r,s)][t] = <J_t| E_rs + E_sr |in> for r > s and
// <J_t| E_rr |in> for r = s, slot(r,s) = r (r + 1) / 2 + s  -- n (n + 1) / 2 rows instead of n^2.
template <bool CONST>
__global__ void __launch_bounds__(256)
build_Dsym_kernel(const double* __restrict__ IN, double* __restrict__ D, int64_t W, int64_t j0, int64_t len,
                  const ERec* __restrict__ etab, int n, const uint32_t* __restrict__ strA,
                  const uint32_t* __restrict__ strB, const int32_t* __restrict__ rankA,
                  const int32_t* __restrict__ rankB, int64_t NB, int64_t row_begin) {
  const int n2 = n * n;
  const ERec* sm = stage_etab<CONST>(etab, n2);
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= W) return;
  const int64_t j = j0 + t;
  const int nS = n * (n + 1) / 2;
  if (j >= len) {
    for (int slot = 0; slot < nS; ++slot) D[(int64_t)slot * W + t] = 0.0;
    return;
  }
  const int64_t ia_loc = j / NB, ib = j - ia_loc * NB;
  const uint32_t a = __ldg(strA + row_begin + ia_loc), b = __ldg(strB + ib);
  auto elem = [&](int slot) -> double {
    double v = 0.0;
    const ERec ra = sm[2 * slot], rb = sm[2 * slot + 1];
    if ((a & ra.tocc) == ra.tocc && (a & ra.temp) == 0u) {
      const uint32_t sa = a ^ ra.flip;
      const int par = (__popc(sa & ra.parS) + __popc(b & ra.parO)) & 1;
      v += (par ? -ra.s0 : ra.s0) * IN[((int64_t)__ldg(rankA + sa) - row_begin) * NB + ib];
    }
    if ((b & rb.tocc) == rb.tocc && (b & rb.temp) == 0u) {
      const uint32_t sb = b ^ rb.flip;
      const int par = (__popc(sb & rb.parS) + __popc(a & rb.parO)) & 1;
      v += (par ? -rb.s0 : rb.s0) * IN[ia_loc * NB + __ldg(rankB + sb)];
    }
    return v;
  };
  int slot = 0;
  for (int r = 0; r < n; ++r)
    for (int q = 0; q <= r; ++q, ++slot) D[(int64_t)slot * W + t] = (r == q) ? elem(r * n + r) : elem(r * n + q) + elem(q * n + r);
}

// OUT[E_pq J] += sign * ( F[pq][t] + k[pq] * IN[J] )   for every determinant J of the panel (scatter form)
template <bool CONST>
__global__ void __launch_bounds__(256)
scatter_E_kernel(const double* __restrict__ IN, double* __restrict__ OUT, const double* __restrict__ F,
                 const double* __restrict__ kmat, const int* __restrict__ frow, int64_t W, int64_t j0, int64_t len,
                 const ERec* __restrict__ etab, int n2, const uint32_t* __restrict__ strA,
                 const uint32_t* __restrict__ strB, const int32_t* __restrict__ rankA,
                 const int32_t* __restrict__ rankB, int64_t NB, int64_t row_begin) {
  const ERec* sm = stage_etab<CONST>(etab, n2);
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t j = j0 + t;
  if (t >= W || j >= len) return;
  const int64_t ia_loc = j / NB, ib = j - ia_loc * NB;
  const uint32_t a = __ldg(strA + row_begin + ia_loc), b = __ldg(strB + ib);
  const double cj = IN[j];
  double diag = 0.0;
  for (int slot = 0; slot < n2; ++slot) {
    const ERec ra = sm[2 * slot], rb = sm[2 * slot + 1];
    const bool va = (a & ra.occ) == ra.occ && (a & ra.emp) == 0u;
    const bool vb = (b & rb.occ) == rb.occ && (b & rb.emp) == 0u;
    if (!va && !vb) continue;
    const double val = F[(int64_t)__ldg(frow + slot) * W + t] + __ldg(kmat + slot) * cj;   // frow: row of F that holds (p,q)
    if (va) {
      const int par = (__popc(a & ra.parS) + __popc(b & ra.parO)) & 1;
      const double sv = (par ? -ra.s0 : ra.s0) * val;
      if (ra.flip == 0u) diag += sv;
      else atomicAdd(OUT + ((int64_t)__ldg(rankA + (a ^ ra.flip)) - row_begin) * NB + ib, sv);
    }
    if (vb) {
      const int par = (__popc(b & rb.parS) + __popc(a & rb.parO)) & 1;
      const double sv = (par ? -rb.s0 : rb.s0) * val;
      if (rb.flip == 0u) diag += sv;
      else atomicAdd(OUT + ia_loc * NB + __ldg(rankB + (b ^ rb.flip)), sv);
    }
  }
  atomicAdd(OUT + j, diag);
}

// ---- table-driven variants (sq_set_option("etab", "tab")): same arithmetic, per-string partner tables --------------------------------
__global__ void __launch_bounds__(256)
build_Dsym_tab_kernel(const double* __restrict__ IN, double* __restrict__ D, int64_t W, int64_t j0, int64_t len,
                      const int32_t* __restrict__ pgA, const int32_t* __restrict__ pgB, const uint32_t* __restrict__ parO, int n,
                      const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB, int64_t NB) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= W) return;
  const int64_t j = j0 + t;
  const int n2 = n * n, nS = n * (n + 1) / 2;
  if (j >= len) {
    for (int slot = 0; slot < nS; ++slot) D[(int64_t)slot * W + t] = 0.0;
    return;
  }
  const int64_t ia = j / NB, ib = j - ia * NB;
  const uint32_t a = __ldg(strA + ia), b = __ldg(strB + ib);
  const int32_t* ta = pgA + ia * n2;          // the same for every thread of the row: broadcast loads
  const double* rowp = IN + ia * NB;
  auto elem = [&](int pq) -> double {
    double v = 0.0;
    const int ea = __ldg(ta + pq);
    if (ea >= 0) {
      const double x = IN[(int64_t)(ea >> 1) * NB + ib];
      v += ((ea ^ __popc(b & __ldg(parO + pq))) & 1) ? -x : x;
    }
    const int eb = __ldg(pgB + (int64_t)pq * NB + ib);
    if (eb >= 0) {
      const double x = rowp[eb >> 1];
      v += ((eb ^ __popc(a & __ldg(parO + n2 + pq))) & 1) ? -x : x;
    }
    return v;
  };
  int slot = 0;
  for (int r = 0; r < n; ++r)
    for (int q = 0; q <= r; ++q, ++slot) D[(int64_t)slot * W + t] = (r == q) ? elem(r * n + r) : elem(r * n + q) + elem(q * n + r);
}

__global__ void __launch_bounds__(256)
scatter_E_tab_kernel(const double* __restrict__ IN, double* __restrict__ OUT, const double* __restrict__ F,
                     const double* __restrict__ kmat, const int* __restrict__ frow, int64_t W, int64_t j0, int64_t len,
                     const int32_t* __restrict__ psA, const int32_t* __restrict__ psB, const uint32_t* __restrict__ parO, int n,
                     const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB, int64_t NB) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t j = j0 + t;
  if (t >= W || j >= len) return;
  const int n2 = n * n;
  const int64_t ia = j / NB, ib = j - ia * NB;
  const uint32_t a = __ldg(strA + ia), b = __ldg(strB + ib);
  const int32_t* ta = psA + ia * n2;
  double* orow = OUT + ia * NB;
  const double cj = IN[j];
  double diag = 0.0;
  for (int p = 0; p < n; ++p)
    for (int q = 0; q < n; ++q) {
      const int pq = p * n + q;
      const int ea = __ldg(ta + pq), eb = __ldg(psB + (int64_t)pq * NB + ib);
      if (ea < 0 && eb < 0) continue;
      const double val = F[(int64_t)__ldg(frow + pq) * W + t] + __ldg(kmat + pq) * cj;
      if (ea >= 0) {
        const double sv = ((ea ^ __popc(b & __ldg(parO + pq))) & 1) ? -val : val;
        if (p == q) diag += sv;
        else atomicAdd(OUT + (int64_t)(ea >> 1) * NB + ib, sv);
      }
      if (eb >= 0) {
        const double sv = ((eb ^ __popc(a & __ldg(parO + n2 + pq))) & 1) ? -val : val;
        if (p == q) diag += sv;
        else atomicAdd(orow + (eb >> 1), sv);
      }
    }
  atomicAdd(OUT + j, diag);
}

// host: the partner tables from the E_pq records (w->h_etab) and the string lists of the space
static int build_partner_tables(sq_space* sp, HamWork* w) {
  const int n = sp->n_orb, n2 = n * n;
  const int64_t NA = sp->NA, NB = sp->NB;
  std::vector<int32_t> gA((size_t)NA * n2), sA((size_t)NA * n2), gB((size_t)n2 * NB), sB((size_t)n2 * NB);
  std::vector<uint32_t> parO(2 * (size_t)n2);
  auto entry = [](const std::vector<int32_t>& rank, uint32_t m, const ERec& r, bool gather) -> int32_t {
    const bool ok = gather ? ((m & r.tocc) == r.tocc && (m & r.temp) == 0u) : ((m & r.occ) == r.occ && (m & r.emp) == 0u);
    if (!ok) return -1;
    const uint32_t other = m ^ r.flip;
    const int par = (__builtin_popcount((gather ? other : m) & r.parS) & 1) ^ (r.s0 < 0 ? 1 : 0);
    const int32_t idx = rank[other];
    return idx < 0 ? -1 : (int32_t)((idx << 1) | par);
  };
  for (int pq = 0; pq < n2; ++pq) {
    const ERec &ra = w->h_etab[2 * pq], &rb = w->h_etab[2 * pq + 1];
    parO[pq] = ra.parO;
    parO[n2 + pq] = rb.parO;
    for (int64_t i = 0; i < NA; ++i) {
      gA[(size_t)i * n2 + pq] = entry(sp->rankA, sp->strA[i], ra, true);
      sA[(size_t)i * n2 + pq] = entry(sp->rankA, sp->strA[i], ra, false);
    }
    for (int64_t i = 0; i < NB; ++i) {
      gB[(size_t)pq * NB + i] = entry(sp->rankB, sp->strB[i], rb, true);
      sB[(size_t)pq * NB + i] = entry(sp->rankB, sp->strB[i], rb, false);
    }
  }
  auto up = [](int32_t** d, const std::vector<int32_t>& v) -> int {
    SQ_CUDA(cudaMalloc(d, sizeof(int32_t) * v.size()));
    SQ_CUDA(cudaMemcpy(*d, v.data(), sizeof(int32_t) * v.size(), cudaMemcpyHostToDevice));
    return SQ_OK;
  };
  SQ_CHECK(up(&w->d_pgA, gA));
  SQ_CHECK(up(&w->d_psA, sA));
  SQ_CHECK(up(&w->d_pgB, gB));
  SQ_CHECK(up(&w->d_psB, sB));
  SQ_CUDA(cudaMalloc(&w->d_parO, sizeof(uint32_t) * parO.size()));
  SQ_CUDA(cudaMemcpy(w->d_parO, parO.data(), sizeof(uint32_t) * parO.size(), cudaMemcpyHostToDevice));
  return SQ_OK;
}
static bool partner_tables_fit(const sq_space* sp) {
  const size_t n2 = (size_t)sp->n_orb * sp->n_orb;
  return sp->world <= 1 && sp->row_begin == 0 && sp->row_end == sp->NA && !sp->alpha_cmask &&
         std::max(sp->NA, sp->NB) * n2 * sizeof(int32_t) <= ((size_t)64 << 20) && std::max(sp->NA, sp->NB) < ((int64_t)1 << 30);
}

// ---- table-free variants (sq_set_option("etab", "alu")): same arithmetic, records from erec_closed ----------------------------------
__global__ void __launch_bounds__(256)
build_D_alu_kernel(const double* __restrict__ IN, double* __restrict__ D, int64_t W, int64_t j0, int64_t len, int n,
                   const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB, const int32_t* __restrict__ rankA,
                   const int32_t* __restrict__ rankB, int64_t NB, int64_t row_begin) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= W) return;
  const int64_t j = j0 + t;
  const int n2 = n * n;
  if (j >= len) {
    for (int slot = 0; slot < n2; ++slot) D[(int64_t)slot * W + t] = 0.0;
    return;
  }
  const int64_t ia_loc = j / NB, ib = j - ia_loc * NB;
  const uint32_t a = __ldg(strA + row_begin + ia_loc), b = __ldg(strB + ib);
  int slot = 0;
  for (int p = 0; p < n; ++p)
    for (int q = 0; q < n; ++q, ++slot) {
      double v = 0.0;
      const ERec ra = erec_closed(p, q, 0), rb = erec_closed(p, q, 1);
      if ((a & ra.tocc) == ra.tocc && (a & ra.temp) == 0u) {
        const uint32_t sa = a ^ ra.flip;
        const int par = (__popc(sa & ra.parS) + __popc(b & ra.parO)) & 1;
        const double x = IN[((int64_t)__ldg(rankA + sa) - row_begin) * NB + ib];
        v += (par ? -ra.s0 : ra.s0) * x;
      }
      if ((b & rb.tocc) == rb.tocc && (b & rb.temp) == 0u) {
        const uint32_t sb = b ^ rb.flip;
        const int par = (__popc(sb & rb.parS) + __popc(a & rb.parO)) & 1;
        const double x = IN[ia_loc * NB + __ldg(rankB + sb)];
        v += (par ? -rb.s0 : rb.s0) * x;
      }
      D[(int64_t)slot * W + t] = v;
    }
}

// Panel of the symmetric / antisymmetric generators (2-RDM of one real vector, see gram_sym_kernel in sqsv_dmma.cu):
//   rows [0, nS):       <J_t| S_pq |in>,  S_pq = E_pq + E_qp (p > q), S_pp = E_pp,        slot = p (p + 1) / 2 + q
//   rows [nS, n^2):     <J_t| A_pq |in>,  A_pq = E_pq - E_qp (p > q),                     slot = nS + p (p - 1) / 2 + q
__global__ void __launch_bounds__(256)
build_DSA_kernel(const double* __restrict__ IN, double* __restrict__ D, int64_t W, int64_t j0, int64_t len, int n,
                 const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB, const int32_t* __restrict__ rankA,
                 const int32_t* __restrict__ rankB, int64_t NB, int64_t row_begin) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= W) return;
  const int64_t j = j0 + t;
  const int n2 = n * n, nS = n * (n + 1) / 2;
  if (j >= len) {
    for (int slot = 0; slot < n2; ++slot) D[(int64_t)slot * W + t] = 0.0;
    return;
  }
  const int64_t ia_loc = j / NB, ib = j - ia_loc * NB;
  const uint32_t a = __ldg(strA + row_begin + ia_loc), b = __ldg(strB + ib);
  auto elem = [&](int p, int q) -> double {
    double v = 0.0;
    const ERec ra = erec_closed(p, q, 0), rb = erec_closed(p, q, 1);
    if ((a & ra.tocc) == ra.tocc && (a & ra.temp) == 0u) {
      const uint32_t sa = a ^ ra.flip;
      const int par = (__popc(sa & ra.parS) + __popc(b & ra.parO)) & 1;
      v += (par ? -ra.s0 : ra.s0) * IN[((int64_t)__ldg(rankA + sa) - row_begin) * NB + ib];
    }
    if ((b & rb.tocc) == rb.tocc && (b & rb.temp) == 0u) {
      const uint32_t sb = b ^ rb.flip;
      const int par = (__popc(sb & rb.parS) + __popc(a & rb.parO)) & 1;
      v += (par ? -rb.s0 : rb.s0) * IN[ia_loc * NB + __ldg(rankB + sb)];
    }
    return v;
  };
  int ss = 0, as = nS;
  for (int p = 0; p < n; ++p)
    for (int q = 0; q <= p; ++q, ++ss) {
      if (p == q) {
        D[(int64_t)ss * W + t] = elem(p, p);
      } else {
        const double x = elem(p, q), y = elem(q, p);
        D[(int64_t)ss * W + t] = x + y;
        D[(int64_t)as * W + t] = x - y;
        ++as;
      }
    }
}

__global__ void __launch_bounds__(256)
build_Dsym_alu_kernel(const double* __restrict__ IN, double* __restrict__ D, int64_t W, int64_t j0, int64_t len, int n,
                      const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB, const int32_t* __restrict__ rankA,
                      const int32_t* __restrict__ rankB, int64_t NB, int64_t row_begin) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= W) return;
  const int64_t j = j0 + t;
  const int nS = n * (n + 1) / 2;
  if (j >= len) {
    for (int slot = 0; slot < nS; ++slot) D[(int64_t)slot * W + t] = 0.0;
    return;
  }
  const int64_t ia_loc = j / NB, ib = j - ia_loc * NB;
  const uint32_t a = __ldg(strA + row_begin + ia_loc), b = __ldg(strB + ib);
  auto elem = [&](int p, int q) -> double {
    double v = 0.0;
    const ERec ra = erec_closed(p, q, 0), rb = erec_closed(p, q, 1);
    if ((a & ra.tocc) == ra.tocc && (a & ra.temp) == 0u) {
      const uint32_t sa = a ^ ra.flip;
      const int par = (__popc(sa & ra.parS) + __popc(b & ra.parO)) & 1;
      v += (par ? -ra.s0 : ra.s0) * IN[((int64_t)__ldg(rankA + sa) - row_begin) * NB + ib];
    }
    if ((b & rb.tocc) == rb.tocc && (b & rb.temp) == 0u) {
      const uint32_t sb = b ^ rb.flip;
      const int par = (__popc(sb & rb.parS) + __popc(a & rb.parO)) & 1;
      v += (par ? -rb.s0 : rb.s0) * IN[ia_loc * NB + __ldg(rankB + sb)];
    }
    return v;
  };
  int slot = 0;
  for (int r = 0; r < n; ++r)
    for (int q = 0; q <= r; ++q, ++slot) D[(int64_t)slot * W + t] = (r == q) ? elem(r, r) : elem(r, q) + elem(q, r);
}

__global__ void __launch_bounds__(256)
scatter_E_alu_kernel(const double* __restrict__ IN, double* __restrict__ OUT, const double* __restrict__ F,
                     const double* __restrict__ kmat, const int* __restrict__ frow, int64_t W, int64_t j0, int64_t len, int n,
                     const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB, const int32_t* __restrict__ rankA,
                     const int32_t* __restrict__ rankB, int64_t NB, int64_t row_begin) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t j = j0 + t;
  if (t >= W || j >= len) return;
  const int64_t ia_loc = j / NB, ib = j - ia_loc * NB;
  const uint32_t a = __ldg(strA + row_begin + ia_loc), b = __ldg(strB + ib);
  const double cj = IN[j];
  double diag = 0.0;
  int slot = 0;
  for (int p = 0; p < n; ++p)
    for (int q = 0; q < n; ++q, ++slot) {
      const ERec ra = erec_closed(p, q, 0), rb = erec_closed(p, q, 1);
      const bool va = (a & ra.occ) == ra.occ && (a & ra.emp) == 0u;
      const bool vb = (b & rb.occ) == rb.occ && (b & rb.emp) == 0u;
      if (!va && !vb) continue;
      const double val = F[(int64_t)__ldg(frow + slot) * W + t] + __ldg(kmat + slot) * cj;   // frow: row of F that holds (p,q)
      if (va) {
        const int par = (__popc(a & ra.parS) + __popc(b & ra.parO)) & 1;
        const double sv = (par ? -ra.s0 : ra.s0) * val;
        if (ra.flip == 0u) diag += sv;
        else atomicAdd(OUT + ((int64_t)__ldg(rankA + (a ^ ra.flip)) - row_begin) * NB + ib, sv);
      }
      if (vb) {
        const int par = (__popc(b & rb.parS) + __popc(a & rb.parO)) & 1;
        const double sv = (par ? -rb.s0 : rb.s0) * val;
        if (rb.flip == 0u) diag += sv;
        else atomicAdd(OUT + ia_loc * NB + __ldg(rankB + (b ^ rb.flip)), sv);
      }
    }
  atomicAdd(OUT + j, diag);
}

// ---- row kernels (measured alternative, off by default) ---------------------------------------------------------------
// Idea: the beta partner of column ib under E_rs and its sign do not depend on the row, so they can be tabulated once per
// space as one 32-bit word per (rs, ib) and read coalesced; and all beta partners of a row lie in that row, so a CTA that
// owns (part of) ONE row can keep the row in shared memory -- the gather reads it there, the scatter accumulates there
// (shared-memory atomics) and flushes once; the alpha part becomes uniform per CTA.  Same arithmetic per element, parity
// tested against the determinant-per-thread kernels (test_sigma_and_rdm_kernel_variants_agree).
// Measured at CAS(16,16) (profiles/r1_ab_rows_kernels_cas16.txt): SLOWER -- sigma 576 ms against 468 ms, RDMs 759 ms against
// 729 ms with 1024 threads per CTA, worse with fewer.  A 103 KB row allows two CTAs per SM and a panel of ~41 rows gives
// few CTAs, while the determinant-per-thread kernels run 64 warps per SM and find the row in L1 anyway (consecutive CTAs
// work on the same row).  Kept behind sq_set_option("rows", "1") with its geometry knob ("rows_cfg") as the evidence.
#define SQ_TINV 0x3FFFFFFFu   // table word: [29:0] partner column (SQ_TINV: none), bit 30 beta-string sign, bit 31 alpha-op parity of b
struct RowU {
  int32_t row;      // alpha partner row under this slot, -1 if the slot does not act on the row's alpha string
  uint32_t bits;    // bit 0: alpha sign (string part and s0); bit 1: parity of the alpha string under the BETA operator
};
static int g_rows_threads = 1024, g_rows_ch = 0;   // CTA size and column chunks per row (0: one wave of two CTAs per SM)
void sq_hamiltonian_set_rows_mode(int on) { g_rows_kernels = on ? 1 : 0; }
void sq_hamiltonian_set_rows_cfg(int threads, int ch) {
  g_rows_threads = (threads >= 32 && threads <= 1024) ? (threads / 32) * 32 : 1024;
  g_rows_ch = ch > 0 ? ch : 0;
}

static size_t rows_smem(const sq_space* sp) {
  const size_t n2 = (size_t)sp->n_orb * sp->n_orb;
  return sizeof(double) * (size_t)sp->NB + sizeof(RowU) * n2 + sizeof(double) * n2 + sizeof(int) * n2;
}
static bool rows_kernels_fit(const sq_space* sp) { return rows_smem(sp) <= 200 * 1024 && sp->NB < (int64_t)SQ_TINV; }

static int build_beta_tables(sq_space* sp, HamWork* w) {
  const int n = sp->n_orb, n2 = n * n;
  const int64_t NB = sp->NB;
  std::vector<uint32_t> tg((size_t)n2 * NB), ts((size_t)n2 * NB);
  for (int slot = 0; slot < n2; ++slot) {
    const ERec ra = w->h_etab[2 * (size_t)slot], rb = w->h_etab[2 * (size_t)slot + 1];
    for (int64_t ib = 0; ib < NB; ++ib) {
      const uint32_t b = sp->strB[ib];
      const uint32_t apar = (uint32_t)(__builtin_popcount(b & ra.parO) & 1) << 31;
      uint32_t g = SQ_TINV, sc = SQ_TINV;
      if ((b & rb.tocc) == rb.tocc && (b & rb.temp) == 0u) {           // gather form: b is the target, sb the source
        const uint32_t sb = b ^ rb.flip;
        const uint32_t sgn = (uint32_t)((__builtin_popcount(sb & rb.parS) & 1) ^ (rb.s0 < 0 ? 1 : 0));
        g = (uint32_t)sp->rankB[sb] | (sgn << 30);
      }
      if ((b & rb.occ) == rb.occ && (b & rb.emp) == 0u) {              // scatter form: b is the source
        const uint32_t sgn = (uint32_t)((__builtin_popcount(b & rb.parS) & 1) ^ (rb.s0 < 0 ? 1 : 0));
        sc = (uint32_t)sp->rankB[b ^ rb.flip] | (sgn << 30);
      }
      tg[(size_t)slot * NB + ib] = g | apar;
      ts[(size_t)slot * NB + ib] = sc | apar;
    }
  }
  SQ_CUDA(cudaMalloc(&w->d_tabG, sizeof(uint32_t) * tg.size()));
  SQ_CUDA(cudaMalloc(&w->d_tabS, sizeof(uint32_t) * ts.size()));
  SQ_CUDA(cudaMemcpy(w->d_tabG, tg.data(), sizeof(uint32_t) * tg.size(), cudaMemcpyHostToDevice));
  SQ_CUDA(cudaMemcpy(w->d_tabS, ts.data(), sizeof(uint32_t) * ts.size(), cudaMemcpyHostToDevice));
  return SQ_OK;
}

// CTA = (row ia, column chunk): D[slot][t] for the determinants of that chunk that fall into the panel [j0, jend)
template <bool SYM>
__global__ void __launch_bounds__(1024)
build_D_rows_kernel(const double* __restrict__ IN, double* __restrict__ D, int64_t W, int64_t j0, int64_t jend,
                    const ERec* __restrict__ etab, int n, const uint32_t* __restrict__ strA,
                    const int32_t* __restrict__ rankA, const uint32_t* __restrict__ tabG, int64_t NB, int64_t ia_first,
                    int CH) {
  extern __shared__ double rowsm[];                 // the row's NB amplitudes, then the per-slot alpha records
  const int n2 = n * n;
  RowU* u = reinterpret_cast<RowU*>(rowsm + NB);
  const int64_t ia = ia_first + blockIdx.x / CH;
  const int ch = blockIdx.x % CH;
  const uint32_t a = __ldg(strA + ia);
  for (int slot = threadIdx.x; slot < n2; slot += blockDim.x) {
    const ERec ra = etab[2 * slot], rb = etab[2 * slot + 1];
    RowU r;
    r.row = -1;
    r.bits = 0u;
    if ((a & ra.tocc) == ra.tocc && (a & ra.temp) == 0u) {
      const uint32_t sa = a ^ ra.flip;
      r.row = __ldg(rankA + sa);
      r.bits = (uint32_t)((__popc(sa & ra.parS) & 1) ^ (ra.s0 < 0 ? 1 : 0));
    }
    r.bits |= (uint32_t)(__popc(a & rb.parO) & 1) << 1;
    u[slot] = r;
  }
  const double* row = IN + ia * NB;
  for (int64_t i = threadIdx.x; i < NB; i += blockDim.x) rowsm[i] = row[i];
  __syncthreads();
  const int64_t c0 = NB * ch / CH, c1 = NB * (ch + 1) / CH;
  for (int64_t ib = c0 + threadIdx.x; ib < c1; ib += blockDim.x) {
    const int64_t j = ia * NB + ib;
    if (j < j0 || j >= jend) continue;
    const int64_t t = j - j0;
    auto elem = [&](int slot) -> double {
      const uint32_t e = __ldg(tabG + (int64_t)slot * NB + ib);
      const RowU r = u[slot];
      double v = 0.0;
      if (r.row >= 0) {
        const double x = IN[(int64_t)r.row * NB + ib];
        v += ((r.bits ^ (e >> 31)) & 1u) ? -x : x;
      }
      const uint32_t idx = e & SQ_TINV;
      if (idx != SQ_TINV) {
        const double x = rowsm[idx];
        v += (((e >> 30) ^ (r.bits >> 1)) & 1u) ? -x : x;
      }
      return v;
    };
    if (SYM) {
      int slot = 0;
      for (int r = 0; r < n; ++r)
        for (int q = 0; q <= r; ++q, ++slot)
          D[(int64_t)slot * W + t] = (r == q) ? elem(r * n + r) : elem(r * n + q) + elem(q * n + r);
    } else {
      for (int slot = 0; slot < n2; ++slot) D[(int64_t)slot * W + t] = elem(slot);
    }
  }
}

// CTA = (row ia, column chunk): OUT[E_pq J] += sign (F[pq][t] + k[pq] IN[J]); beta targets accumulate in the shared row
__global__ void __launch_bounds__(1024)
scatter_E_rows_kernel(const double* __restrict__ IN, double* __restrict__ OUT, const double* __restrict__ F,
                      const double* __restrict__ kmat, const int* __restrict__ frow, int64_t W, int64_t j0, int64_t jend,
                      const ERec* __restrict__ etab, int n2, const uint32_t* __restrict__ strA,
                      const int32_t* __restrict__ rankA, const uint32_t* __restrict__ tabS, int64_t NB, int64_t ia_first,
                      int CH) {
  extern __shared__ double acc[];                   // NB accumulators, alpha records, k_pq, F rows
  RowU* u = reinterpret_cast<RowU*>(acc + NB);
  double* ks = reinterpret_cast<double*>(u + n2);
  int* fr = reinterpret_cast<int*>(ks + n2);
  const int64_t ia = ia_first + blockIdx.x / CH;
  const int ch = blockIdx.x % CH;
  const uint32_t a = __ldg(strA + ia);
  for (int slot = threadIdx.x; slot < n2; slot += blockDim.x) {
    const ERec ra = etab[2 * slot], rb = etab[2 * slot + 1];
    RowU r;
    r.row = -1;
    r.bits = 0u;
    if ((a & ra.occ) == ra.occ && (a & ra.emp) == 0u) {
      r.row = __ldg(rankA + (a ^ ra.flip));
      r.bits = (uint32_t)((__popc(a & ra.parS) & 1) ^ (ra.s0 < 0 ? 1 : 0));
    }
    r.bits |= (uint32_t)(__popc(a & rb.parO) & 1) << 1;
    u[slot] = r;
    ks[slot] = kmat[slot];
    fr[slot] = frow[slot];
  }
  for (int64_t i = threadIdx.x; i < NB; i += blockDim.x) acc[i] = 0.0;
  __syncthreads();
  const int64_t c0 = NB * ch / CH, c1 = NB * (ch + 1) / CH;
  for (int64_t ib = c0 + threadIdx.x; ib < c1; ib += blockDim.x) {
    const int64_t j = ia * NB + ib;
    if (j < j0 || j >= jend) continue;
    const int64_t t = j - j0;
    const double cj = IN[j];
    double diag = 0.0;
    for (int slot = 0; slot < n2; ++slot) {
      const uint32_t e = __ldg(tabS + (int64_t)slot * NB + ib);
      const RowU r = u[slot];
      const uint32_t idx = e & SQ_TINV;
      if (r.row < 0 && idx == SQ_TINV) continue;
      const double val = F[(int64_t)fr[slot] * W + t] + ks[slot] * cj;
      if (r.row >= 0) {
        const double sv = ((r.bits ^ (e >> 31)) & 1u) ? -val : val;
        if (r.row == ia) diag += sv;
        else atomicAdd(OUT + (int64_t)r.row * NB + ib, sv);
      }
      if (idx != SQ_TINV) {
        const double sv = (((e >> 30) ^ (r.bits >> 1)) & 1u) ? -val : val;
        if ((int64_t)idx == ib) diag += sv;
        else atomicAdd(acc + idx, sv);
      }
    }
    atomicAdd(acc + ib, diag);
  }
  __syncthreads();
  double* orow = OUT + ia * NB;
  for (int64_t i = threadIdx.x; i < NB; i += blockDim.x) {
    const double v = acc[i];
    if (v != 0.0) atomicAdd(orow + i, v);
  }
}

// launch geometry of the row kernels for the panel [j0, j0 + W)
struct RowGrid {
  int64_t ia_first, jend;
  int CH;
  unsigned grid;
};
static RowGrid row_grid(const sq_space* sp, const HamWork* w, int64_t j0) {
  RowGrid g;
  const int64_t len = sp->local_len();
  g.jend = (j0 + w->W < len) ? j0 + w->W : len;
  g.ia_first = j0 / sp->NB;
  const int64_t ia_last = (g.jend - 1) / sp->NB;
  const int64_t nrows = ia_last - g.ia_first + 1;
  int ch = g_rows_ch > 0 ? g_rows_ch : (int)((2 * 148) / nrows);   // default: two CTAs per SM (103 KB row at CAS(16,16)), one wave
  g.CH = ch < 1 ? 1 : (ch > 64 ? 64 : ch);
  g.grid = (unsigned)(nrows * g.CH);
  return g;
}

static int check_full_space(sq_space* sp, const char* who) {
  if (sp->device < 0) {
    sq_set_error("%s: host-only space (device = -1) cannot run kernels", who);
    return SQ_ERR_INVALID;
  }
  if (sp->row_begin != 0 || sp->row_end != sp->NA) {
    sq_set_error("%s: not available on an alpha-sharded vector yet", who);
    return SQ_ERR_UNSUPPORTED;
  }
  return SQ_OK;
}

// Make c_etab hold this space's table (no-op when it already does); *use_const = false -> shared-memory kernels.
static int bind_etab(sq_space* sp, HamWork* w, cudaStream_t st, bool* use_const) {
  const int n = sp->n_orb;
  *use_const = g_etab_const && n <= SQ_ETAB_CONST_ORBS;
  if (!*use_const) return SQ_OK;
  const sq_space*& owner = g_etab_owner[sp->device];
  if (owner != sp) {
    if (owner) SQ_CUDA(cudaDeviceSynchronize());   // kernels of the previous owner may still be reading the table
    SQ_CUDA(cudaMemcpyToSymbolAsync(c_etab, w->d_etab, sizeof(ERec) * 2 * (size_t)n * n, 0, cudaMemcpyDeviceToDevice, st));
    owner = sp;
  }
  return SQ_OK;
}

template <typename K>
static void allow_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

static int launch_error(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sq_set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}

// D panel of [j0, j0 + W): n^2 rows, or the n (n + 1) / 2 symmetrised rows when sym
static int launch_build_D(sq_space* sp, HamWork* w, const double* in, double* D, int64_t j0, cudaStream_t st, bool use_const,
                          bool sym = false, const PeerView* pv = nullptr) {
  const int n = sp->n_orb, n2 = n * n;
  if (pv) {   // alpha-sharded vector: general n^2 panel, alpha partners through peer memory
    const size_t smem_p = sizeof(ERec) * 2 * (size_t)n2;
    allow_smem(build_D_peer_kernel, smem_p);
    build_D_peer_kernel<<<(unsigned)(w->W / 256), 256, smem_p, st>>>(*pv, in, D, w->W, j0, sp->local_len(), w->d_etab, n2,
                                                                     sp->d_strA, sp->d_strB, sp->d_rankA, sp->d_rankB, sp->NB,
                                                                     sp->row_begin);
    return launch_error("build_D_peer_kernel");
  }
  if (g_etab_tab && sym && w->d_pgA) {   // per-string partner tables
    build_Dsym_tab_kernel<<<(unsigned)(w->W / 256), 256, 0, st>>>(in, D, w->W, j0, sp->local_len(), w->d_pgA, w->d_pgB, w->d_parO, n,
                                                                  sp->d_strA, sp->d_strB, sp->NB);
    return launch_error("build_Dsym_tab_kernel");
  }
  if (g_etab_alu) {   // table-free records (candidate, off by default)
    const unsigned grid_a = (unsigned)(w->W / 256);
    if (sym)
      build_Dsym_alu_kernel<<<grid_a, 256, 0, st>>>(in, D, w->W, j0, sp->local_len(), n, sp->d_strA, sp->d_strB, sp->d_rankA, sp->d_rankB,
                                                     sp->NB, sp->row_begin);
    else
      build_D_alu_kernel<<<grid_a, 256, 0, st>>>(in, D, w->W, j0, sp->local_len(), n, sp->d_strA, sp->d_strB, sp->d_rankA, sp->d_rankB,
                                                  sp->NB, sp->row_begin);
    return launch_error("build_D_alu_kernel");
  }
  if (g_rows_kernels && w->d_tabG) {
    const RowGrid g = row_grid(sp, w, j0);
    const size_t smem = rows_smem(sp);
    if (g.jend - j0 < w->W)   // last panel: the GEMM reads all W columns, the kernel writes only the live ones
      SQ_CUDA(cudaMemsetAsync(D, 0, sizeof(double) * (size_t)(sym ? n * (n + 1) / 2 : n2) * (size_t)w->W, st));
    if (sym) {
      allow_smem(build_D_rows_kernel<true>, smem);
      build_D_rows_kernel<true><<<g.grid, g_rows_threads, smem, st>>>(in, D, w->W, j0, g.jend, w->d_etab, n, sp->d_strA, sp->d_rankA,
                                                           w->d_tabG, sp->NB, g.ia_first, g.CH);
    } else {
      allow_smem(build_D_rows_kernel<false>, smem);
      build_D_rows_kernel<false><<<g.grid, g_rows_threads, smem, st>>>(in, D, w->W, j0, g.jend, w->d_etab, n, sp->d_strA, sp->d_rankA,
                                                            w->d_tabG, sp->NB, g.ia_first, g.CH);
    }
    return launch_error("build_D_rows_kernel");
  }
  const unsigned grid = (unsigned)(w->W / 256);
  const size_t smem = use_const ? 0 : sizeof(ERec) * 2 * (size_t)n2;
  if (sym) {
    if (use_const) {
      build_Dsym_kernel<true><<<grid, 256, 0, st>>>(in, D, w->W, j0, sp->local_len(), w->d_etab, n, sp->d_strA, sp->d_strB,
                                                    sp->d_rankA, sp->d_rankB, sp->NB, sp->row_begin);
    } else {
      allow_smem(build_Dsym_kernel<false>, smem);
      build_Dsym_kernel<false><<<grid, 256, smem, st>>>(in, D, w->W, j0, sp->local_len(), w->d_etab, n, sp->d_strA, sp->d_strB,
                                                        sp->d_rankA, sp->d_rankB, sp->NB, sp->row_begin);
    }
    return launch_error("build_Dsym_kernel");
  }
  if (use_const) {
    build_D_kernel<true><<<grid, 256, 0, st>>>(in, D, w->W, j0, sp->local_len(), w->d_etab, n2, sp->d_strA, sp->d_strB,
                                               sp->d_rankA, sp->d_rankB, sp->NB, sp->row_begin);
  } else {
    allow_smem(build_D_kernel<false>, smem);
    build_D_kernel<false><<<grid, 256, smem, st>>>(in, D, w->W, j0, sp->local_len(), w->d_etab, n2, sp->d_strA, sp->d_strB,
                                                   sp->d_rankA, sp->d_rankB, sp->NB, sp->row_begin);
  }
  return launch_error("build_D_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------------
// Fused sigma kernel: gather -> DMMA -> scatter for a tile of 64 determinants, without the HBM round trips of the D and F
// panels (round 1: three kernels per panel, D and F written to and read from 1 GiB buffers).  Per tile:
//   A  every thread gathers its share of D[rs][t] = <J_t|E_rs (+ E_sr)|in> straight into shared memory (table-free E records);
//   B  F = Gm . D on the fp64 tensor cores (mma.sync m8n8k4, 4 x 2 warps, accumulators in registers), the integral matrix
//      streams through shared memory in double-buffered chunks of 8 columns (cp.async, it lives in L2);
//   C  the accumulators replace the D tile in shared memory and every thread scatters its share:
//      out[E_pq J_t] += sign (F[pq][t] + k_pq in[J_t]).
// Two CTAs per SM: while one is in phase B (tensor pipe) the other gathers or scatters (LSU pipe).  Same arithmetic as
// build_Dsym_kernel / sigma_dmma_kernel / scatter_E_kernel.  MEASURED SLOWER than that three-kernel pipeline (728 ms against
// 445 ms at CAS(16,16)): the gathers want 64 warps per SM to hide their latency and get 16 here, so the tensor pipe idles at 27 %.
// Parity-green and kept behind sq_set_option("sigma_fused", "1") as evidence; the default is the panel pipeline.
// ---------------------------------------------------------------------------------------------------------------------------
#define SF_BN 64
#define SF_LDB 68            // D / F tile row stride (doubles): = 4 (mod 16) -> conflict-free B fragments
#define SF_KC 8
#define SF_LDA 12            // Gm chunk [m][8] row stride: = 12 (mod 16) -> conflict-free A fragments
#define SF_THREADS 256

template <int MQ, bool SYM>   // MQ = row fragments per warp: 4 warps x MQ x 8 >= nrow generator rows
__global__ void __launch_bounds__(SF_THREADS, 2)
sigma_fused_kernel(const double* __restrict__ IN, double* __restrict__ OUT, const double* __restrict__ Gm, int ldg, int nrow,
                   const double* __restrict__ kmat, const int* __restrict__ frow, int n, int64_t len, int64_t n_tiles,
                   const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB, const int32_t* __restrict__ rankA,
                   const int32_t* __restrict__ rankB, int64_t NB, int64_t row_begin) {
  extern __shared__ __align__(16) double fsm[];
  constexpr int MP = 4 * MQ * 8;                                  // padded generator rows of the A chunks
  const int krows = (nrow + 7) & ~7;                              // rows of the D tile (k index), zero-padded
  double* const Ds = fsm;                                         // [krows][SF_LDB]   (D, later F)
  double* const As = Ds + (size_t)krows * SF_LDB;                 // 2 stages of [MP][SF_LDA]
  double* const ks = As + 2 * MP * SF_LDA;                        // [n * n]
  int* const frs = reinterpret_cast<int*>(ks + n * n);            // [n * n]
  short2* const slotrq = reinterpret_cast<short2*>(frs + n * n);  // [nrow] (r, q) of every D row
  const int n2 = n * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int t = threadIdx.x & (SF_BN - 1), part = threadIdx.x >> 6;   // gather / scatter: determinant t of the tile, quarter of the rows
  for (int i = threadIdx.x; i < n2; i += SF_THREADS) {
    ks[i] = __ldg(kmat + i);
    frs[i] = __ldg(frow + i);
  }
  if (SYM) {
    for (int r = 0, s = 0; r < n; ++r)
      for (int q = 0; q <= r; ++q, ++s)
        if (s % SF_THREADS == (int)threadIdx.x) slotrq[s] = make_short2((short)r, (short)q);
  } else {
    for (int s = threadIdx.x; s < n2; s += SF_THREADS) slotrq[s] = make_short2((short)(s / n), (short)(s % n));
  }
  const uint32_t abase = (uint32_t)__cvta_generic_to_shared(As);
  const int n_chunks = (nrow + SF_KC - 1) / SF_KC;
  auto issue_A = [&](int c, int slot) {
    const int k0 = c * SF_KC;
    const uint32_t sa = abase + (uint32_t)(slot * MP * SF_LDA) * 8u;
    for (int id = threadIdx.x; id < MP * 4; id += SF_THREADS) {
      const int m = id >> 2, ch = id & 3;
      const bool ok = m < nrow && k0 + ch * 2 < ldg;
      cp16(sa + (uint32_t)(m * SF_LDA + ch * 2) * 8u, Gm + (size_t)(ok ? m : 0) * ldg + (ok ? k0 + ch * 2 : 0), ok ? 16 : 0);
    }
  };
  __syncthreads();
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t j = tile * SF_BN + t;
    const bool live = j < len;
    int64_t ia_loc = 0, ib = 0;
    uint32_t a = 0, b = 0;
    if (live) {
      ia_loc = j / NB;
      ib = j - ia_loc * NB;
      a = __ldg(strA + row_begin + ia_loc);
      b = __ldg(strB + ib);
    }
    issue_A(0, 0);            // the first integral chunk arrives while the gathers run
    cp_commit();
    // ---- phase A: D tile ----
    auto elem = [&](int p, int q) -> double {   // <J|E_pq|in>, gather form (build_D_alu_kernel)
      double v = 0.0;
      const ERec ra = erec_closed(p, q, 0), rb = erec_closed(p, q, 1);
      if ((a & ra.tocc) == ra.tocc && (a & ra.temp) == 0u) {
        const uint32_t sa = a ^ ra.flip;
        const int par = (__popc(sa & ra.parS) + __popc(b & ra.parO)) & 1;
        const double x = IN[((int64_t)__ldg(rankA + sa) - row_begin) * NB + ib];
        v += (par ? -ra.s0 : ra.s0) * x;
      }
      if ((b & rb.tocc) == rb.tocc && (b & rb.temp) == 0u) {
        const uint32_t sb = b ^ rb.flip;
        const int par = (__popc(sb & rb.parS) + __popc(a & rb.parO)) & 1;
        const double x = IN[ia_loc * NB + __ldg(rankB + sb)];
        v += (par ? -rb.s0 : rb.s0) * x;
      }
      return v;
    };
    for (int s = part; s < krows; s += 4) {
      double v = 0.0;
      if (live && s < nrow) {
        const short2 rq = slotrq[s];
        v = (SYM && rq.x != rq.y) ? elem(rq.x, rq.y) + elem(rq.y, rq.x) : elem(rq.x, rq.y);
      }
      Ds[s * SF_LDB + t] = v;
    }
    const double cj = live ? IN[j] : 0.0;
    __syncthreads();
    // ---- phase B: F = Gm . D ----
    double acc[MQ][4][2];
#pragma unroll
    for (int i = 0; i < MQ; ++i)
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) acc[i][jn][0] = acc[i][jn][1] = 0.0;
    for (int it = 0; it < n_chunks; ++it) {
      cp_wait<0>();
      __syncthreads();                                   // chunk `it` has landed; the other stage is free
      if (it + 1 < n_chunks) issue_A(it + 1, (it + 1) & 1);
      cp_commit();
      const double* as = As + (it & 1) * MP * SF_LDA + (wm * MQ * 8) * SF_LDA;
      const double* bs = Ds + (size_t)(it * SF_KC) * SF_LDB + wn * 32;
#pragma unroll
      for (int k4 = 0; k4 < SF_KC / 4; ++k4) {
        double bf[4];
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) bf[jn] = bs[(k4 * 4 + (lane & 3)) * SF_LDB + jn * 8 + (lane >> 2)];
#pragma unroll
        for (int i = 0; i < MQ; ++i) {
          const double av = as[(i * 8 + (lane >> 2)) * SF_LDA + k4 * 4 + (lane & 3)];
#pragma unroll
          for (int jn = 0; jn < 4; ++jn) dmma884(acc[i][jn][0], acc[i][jn][1], av, bf[jn]);
        }
      }
    }
    __syncthreads();                                     // every warp is done reading the D tile
    // ---- phase C: F tile into shared memory, scatter ----
#pragma unroll
    for (int i = 0; i < MQ; ++i) {
      const int m = (wm * MQ + i) * 8 + (lane >> 2);
      if (m < nrow) {
        double* p = Ds + (size_t)m * SF_LDB + wn * 32 + 2 * (lane & 3);
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) *reinterpret_cast<double2*>(p + jn * 8) = make_double2(acc[i][jn][0], acc[i][jn][1]);
      }
    }
    __syncthreads();
    if (live) {
      double diag = 0.0;
      for (int slot = part; slot < n2; slot += 4) {
        const int p = slot / n, q = slot - p * n;
        const ERec ra = erec_closed(p, q, 0), rb = erec_closed(p, q, 1);
        const bool va = (a & ra.occ) == ra.occ && (a & ra.emp) == 0u;
        const bool vb = (b & rb.occ) == rb.occ && (b & rb.emp) == 0u;
        if (!va && !vb) continue;
        const double val = Ds[(size_t)frs[slot] * SF_LDB + t] + ks[slot] * cj;
        if (va) {
          const int par = (__popc(a & ra.parS) + __popc(b & ra.parO)) & 1;
          const double sv = (par ? -ra.s0 : ra.s0) * val;
          if (ra.flip == 0u) diag += sv;
          else atomicAdd(OUT + ((int64_t)__ldg(rankA + (a ^ ra.flip)) - row_begin) * NB + ib, sv);
        }
        if (vb) {
          const int par = (__popc(b & rb.parS) + __popc(a & rb.parO)) & 1;
          const double sv = (par ? -rb.s0 : rb.s0) * val;
          if (rb.flip == 0u) diag += sv;
          else atomicAdd(OUT + ia_loc * NB + __ldg(rankB + (b ^ rb.flip)), sv);
        }
      }
      atomicAdd(OUT + j, diag);
    }
    __syncthreads();                                     // the tile buffer is rewritten by the next tile's gathers
  }
}

static int g_sigma_fused = 0;   // measured at CAS(16,16): 728 ms fused against 445 ms for the three-kernel panel pipeline (profiles/r2_visit6_ab_sigma_fused.txt):
                                // with 16 warps per SM the gathers of phase A are latency-bound (tensor pipe 27 %); kept as sq_set_option("sigma_fused", "1")
void sq_hamiltonian_set_sigma_fused(int on) { g_sigma_fused = on ? 1 : 0; }

static size_t sigma_fused_smem(int mq, int nrow, int n) {
  const int krows = (nrow + 7) & ~7, n2 = n * n;
  return sizeof(double) * ((size_t)krows * SF_LDB + 2 * (size_t)(4 * mq * 8) * SF_LDA + n2) + sizeof(int) * n2 + sizeof(short2) * (size_t)(n2 > nrow ? n2 : nrow) + 16;
}

template <int MQ, bool SYM>
static int launch_sigma_fused_t(sq_space* sp, const double* in, double* out, const double* d_G, int ldg, int nrow, const double* d_k,
                                const int* d_frow, int n_sm, cudaStream_t st) {
  const int n = sp->n_orb;
  const size_t smem = sigma_fused_smem(MQ, nrow, n);
  static size_t attr = 0;
  if (smem > attr) {
    SQ_CUDA(cudaFuncSetAttribute(sigma_fused_kernel<MQ, SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int64_t len = sp->local_len(), n_tiles = (len + SF_BN - 1) / SF_BN;
  const int64_t grid = std::min<int64_t>(n_tiles, (int64_t)2 * n_sm * 4);   // persistent CTAs, a few waves for load balance
  sigma_fused_kernel<MQ, SYM><<<(unsigned)grid, SF_THREADS, smem, st>>>(in, out, d_G, ldg, nrow, d_k, d_frow, n, len, n_tiles, sp->d_strA,
                                                                       sp->d_strB, sp->d_rankA, sp->d_rankB, sp->NB, sp->row_begin);
  return launch_error("sigma_fused_kernel");
}

// returns SQ_ERR_UNSUPPORTED (without an error message) when the fused kernel does not cover this shape
static int launch_sigma_fused(sq_space* sp, const double* in, double* out, const double* d_G, int ldg, int nrow, bool sym,
                              const double* d_k, const int* d_frow, int n_sm, cudaStream_t st) {
  const int mq = (nrow + 31) / 32;
  if (mq < 1 || mq > 5 || sigma_fused_smem(mq, nrow, sp->n_orb) > 110 * 1024) return SQ_ERR_UNSUPPORTED;
#define SF_CASE(M)                                                                                                     \
  case M:                                                                                                              \
    return sym ? launch_sigma_fused_t<M, true>(sp, in, out, d_G, ldg, nrow, d_k, d_frow, n_sm, st)                     \
               : launch_sigma_fused_t<M, false>(sp, in, out, d_G, ldg, nrow, d_k, d_frow, n_sm, st);
  switch (mq) {
    SF_CASE(1) SF_CASE(2) SF_CASE(3) SF_CASE(4) SF_CASE(5)
    default: break;
  }
#undef SF_CASE
  return SQ_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Spin-flip symmetric vectors: half of the sigma build.
//
// U = "flip the spin of every electron" maps |A,B> (alpha mask A, beta mask B; spin orbitals interleaved a0 b0 a1 b1 ...) to
// phi(A,B) |B,A> with phi = (-1)^popc(A & B) (one transposition per doubly occupied orbital).  A spin-free Hamiltonian commutes
// with U, and so does every spin-adapted ansatz operator (sa_single, pair doubles, sa_double_k): a tUPS / QNP / SA-UCC state
// built on a closed-shell reference is an eigenvector, U psi = lambda psi with lambda = +-1, i.e.
//     c[B,A] = lambda phi(A,B) c[A,B],
// and the same holds for sigma = H psi and for every column F[pq][.] of the contraction with the integrals.  With n_alpha = n_beta
// the panels of the Knowles-Handy build then only need the determinants J = (A,B) with index(A) <= index(B): for a contribution
// w of J to a target T,   T above / on the diagonal: sigma[T] += w;   T below / on the diagonal: sigma[T^T] += lambda phi(T) w
// (a diagonal T gets both; a diagonal J contributes only to targets above / on the diagonal), and the lower triangle is filled
// from the upper one at the end.  Half of the gathers, of the tensor-core work and of the atomics.  The symmetry is a property
// of the VECTOR: it is measured before every build (max |c[B,A] - lambda phi c[A,B]| <= 1e-12 max|c|), anything else takes the
// full build.  Reference: the full expectation-value loop of ups_wavefunction.py:770-784 makes no use of it.
// ---------------------------------------------------------------------------------------------------------------------------
static int g_sigma_spinsym = 1;   // sq_set_option("sigma_spinsym", "0"): always the full build
void sq_hamiltonian_set_sigma_spinsym(int on) { g_sigma_spinsym = on ? 1 : 0; }

__device__ __forceinline__ void tri_unrank(int64_t j, int64_t N, int64_t* ia, int64_t* ib) {
  // j = ia * N - ia (ia - 1) / 2 + (ib - ia), ib >= ia
  const double b = 2.0 * (double)N + 1.0;
  int64_t r = (int64_t)((b - sqrt(b * b - 8.0 * (double)j)) * 0.5);
  if (r < 0) r = 0;
  if (r > N - 1) r = N - 1;
  while (r > 0 && r * N - r * (r - 1) / 2 > j) --r;
  while (r + 1 < N && (r + 1) * N - (r + 1) * r / 2 <= j) ++r;
  *ia = r;
  *ib = r + (j - (r * N - r * (r - 1) / 2));
}

// res[0] = max |c|, res[1] = max |c[B,A] - phi c[A,B]|, res[2] = max |c[B,A] + phi c[A,B]| (bit patterns of non-negative doubles
// order like integers: atomicMax on the 64-bit words); 32 x 32 tiles transposed through shared memory, upper tiles only
__global__ void __launch_bounds__(256)
spinsym_check_kernel(const double* __restrict__ C, int64_t N, const uint32_t* __restrict__ str, unsigned long long* __restrict__ res) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t bi = blockIdx.y, bj = blockIdx.x;
  double m0 = 0.0, m1 = 0.0, m2 = 0.0;
  if (bj >= bi) {
    for (int k = ty; k < 32; k += 8) {   // lower tile (rows of block bj, columns of block bi), read along its rows
      const int64_t r = bj * 32 + k, c = bi * 32 + tx;
      tile[k][tx] = (r < N && c < N) ? C[r * N + c] : 0.0;
    }
  }
  __syncthreads();
  if (bj >= bi) {
    for (int k = ty; k < 32; k += 8) {
      const int64_t r = bi * 32 + k, c = bj * 32 + tx;   // element (r, c) of the upper tile; its mirror is tile[tx][k]
      if (r < N && c < N) {
        const double x = C[r * N + c], y = tile[tx][k];
        const double ph = (__popc(__ldg(str + r) & __ldg(str + c)) & 1) ? -x : x;
        m0 = fmax(m0, fmax(fabs(x), fabs(y)));
        m1 = fmax(m1, fabs(y - ph));
        m2 = fmax(m2, fabs(y + ph));
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    m0 = fmax(m0, __shfl_down_sync(0xffffffffu, m0, off));
    m1 = fmax(m1, __shfl_down_sync(0xffffffffu, m1, off));
    m2 = fmax(m2, __shfl_down_sync(0xffffffffu, m2, off));
  }
  if (tx == 0 && bj >= bi) {
    atomicMax(res + 0, (unsigned long long)__double_as_longlong(m0));
    atomicMax(res + 1, (unsigned long long)__double_as_longlong(m1));
    atomicMax(res + 2, (unsigned long long)__double_as_longlong(m2));
  }
}

// OUT[B,A] = lambda phi(A,B) OUT[A,B] for index(A) < index(B): the lower triangle from the upper one
__global__ void __launch_bounds__(256)
spinsym_mirror_kernel(double* __restrict__ OUT, int64_t N, const uint32_t* __restrict__ str, double lambda) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  for (int k = ty; k < 32; k += 8) {
    const int64_t r = bi * 32 + k, c = bj * 32 + tx;
    double v = 0.0;
    if (r < N && c < N) {
      v = OUT[r * N + c];
      if (__popc(__ldg(str + r) & __ldg(str + c)) & 1) v = -v;
    }
    tile[k][tx] = lambda * v;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int64_t r = bj * 32 + k, c = bi * 32 + tx;   // element (r, c) of the lower tile = mirror of upper (c, r)
    if (r < N && c < N && r > c) OUT[r * N + c] = tile[tx][k];
  }
}

// the gather of build_Dsym_kernel for the determinants above / on the diagonal (linear index j over the upper triangle)
__global__ void __launch_bounds__(256)
build_Dsym_tri_kernel(const double* __restrict__ IN, double* __restrict__ D, int64_t W, int64_t j0, int64_t len,
                      const ERec* __restrict__ etab, int n, const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB,
                      const int32_t* __restrict__ rankA, const int32_t* __restrict__ rankB, int64_t NB) {
  const int n2 = n * n;
  const ERec* sm = stage_etab<false>(etab, n2);
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= W) return;
  const int64_t j = j0 + t;
  const int nS = n * (n + 1) / 2;
  if (j >= len) {
    for (int slot = 0; slot < nS; ++slot) D[(int64_t)slot * W + t] = 0.0;
    return;
  }
  int64_t ia, ib;
  tri_unrank(j, NB, &ia, &ib);
  const uint32_t a = __ldg(strA + ia), b = __ldg(strB + ib);
  auto elem = [&](int slot) -> double {
    double v = 0.0;
    const ERec ra = sm[2 * slot], rb = sm[2 * slot + 1];
    if ((a & ra.tocc) == ra.tocc && (a & ra.temp) == 0u) {
      const uint32_t sa = a ^ ra.flip;
      const int par = (__popc(sa & ra.parS) + __popc(b & ra.parO)) & 1;
      v += (par ? -ra.s0 : ra.s0) * IN[(int64_t)__ldg(rankA + sa) * NB + ib];
    }
    if ((b & rb.tocc) == rb.tocc && (b & rb.temp) == 0u) {
      const uint32_t sb = b ^ rb.flip;
      const int par = (__popc(sb & rb.parS) + __popc(a & rb.parO)) & 1;
      v += (par ? -rb.s0 : rb.s0) * IN[ia * NB + __ldg(rankB + sb)];
    }
    return v;
  };
  int slot = 0;
  for (int r = 0; r < n; ++r)
    for (int q = 0; q <= r; ++q, ++slot) D[(int64_t)slot * W + t] = (r == q) ? elem(r * n + r) : elem(r * n + q) + elem(q * n + r);
}

// the scatter of scatter_E_kernel for sources above / on the diagonal; targets below the diagonal are folded onto their mirror
__global__ void __launch_bounds__(256)
scatter_E_tri_kernel(const double* __restrict__ IN, double* __restrict__ OUT, const double* __restrict__ F,
                     const double* __restrict__ kmat, const int* __restrict__ frow, int64_t W, int64_t j0, int64_t len,
                     const ERec* __restrict__ etab, int n2, const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB,
                     const int32_t* __restrict__ rankA, const int32_t* __restrict__ rankB, int64_t NB, double lambda) {
  const ERec* sm = stage_etab<false>(etab, n2);
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t j = j0 + t;
  if (t >= W || j >= len) return;
  int64_t ia, ib;
  tri_unrank(j, NB, &ia, &ib);
  const uint32_t a = __ldg(strA + ia), b = __ldg(strB + ib);
  const bool src_diag = ia == ib;
  const double cj = IN[ia * NB + ib];
  double diag = 0.0;
  // one contribution w to target (ra_, rb_) with masks (ma, mb)
  auto put = [&](int64_t ra_, int64_t rb_, uint32_t ma, uint32_t mb, double wv) {
    if (ra_ <= rb_) atomicAdd(OUT + ra_ * NB + rb_, wv);
    if (ra_ >= rb_ && !src_diag) {
      const double mv = lambda * ((__popc(ma & mb) & 1) ? -wv : wv);
      atomicAdd(OUT + rb_ * NB + ra_, mv);
    }
  };
  for (int slot = 0; slot < n2; ++slot) {
    const ERec ra = sm[2 * slot], rb = sm[2 * slot + 1];
    const bool va = (a & ra.occ) == ra.occ && (a & ra.emp) == 0u;
    const bool vb = (b & rb.occ) == rb.occ && (b & rb.emp) == 0u;
    if (!va && !vb) continue;
    const double val = F[(int64_t)__ldg(frow + slot) * W + t] + __ldg(kmat + slot) * cj;
    if (va) {
      const int par = (__popc(a & ra.parS) + __popc(b & ra.parO)) & 1;
      const double sv = (par ? -ra.s0 : ra.s0) * val;
      if (ra.flip == 0u) diag += sv;
      else put((int64_t)__ldg(rankA + (a ^ ra.flip)), ib, a ^ ra.flip, b, sv);
    }
    if (vb) {
      const int par = (__popc(b & rb.parS) + __popc(a & rb.parO)) & 1;
      const double sv = (par ? -rb.s0 : rb.s0) * val;
      if (rb.flip == 0u) diag += sv;
      else put(ia, (int64_t)__ldg(rankB + (b ^ rb.flip)), a, b ^ rb.flip, sv);
    }
  }
  // the determinant itself: above the diagonal it is a plain upper target; on the diagonal it gets the direct term only
  atomicAdd(OUT + ia * NB + ib, diag);
}

// The S / A generator panel of build_DSA_kernel for the determinants above / on the diagonal of a spin-flip symmetric vector:
// <J^T|O|psi> = lambda phi(J) <J|O|psi> for every spin-free O, so the Gram matrices sum_J d_J d_J^T over all determinants are
// 2 sum_{J above} + sum_{J on the diagonal}; the weight enters as a factor sqrt(2) on the columns above the diagonal.
__global__ void __launch_bounds__(256)
build_DSA_tri_kernel(const double* __restrict__ IN, double* __restrict__ D, int64_t W, int64_t j0, int64_t len, int n,
                     const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB, const int32_t* __restrict__ rankA,
                     const int32_t* __restrict__ rankB, int64_t NB) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= W) return;
  const int64_t j = j0 + t;
  const int n2 = n * n, nS = n * (n + 1) / 2;
  if (j >= len) {
    for (int slot = 0; slot < n2; ++slot) D[(int64_t)slot * W + t] = 0.0;
    return;
  }
  int64_t ia, ib;
  tri_unrank(j, NB, &ia, &ib);
  const uint32_t a = __ldg(strA + ia), b = __ldg(strB + ib);
  const double wgt = (ia == ib) ? 1.0 : 1.4142135623730951;
  auto elem = [&](int p, int q) -> double {
    double v = 0.0;
    const ERec ra = erec_closed(p, q, 0), rb = erec_closed(p, q, 1);
    if ((a & ra.tocc) == ra.tocc && (a & ra.temp) == 0u) {
      const uint32_t sa = a ^ ra.flip;
      const int par = (__popc(sa & ra.parS) + __popc(b & ra.parO)) & 1;
      v += (par ? -ra.s0 : ra.s0) * IN[(int64_t)__ldg(rankA + sa) * NB + ib];
    }
    if ((b & rb.tocc) == rb.tocc && (b & rb.temp) == 0u) {
      const uint32_t sb = b ^ rb.flip;
      const int par = (__popc(sb & rb.parS) + __popc(a & rb.parO)) & 1;
      v += (par ? -rb.s0 : rb.s0) * IN[ia * NB + __ldg(rankB + sb)];
    }
    return v;
  };
  int ss = 0, as = nS;
  for (int p = 0; p < n; ++p)
    for (int q = 0; q <= p; ++q, ++ss) {
      if (p == q) {
        D[(int64_t)ss * W + t] = wgt * elem(p, p);
      } else {
        const double x = elem(p, q), y = elem(q, p);
        D[(int64_t)ss * W + t] = wgt * (x + y);
        D[(int64_t)as * W + t] = wgt * (x - y);
        ++as;
      }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Spin-flip symmetric vectors, blocked panels: EVERY gather and EVERY atomic of the half build as a contiguous 256-byte run.
//
// The determinant-per-thread kernels above run lanes along the beta index: alpha partners (another row, same columns) are
// coalesced, beta partners (same row, scattered columns) cost one 32-byte sector per lane plus a rank look-up per lane, and those
// sector requests are what bounds them (L1 wavefronts).  For a spin-flip symmetric vector the beta partner can be read through its
// mirror, c[A, B'] = lambda phi(A, B') c[B', A]: row B', columns along A.  So a CTA takes a 32 x 32 block of determinants (rows
// ia0.., columns ib0.. of a block above / on the diagonal) and works on it in two thread mappings,
//     alpha mapping: a warp holds a row (string uniform), lanes along the columns  -> alpha partners c[A', B0 + lane]
//     beta mapping : a warp holds a column (string uniform), lanes along the rows  -> beta partners lambda phi c[B', A0 + lane]
// exchanging the per-determinant values between the mappings through a padded shared-memory tile (double-buffered: one barrier
// per generator pair).  Screens, parities and rank look-ups are warp-uniform.
// Scatter: with X = sum over the kept sources (weight 1/2 on the diagonal) of their alpha and beta contributions, H psi =
// X + lambda U X; the beta contributions are added at their MIRRORED targets (row B', columns along A: coalesced atomics), i.e. the
// kernel accumulates Y = X_alpha + lambda U X_beta, and sigma = Y + lambda U Y comes from one in-place symmetrisation pass
// (out starts as e_core / 2 * in).  Panel column t <-> (block t / 1024 of the panel, row (t % 1024) / 32, column t % 32); blocks in
// tri_unrank order over the (NA / 32)^2 block grid; lower-triangle determinants of diagonal blocks carry zero weight.
// ---------------------------------------------------------------------------------------------------------------------------
#define BLK_TS (32 * 33)
#define BLK_SG 2   // generator pairs per barrier round: the loads of a round (one per determinant, spin and pair) are in flight together

// x with its sign flipped when bit 31 of `neg` is set (one logic operation on the high word instead of a negation and a select)
__device__ __forceinline__ double blk_flip_sign(double x, uint32_t neg) {
  return __hiloint2double(__double2hiint(x) ^ (int)(neg & 0x80000000u), __double2loint(x));
}

// For p != q at most one of E_pq, E_qp passes the screen of a string (p occupied and q empty, or the reverse): one record, one access.
struct BlkSel {
  uint32_t flip, parS;
  int s0;
  bool any, second;
};
template <bool SRC>
__device__ __forceinline__ BlkSel blk_select(const ERec& r1, const ERec& r2, bool pair, uint32_t s) {
  const bool v1 = SRC ? ((s & r1.occ) == r1.occ && (s & r1.emp) == 0u) : ((s & r1.tocc) == r1.tocc && (s & r1.temp) == 0u);
  const bool v2 = pair && (SRC ? ((s & r2.occ) == r2.occ && (s & r2.emp) == 0u) : ((s & r2.tocc) == r2.tocc && (s & r2.temp) == 0u));
  return {v1 ? r1.flip : r2.flip, v1 ? r1.parS : r2.parS, v1 ? r1.s0 : r2.s0, v1 || v2, v2};
}

// Everything that depends on (string, generator pair) only is worked out ONCE per CTA for its 32 row and 32 column strings:
//   act[(role * nSp + slot) * 32 + j], j = 4 * (r % 8) + r / 8 for string r of the block (role 0: row strings under E^alpha, role 1:
//   column strings under E^beta): -1 (all ones) if neither E_pq nor E_qp acts, else
//       rank of the partner string (24 bits) | E_qp taken << 29 | sign << 31
//   (sign = s0 and the same-spin parity of the SOURCE string; SRC: the string is the source, else the target of the operator);
//   info[slot] = {other-spin parity mask of E^alpha_pq (= that of E^alpha_qp, checked on the host), the same for E^beta, flip mask,
//   p << 8 | q}; slots padded to whole rounds (no action).
// The main loops are left with one uniform 16-byte table load per 4 determinants, one popcount and the memory access itself.
struct BlkSmem {
  int32_t* act;
  uint4* info;
  double* T;
};
template <bool SRC>
__device__ __forceinline__ BlkSmem blk_stage_actions(const ERec* __restrict__ etab, int n, const uint32_t* __restrict__ str,
                                                     const int32_t* __restrict__ rank, int64_t N, int64_t ia0, int64_t ib0) {
  extern __shared__ uint4 blk_raw[];
  const int n2 = n * n, nS = n * (n + 1) / 2, nSp = (nS + BLK_SG - 1) / BLK_SG * BLK_SG;
  BlkSmem S;
  S.act = reinterpret_cast<int32_t*>(blk_raw);
  S.info = blk_raw + 2 * nSp * 8;
  S.T = reinterpret_cast<double*>(S.info + nSp);
  ERec* const et = reinterpret_cast<ERec*>(S.T);   // the records, staged in the tile area while the tables are built
  for (int w = threadIdx.x; w < 2 * n2 * (int)(sizeof(ERec) / 4); w += 256) reinterpret_cast<uint32_t*>(et)[w] = reinterpret_cast<const uint32_t*>(etab)[w];
  __syncthreads();
  for (int s = threadIdx.x; s < nSp; s += 256) {
    if (s >= nS) {
      S.info[s] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    int p = (int)((sqrtf(8.0f * (float)s + 1.0f) - 1.0f) * 0.5f);
    while (p * (p + 1) / 2 > s) --p;
    while ((p + 1) * (p + 2) / 2 <= s) ++p;
    const int q = s - p * (p + 1) / 2;
    S.info[s] = make_uint4(et[2 * (p * n + q)].parO, et[2 * (p * n + q) + 1].parO, et[2 * (p * n + q) + 1].flip, (uint32_t)(p << 8 | q));
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * nSp * 32; idx += 256) {
    const int role = idx >= nSp * 32, rem = idx - role * nSp * 32, s = rem >> 5, j = rem & 31, r = (j >> 2) + 8 * (j & 3);
    if (s >= nS) {
      S.act[idx] = -1;
      continue;
    }
    const uint32_t pq = S.info[s].w;
    const int p = (int)(pq >> 8), q = (int)(pq & 255u);
    const int64_t gi = (role ? ib0 : ia0) + r;
    const uint32_t sstr = gi < N ? __ldg(str + gi) : 0u;
    const BlkSel e = blk_select<SRC>(et[2 * (p * n + q) + role], et[2 * (q * n + p) + role], p != q, sstr);
    int32_t word = -1;
    if (e.any) {
      const uint32_t partner = sstr ^ e.flip;
      const uint32_t neg = (__popc((SRC ? sstr : partner) & e.parS) & 1) ^ (e.s0 < 0 ? 1u : 0u);
      word = (int32_t)((uint32_t)__ldg(rank + partner) | (e.second ? 1u << 29 : 0u) | neg << 31);
    }
    S.act[idx] = word;
  }
  __syncthreads();
  return S;
}

// N8: bytes per row of the vector (N < 2^24), a kernel parameter so that every address is one 32 x 32 -> 64-bit multiply-add.
// Lanes beyond the last row / column of an edge block address row / column N - 1 instead: their values carry weight zero.
template <bool DSA>
__global__ void __launch_bounds__(256, 4)
build_blk_kernel(const double* __restrict__ IN, double* __restrict__ D, int64_t W, int64_t k0, int64_t nblk, int64_t nbg,
                 const ERec* __restrict__ etab, int n, const uint32_t* __restrict__ str, const int32_t* __restrict__ rank, int64_t N,
                 uint32_t N8, double lambda) {
  const int n2 = n * n, nS = n * (n + 1) / 2;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int64_t kb = k0 + blockIdx.x, t0 = (int64_t)blockIdx.x * 1024;
  if (kb >= nblk) {   // beyond the last block: zero columns
    const int rows = DSA ? n2 : nS;
    for (int slot = 0; slot < rows; ++slot)
#pragma unroll
      for (int i = 0; i < 4; ++i) D[(int64_t)slot * W + t0 + (wp + 8 * i) * 32 + lane] = 0.0;
    return;
  }
  int64_t bi, bj;
  tri_unrank(kb, nbg, &bi, &bj);
  const int64_t ia0 = bi * 32, ib0 = bj * 32;
  const BlkSmem S = blk_stage_actions<false>(etab, n, str, rank, N, ia0, ib0);
  // alpha mapping: rows wp + 8 i (uniform strings), column lane (string bL)
  // beta mapping : columns wp + 8 i (uniform strings bC), row lane (string aL)
  uint32_t bC[4];
  double wgt[4];
  const int64_t ibL = min(ib0 + lane, N - 1), iaL = min(ia0 + lane, N - 1);
  const uint32_t bL = __ldg(str + ibL), aL = __ldg(str + iaL);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t ia = ia0 + wp + 8 * i, ib = ib0 + wp + 8 * i;
    bC[i] = ib < N ? __ldg(str + ib) : 0u;
    wgt[i] = (ia < N && ib0 + lane < N && ia <= ib0 + lane) ? ((DSA && ia != ib0 + lane) ? 1.4142135623730951 : 1.0) : 0.0;
  }
  const char* const INa = reinterpret_cast<const char*>(IN + iaL);   // beta partners through their mirrors: row of the partner, column iaL
  const char* const INb = reinterpret_cast<const char*>(IN + ibL);   // alpha partners: row of the partner, column ibL
  const int nSp = (nS + BLK_SG - 1) / BLK_SG * BLK_SG;
  const int4* const actA = reinterpret_cast<const int4*>(S.act) + wp;
  const int4* const actB = actA + nSp * 8;
  const int32_t* const actL = S.act + nSp * 32 + (((lane & 7) << 2) | (lane >> 3));   // table entries of column string `lane`
  double* const dst = D + t0 + wp * 32 + lane;
  double* const Tw = S.T + wp * 33 + lane;        // beta mapping writes [column wp + 8 i][row lane]
  const double* const Tr = S.T + lane * 33 + wp;  // alpha mapping reads [column lane][row wp + 8 i]
  for (int s0 = 0; s0 < nS; s0 += BLK_SG) {
    double xb[BLK_SG][4], xa[BLK_SG][4];
    uint32_t sec[BLK_SG];   // DSA: bit i: the row string of determinant i takes E_qp; bit 4: the column string of this lane does
#pragma unroll
    for (int s = 0; s < BLK_SG; ++s) {
      const uint4 inf = S.info[s0 + s];
      const int4 wa4 = actA[(s0 + s) * 8], wb4 = actB[(s0 + s) * 8];
      const int wa[4] = {wa4.x, wa4.y, wa4.z, wa4.w}, wb[4] = {wb4.x, wb4.y, wb4.z, wb4.w};
      const uint32_t ma = bL & inf.x;                 // alpha operators: other-spin parity on the column string
      const uint32_t pa = (uint32_t)__popc(ma) << 31;
      sec[s] = 0u;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        {   // <J|E^beta|in> through the mirror of the beta partner: lambda phi(a, b') in[b', a]
          const int w = wb[i];
          const uint32_t neg = (uint32_t)w ^ ((uint32_t)__popc(aL & (inf.y ^ bC[i] ^ inf.z)) << 31);
          const double v = (w != -1) ? *reinterpret_cast<const double*>(INa + (uint64_t)((uint32_t)w & 0xffffffu) * N8) : 0.0;
          xb[s][i] = blk_flip_sign(v, neg);
        }
        {   // <J|E^alpha|in>
          const int w = wa[i];
          const double v = (w != -1) ? *reinterpret_cast<const double*>(INb + (uint64_t)((uint32_t)w & 0xffffffu) * N8) : 0.0;
          xa[s][i] = blk_flip_sign(v, (uint32_t)w ^ pa);
          if (DSA) sec[s] |= (uint32_t)((w >> 29) & 1) << i;
        }
      }
      if (DSA) sec[s] |= (uint32_t)((actL[(s0 + s) * 32] >> 29) & 1) << 4;
    }
#pragma unroll
    for (int s = 0; s < BLK_SG; ++s)
#pragma unroll
      for (int i = 0; i < 4; ++i) Tw[s * BLK_TS + i * (8 * 33)] = lambda * xb[s][i];
    __syncthreads();
#pragma unroll
    for (int s = 0; s < BLK_SG; ++s) {
      if (s0 + s >= nS) continue;
      const uint32_t pq = S.info[s0 + s].w;
      const int p = (int)(pq >> 8), q = (int)(pq & 255u);
      const int64_t ss = s0 + s, as = nS + p * (p - 1) / 2 + q;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double tb = Tr[s * BLK_TS + i * 8];
        dst[ss * W + i * 256] = wgt[i] * (xa[s][i] + tb);
        if (DSA && p != q) dst[as * W + i * 256] = wgt[i] * (blk_flip_sign(xa[s][i], sec[s] << (31 - i)) + blk_flip_sign(tb, sec[s] << 27));
      }
    }
    __syncthreads();
  }
}

// Y += contributions of the kept sources of one panel (see above); F has the n (n + 1) / 2 symmetrised rows, k_pq = k_qp
__global__ void __launch_bounds__(256, 4)
scatter_blk_kernel(const double* __restrict__ IN, double* __restrict__ OUT, const double* __restrict__ F,
                   const double* __restrict__ kmat, int64_t W, int64_t k0, int64_t nblk, int64_t nbg, const ERec* __restrict__ etab, int n,
                   const uint32_t* __restrict__ str, const int32_t* __restrict__ rank, int64_t N, uint32_t N8, double lambda) {
  const int nS = n * (n + 1) / 2;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int64_t kb = k0 + blockIdx.x, t0 = (int64_t)blockIdx.x * 1024;
  if (kb >= nblk) return;
  int64_t bi, bj;
  tri_unrank(kb, nbg, &bi, &bj);
  const int64_t ia0 = bi * 32, ib0 = bj * 32;
  const BlkSmem S = blk_stage_actions<true>(etab, n, str, rank, N, ia0, ib0);
  uint32_t bC[4];
  double wJ[4], cjw[4];
  const int64_t ibL = min(ib0 + lane, N - 1), iaL = min(ia0 + lane, N - 1);
  const uint32_t bL = __ldg(str + ibL), aL = __ldg(str + iaL);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t ia = ia0 + wp + 8 * i, ib = ib0 + wp + 8 * i;
    bC[i] = ib < N ? __ldg(str + ib) : 0u;
    const bool kept = ia < N && ib0 + lane < N && ia <= ib0 + lane;
    wJ[i] = kept ? (ia == ib0 + lane ? 0.5 : 1.0) : 0.0;   // masked determinants (lower triangle of a diagonal block, beyond N) carry value 0 everywhere below
    cjw[i] = kept ? wJ[i] * IN[ia * N + ibL] : 0.0;
  }
  char* const OUTa = reinterpret_cast<char*>(OUT + iaL);   // mirrored beta targets: row of the target string, column iaL
  char* const OUTb = reinterpret_cast<char*>(OUT + ibL);   // alpha targets
  const int nSp = (nS + BLK_SG - 1) / BLK_SG * BLK_SG;
  const double* const Fb = F + t0 + wp * 32 + lane;
  const int4* const actA = reinterpret_cast<const int4*>(S.act) + wp;
  const int4* const actB = actA + nSp * 8;
  double* const Tw = S.T + wp * 33 + lane;        // alpha mapping writes [row wp + 8 i][column lane]
  const double* const Tr = S.T + lane * 33 + wp;  // beta mapping reads [row lane][column wp + 8 i]
  for (int s0 = 0; s0 < nS; s0 += BLK_SG) {
    double val[BLK_SG][4];
#pragma unroll
    for (int s = 0; s < BLK_SG; ++s)
#pragma unroll
      for (int i = 0; i < 4; ++i) val[s][i] = (s0 + s < nS && wJ[i] != 0.0) ? Fb[(int64_t)(s0 + s) * W + i * 256] : 0.0;
    // ---- alpha mapping: value of every source, alpha contributions at their targets
#pragma unroll
    for (int s = 0; s < BLK_SG; ++s) {
      const uint4 inf = S.info[s0 + s];
      const double kk = __ldg(kmat + (inf.w >> 8) * n + (inf.w & 255u));   // a padded slot reads k_00 and adds nothing (no action)
      const int4 wa4 = actA[(s0 + s) * 8];
      const int wa[4] = {wa4.x, wa4.y, wa4.z, wa4.w};
      const uint32_t pa = (uint32_t)__popc(bL & inf.x) << 31;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double v = wJ[i] * val[s][i] + kk * cjw[i];
        Tw[s * BLK_TS + i * (8 * 33)] = v;
        const int w = wa[i];
        if (w != -1 && v != 0.0)
          atomicAdd(reinterpret_cast<double*>(OUTb + (uint64_t)((uint32_t)w & 0xffffffu) * N8), blk_flip_sign(v, (uint32_t)w ^ pa));
      }
    }
    __syncthreads();
    // ---- beta mapping: beta contributions at the mirrors of their targets
#pragma unroll
    for (int s = 0; s < BLK_SG; ++s) {
      const uint4 inf = S.info[s0 + s];
      const int4 wb4 = actB[(s0 + s) * 8];
      const int wb[4] = {wb4.x, wb4.y, wb4.z, wb4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double v = Tr[s * BLK_TS + i * 8];
        const int w = wb[i];
        if (w != -1 && v != 0.0) {
          const uint32_t neg = (uint32_t)w ^ ((uint32_t)__popc(aL & (inf.y ^ bC[i] ^ inf.z)) << 31);
          atomicAdd(reinterpret_cast<double*>(OUTa + (uint64_t)((uint32_t)w & 0xffffffu) * N8), lambda * blk_flip_sign(v, neg));
        }
      }
    }
    __syncthreads();
  }
}

// dynamic shared memory of the two kernels: action table, generator info, tiles (the staged records share the tile area)
static size_t blk_smem_bytes(int n) {
  const size_t nSp = ((size_t)n * (n + 1) / 2 + BLK_SG - 1) / BLK_SG * BLK_SG, tiles = sizeof(double) * BLK_SG * BLK_TS, recs = sizeof(ERec) * 2 * (size_t)n * n;
  return 2 * nSp * 32 * sizeof(int32_t) + nSp * sizeof(uint4) + std::max(tiles, recs);
}

// the other-spin parity mask of E_pq must equal that of E_qp (it counts the other spin's electrons between p and q) for the
// kernels above to use one mask per generator pair: checked on the host records, anything else keeps the determinant-per-thread route
static bool blk_tables_ok(const std::vector<ERec>& tab, int n) {
  for (int p = 0; p < n; ++p)
    for (int q = 0; q < p; ++q)
      for (int spin = 0; spin < 2; ++spin) {
        const ERec &a = tab[2 * ((size_t)p * n + q) + spin], &b = tab[2 * ((size_t)q * n + p) + spin];
        if (a.parO != b.parO || a.flip != b.flip) return false;
      }
  return true;
}

// OUT <- OUT + lambda U OUT in place: (U x)[A,B] = phi(A,B) x[B,A]; 32 x 32 tile pairs through shared memory
__global__ void __launch_bounds__(256)
spinsym_symmetrize_kernel(double* __restrict__ OUT, int64_t N, const uint32_t* __restrict__ str, double lambda) {
  __shared__ double tl[32][33], tu[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  for (int k = ty; k < 32; k += 8) {   // lower tile, along its rows
    const int64_t r = bj * 32 + k, c = bi * 32 + tx;
    tl[k][tx] = (r < N && c < N) ? OUT[r * N + c] : 0.0;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int64_t r = bi * 32 + k, c = bj * 32 + tx;
    double m = 0.0;
    if (r < N && c < N) {
      const bool neg = __popc(__ldg(str + r) & __ldg(str + c)) & 1;
      const double y = tl[tx][k];
      const double v = OUT[r * N + c] + lambda * (neg ? -y : y);
      OUT[r * N + c] = v;
      m = lambda * (neg ? -v : v);
    }
    tu[k][tx] = m;
  }
  __syncthreads();
  if (bj > bi)
    for (int k = ty; k < 32; k += 8) {
      const int64_t r = bj * 32 + k, c = bi * 32 + tx;
      if (r < N && c < N) OUT[r * N + c] = tu[tx][k];
    }
}

static int g_spinsym_blk = 1;   // sq_set_option("sigma_spinsym", "tri"): the determinant-per-thread kernels of the half build
void sq_hamiltonian_set_spinsym_blk(int on) { g_spinsym_blk = on ? 1 : 0; }

// measures the spin-flip symmetry of a vector (see above): *lambda = +-1 if c[B,A] = lambda phi(A,B) c[A,B] to 1e-12 max|c|, else 0
static int spinsym_measure(sq_space* sp, HamWork* w, const double* vec, cudaStream_t st, double* lambda) {
  *lambda = 0.0;
  if (sp->n_alpha != sp->n_beta || sp->NA != sp->NB || sp->NA < 2) return SQ_OK;
  if (!w->d_symres) SQ_CUDA(cudaMalloc(&w->d_symres, 3 * sizeof(unsigned long long)));
  SQ_CUDA(cudaMemsetAsync(w->d_symres, 0, 3 * sizeof(unsigned long long), st));
  const unsigned nb32 = (unsigned)((sp->NA + 31) / 32);
  spinsym_check_kernel<<<dim3(nb32, nb32), 256, 0, st>>>(vec, sp->NA, sp->d_strA, w->d_symres);
  SQ_CHECK(launch_error("spinsym_check_kernel"));
  double res[3];
  SQ_CUDA(cudaMemcpyAsync(res, w->d_symres, sizeof(res), cudaMemcpyDeviceToHost, st));
  SQ_CUDA(cudaStreamSynchronize(st));
  const double tol_sym = 1e-12 * res[0];
  if (res[0] > 0.0 && res[1] <= tol_sym) *lambda = 1.0;
  else if (res[0] > 0.0 && res[2] <= tol_sym) *lambda = -1.0;
  return SQ_OK;
}

static int launch_scatter_E(sq_space* sp, HamWork* w, const double* in, double* out, const double* F, const double* d_k,
                            int64_t j0, cudaStream_t st, bool use_const) {
  const int n2 = sp->n_orb * sp->n_orb;
  if (g_etab_tab && w->d_psA) {   // per-string partner tables
    scatter_E_tab_kernel<<<(unsigned)(w->W / 256), 256, 0, st>>>(in, out, F, d_k, w->d_frow, w->W, j0, sp->local_len(), w->d_psA, w->d_psB,
                                                                 w->d_parO, sp->n_orb, sp->d_strA, sp->d_strB, sp->NB);
    return launch_error("scatter_E_tab_kernel");
  }
  if (g_etab_alu) {   // table-free records (candidate, off by default)
    scatter_E_alu_kernel<<<(unsigned)(w->W / 256), 256, 0, st>>>(in, out, F, d_k, w->d_frow, w->W, j0, sp->local_len(), sp->n_orb,
                                                                 sp->d_strA, sp->d_strB, sp->d_rankA, sp->d_rankB, sp->NB, sp->row_begin);
    return launch_error("scatter_E_alu_kernel");
  }
  if (g_rows_kernels && w->d_tabS) {
    const RowGrid g = row_grid(sp, w, j0);
    const size_t smem = rows_smem(sp);
    allow_smem(scatter_E_rows_kernel, smem);
    scatter_E_rows_kernel<<<g.grid, g_rows_threads, smem, st>>>(in, out, F, d_k, w->d_frow, w->W, j0, g.jend, w->d_etab, n2, sp->d_strA,
                                                     sp->d_rankA, w->d_tabS, sp->NB, g.ia_first, g.CH);
    return launch_error("scatter_E_rows_kernel");
  }
  const unsigned grid = (unsigned)(w->W / 256);
  if (use_const) {
    scatter_E_kernel<true><<<grid, 256, 0, st>>>(in, out, F, d_k, w->d_frow, w->W, j0, sp->local_len(), w->d_etab, n2, sp->d_strA,
                                                 sp->d_strB, sp->d_rankA, sp->d_rankB, sp->NB, sp->row_begin);
  } else {
    const size_t smem = sizeof(ERec) * 2 * (size_t)n2;
    allow_smem(scatter_E_kernel<false>, smem);
    scatter_E_kernel<false><<<grid, 256, smem, st>>>(in, out, F, d_k, w->d_frow, w->W, j0, sp->local_len(), w->d_etab, n2,
                                                     sp->d_strA, sp->d_strB, sp->d_rankA, sp->d_rankB, sp->NB, sp->row_begin);
  }
  return launch_error("scatter_E_kernel");
}

extern "C" int sq_sigma(sq_space* sp, double e_core, const double* h_act_host, const double* g_act_host,
                        const double* in_dev, double* out_dev, void* stream) {
  SqRange nvtx_range("sq_sigma");
  if (!sp || !h_act_host || !g_act_host || !in_dev || !out_dev) return SQ_ERR_INVALID;
  if (in_dev == out_dev) {
    sq_set_error("sq_sigma: in and out must not alias");
    return SQ_ERR_INVALID;
  }
  SQ_CHECK(check_full_space(sp, "sq_sigma"));
  cudaStream_t st = (cudaStream_t)stream;
  SQ_CUDA(cudaSetDevice(sp->device));
  HamWork* w = nullptr;
  SQ_CHECK(get_work(sp, false, true, &w));
  const int n = sp->n_orb, n2 = n * n;
  // Gm[pq][rs] = 1/2 g_pqrs ; k_pq = h_pq - 1/2 sum_r g_prrq   (from e_pqrs = E_pq E_rs - delta_qr E_ps)
  std::vector<double> k((size_t)n2);
  double gmax = 0.0;
  for (size_t i = 0; i < (size_t)n2 * n2; ++i) gmax = std::max(gmax, std::fabs(g_act_host[i]));
  for (int p = 0; p < n; ++p)
    for (int q = 0; q < n; ++q) {
      double v = h_act_host[p * n + q];
      for (int r = 0; r < n; ++r) v -= 0.5 * g_act_host[(((size_t)p * n + r) * n + r) * n + q];
      k[(size_t)p * n + q] = v;
    }
  // Real-orbital integrals have g_pqrs = g_qprs = g_pqsr (and then k_pq = k_qp): the panel shrinks to the
  // n (n + 1) / 2 symmetrised generators E_rs + E_sr -- 3.5 x fewer DGEMM flops at n = 16.  Anything else takes the
  // general n^2 path (no symmetry assumed).
  bool sym = true;
  const double tol = 1e-13 * (gmax > 0 ? gmax : 1.0);
  auto G = [&](int p, int q, int r, int t) { return g_act_host[(((size_t)p * n + q) * n + r) * n + t]; };
  for (int p = 0; p < n && sym; ++p)
    for (int q = 0; q < n && sym; ++q) {
      if (std::fabs(k[(size_t)p * n + q] - k[(size_t)q * n + p]) > 1e-13 * (1.0 + std::fabs(k[(size_t)p * n + q]))) sym = false;
      for (int r = 0; r < n && sym; ++r)
        for (int t = 0; t < n; ++t)
          if (std::fabs(G(p, q, r, t) - G(q, p, r, t)) > tol || std::fabs(G(p, q, r, t) - G(p, q, t, r)) > tol) {
            sym = false;
            break;
          }
    }
  const int nS = n * (n + 1) / 2;
  const int nrow = sym ? nS : n2;   // rows of the D and F panels
  const int ldg = (nrow + 1) & ~1;  // even leading dimension of the integral matrix (16-byte rows for the async copies)
  std::vector<double> Gm((size_t)nrow * ldg, 0.0);
  std::vector<int> frow((size_t)n2);
  if (sym) {
    auto slot = [](int r, int t) { return r >= t ? r * (r + 1) / 2 + t : t * (t + 1) / 2 + r; };
    for (int p = 0; p < n; ++p)
      for (int q = 0; q <= p; ++q)
        for (int r = 0; r < n; ++r)
          for (int t = 0; t <= r; ++t) Gm[(size_t)slot(p, q) * ldg + slot(r, t)] = 0.5 * G(p, q, r, t);
    for (int p = 0; p < n; ++p)
      for (int q = 0; q < n; ++q) frow[(size_t)p * n + q] = slot(p, q);
  } else {
    for (int a = 0; a < n2; ++a)
      for (int b = 0; b < n2; ++b) Gm[(size_t)a * ldg + b] = 0.5 * g_act_host[(size_t)a * n2 + b];
    for (int i = 0; i < n2; ++i) frow[i] = i;
  }
  double* d_G = w->d_small;
  double* d_k = w->d_small + (size_t)n2 * (n2 + 1);
  SQ_CUDA(cudaMemcpyAsync(d_G, Gm.data(), sizeof(double) * Gm.size(), cudaMemcpyHostToDevice, st));
  SQ_CUDA(cudaMemcpyAsync(d_k, k.data(), sizeof(double) * k.size(), cudaMemcpyHostToDevice, st));
  SQ_CUDA(cudaMemcpyAsync(w->d_frow, frow.data(), sizeof(int) * frow.size(), cudaMemcpyHostToDevice, st));
  SQ_CUDA(cudaStreamSynchronize(st));   // host vectors go out of scope below
  const int64_t len = sp->local_len();
  bool use_const = false;
  SQ_CHECK(bind_etab(sp, w, st, &use_const));
  // Spin-flip symmetric input (see "half of the sigma build" above): measured, not assumed
  bool tri = false;
  double lambda = 1.0;
  if (g_sigma_spinsym && !g_sigma_fused && sym && !use_const && !g_etab_alu && !g_etab_tab && !(g_rows_kernels && w->d_tabG) &&
      sp->n_alpha == sp->n_beta && sp->NA == sp->NB && sp->NA > 1) {
    SQ_CHECK(spinsym_measure(sp, w, in_dev, st, &lambda));
    tri = lambda != 0.0;
  }
  // blocked panels (32 x 32 blocks of determinants, every access a contiguous run) when a panel holds at least one block
  const bool blk = tri && g_spinsym_blk && w->W >= 1024 && sp->NA < ((int64_t)1 << 24) && n < 256 && blk_smem_bytes(n) <= 200 * 1024 && blk_tables_ok(w->h_etab, n);
  const int64_t nbg = (sp->NA + 31) / 32, nblk = nbg * (nbg + 1) / 2, bpp = w->W / 1024;   // block grid, kept blocks, blocks per panel
  SQ_CHECK(sq_launch_scale_copy(sp, blk ? 0.5 * e_core : e_core, in_dev, out_dev, st));
  if (g_sigma_fused) {   // one fused gather -> DMMA -> scatter kernel, no D / F panels in HBM
    const int rc = launch_sigma_fused(sp, in_dev, out_dev, d_G, ldg, nrow, sym, d_k, w->d_frow, w->n_sm, st);
    if (rc != SQ_ERR_UNSUPPORTED) return rc;
  }
  const int64_t len_eff = blk ? ((nblk + bpp - 1) / bpp) * w->W : (tri ? sp->NA * (sp->NA + 1) / 2 : len);
  const size_t blk_smem = blk_smem_bytes(n);
  if (blk) {
    allow_smem(build_blk_kernel<false>, blk_smem);
    allow_smem(scatter_blk_kernel, blk_smem);
  }
  auto build_panel = [&](double* Dp, int64_t j0, cudaStream_t s) -> int {
    if (blk) {
      if (w->W > bpp * 1024)   // columns behind the last block of the panel
        SQ_CUDA(cudaMemset2DAsync(Dp + bpp * 1024, sizeof(double) * (size_t)w->W, 0, sizeof(double) * (size_t)(w->W - bpp * 1024), (size_t)nrow, s));
      build_blk_kernel<false><<<(unsigned)bpp, 256, blk_smem, s>>>(in_dev, Dp, w->W, (j0 / w->W) * bpp, nblk, nbg, w->d_etab, n, sp->d_strA,
                                                                  sp->d_rankA, sp->NA, (uint32_t)(sp->NA * 8), lambda);
      return launch_error("build_blk_kernel");
    }
    if (!tri) return launch_build_D(sp, w, in_dev, Dp, j0, s, use_const, sym);
    const size_t smem = sizeof(ERec) * 2 * (size_t)n2;
    allow_smem(build_Dsym_tri_kernel, smem);
    build_Dsym_tri_kernel<<<(unsigned)(w->W / 256), 256, smem, s>>>(in_dev, Dp, w->W, j0, len_eff, w->d_etab, n, sp->d_strA, sp->d_strB,
                                                                     sp->d_rankA, sp->d_rankB, sp->NB);
    return launch_error("build_Dsym_tri_kernel");
  };
  auto scatter_panel = [&](const double* Fp, int64_t j0, cudaStream_t s) -> int {
    if (blk) {
      scatter_blk_kernel<<<(unsigned)bpp, 256, blk_smem, s>>>(in_dev, out_dev, Fp, d_k, w->W, (j0 / w->W) * bpp, nblk, nbg, w->d_etab, n,
                                                             sp->d_strA, sp->d_rankA, sp->NA, (uint32_t)(sp->NA * 8), lambda);
      return launch_error("scatter_blk_kernel");
    }
    if (!tri) return launch_scatter_E(sp, w, in_dev, out_dev, Fp, d_k, j0, s, use_const);
    const size_t smem = sizeof(ERec) * 2 * (size_t)n2;
    allow_smem(scatter_E_tri_kernel, smem);
    scatter_E_tri_kernel<<<(unsigned)(w->W / 256), 256, smem, s>>>(in_dev, out_dev, Fp, d_k, w->d_frow, w->W, j0, len_eff, w->d_etab, n2,
                                                                    sp->d_strA, sp->d_strB, sp->d_rankA, sp->d_rankB, sp->NB, lambda);
    return launch_error("scatter_E_tri_kernel");
  };
  // Three-stage pipeline over the panels: while the DGEMM of panel k runs on the tensor cores, the gather of panel k+1
  // and the scatter of panel k-1 (both address-bound) run beside it.  Two D and two F panels are in flight; with
  // pipeline off (or a single panel) all three stages are issued on the caller's stream.
  const bool piped = g_panel_pipeline && w->d_D[1] && w->d_F[1];
  cudaStream_t s_build = piped ? w->s_build : st, s_gemm = piped ? w->s_gemm : st, s_scat = piped ? w->s_scat : st;
  if (piped) {
    SQ_CUDA(cudaEventRecord(w->ev_start, st));
    SQ_CUDA(cudaStreamWaitEvent(s_build, w->ev_start, 0));
    SQ_CUDA(cudaStreamWaitEvent(s_gemm, w->ev_start, 0));
    SQ_CUDA(cudaStreamWaitEvent(s_scat, w->ev_start, 0));
  }
  int64_t ip = 0;
  for (int64_t j0 = 0; j0 < len_eff; j0 += w->W, ++ip) {
    const int b = piped ? (int)(ip & 1) : 0;
    double* Dp = w->d_D[b];
    double* Fp = w->d_F[b];
    if (piped && ip >= 2) SQ_CUDA(cudaStreamWaitEvent(s_build, w->ev_gemm[b], 0));   // D[b] is free once GEMM k-2 has read it
    SQ_CHECK(build_panel(Dp, j0, s_build));
    if (piped) {
      SQ_CUDA(cudaEventRecord(w->ev_built[b], s_build));
      SQ_CUDA(cudaStreamWaitEvent(s_gemm, w->ev_built[b], 0));
      if (ip >= 2) SQ_CUDA(cudaStreamWaitEvent(s_gemm, w->ev_scat[b], 0));             // F[b] is free once scatter k-2 has read it
    }
    // F[pq][t] = sum_rs Gm[pq][rs] D[rs][t]: hand-written DMMA kernel (sqsv_dmma.cu)
    SQ_CHECK(sq_sigma_gemm(d_G, ldg, Dp, Fp, nrow, w->W, s_gemm));
    if (piped) {
      SQ_CUDA(cudaEventRecord(w->ev_gemm[b], s_gemm));
      SQ_CUDA(cudaStreamWaitEvent(s_scat, w->ev_gemm[b], 0));
    }
    SQ_CHECK(scatter_panel(Fp, j0, s_scat));
    if (piped) SQ_CUDA(cudaEventRecord(w->ev_scat[b], s_scat));
  }
  if (piped) {   // the caller's stream continues after the last scatter (which is after everything else)
    SQ_CUDA(cudaEventRecord(w->ev_start, s_scat));
    SQ_CUDA(cudaStreamWaitEvent(st, w->ev_start, 0));
    SQ_CUDA(cudaEventRecord(w->ev_start, s_build));
    SQ_CUDA(cudaStreamWaitEvent(st, w->ev_start, 0));
  }
  if (blk) {   // sigma = Y + lambda U Y
    spinsym_symmetrize_kernel<<<dim3((unsigned)nbg, (unsigned)nbg), 256, 0, st>>>(out_dev, sp->NA, sp->d_strA, lambda);
    SQ_CHECK(launch_error("spinsym_symmetrize_kernel"));
  } else if (tri) {   // the lower triangle from the upper one
    const unsigned nb32 = (unsigned)((sp->NA + 31) / 32);
    spinsym_mirror_kernel<<<dim3(nb32, nb32), 256, 0, st>>>(out_dev, sp->NA, sp->d_strA, lambda);
    SQ_CHECK(launch_error("spinsym_mirror_kernel"));
  }
  return SQ_OK;
}

static int rdm12_impl(sq_space* sp, const double* bra_dev, const double* ket_dev, const PeerView* pv_bra,
                      const PeerView* pv_ket, double* rdm1_host, double* rdm2_host, void* stream);
static int launch_build_DSA_peer(sq_space* sp, const PeerView* pv, const double* in, double* D, int64_t W, int64_t j0, bool half,
                                 cudaStream_t st);
static int64_t half_len_host(const sq_space* sp);
static thread_local double t_rdm_dist_lambda = 0.0;   // set by sq_rdm12_dist_sym around its call of rdm12_impl

extern "C" int sq_rdm12(sq_space* sp, const double* bra_dev, const double* ket_dev, double* rdm1_host,
                        double* rdm2_host, void* stream) {
  if (!sp || !bra_dev || !ket_dev || !rdm1_host) return SQ_ERR_INVALID;
  SQ_CHECK(check_full_space(sp, "sq_rdm12"));
  return rdm12_impl(sp, bra_dev, ket_dev, nullptr, nullptr, rdm1_host, rdm2_host, stream);
}

// This rank's PARTIAL sums over its rows of an alpha-sharded vector (the caller adds the ranks' results: everything is
// linear in the partial sums).  *_ptrs_host[r] = base pointer of rank r's shard as mapped into this process.
// sq_rdm12_dist for a spin-flip symmetric vector (bra == ket; lambda = +-1 measured by sq_spinsym_measure_dist + MAX all-reduce):
// the S / A panels of the kept half of this rank's rows, weighted.  lambda = 0: same as sq_rdm12_dist.
extern "C" int sq_rdm12_dist(sq_space* sp, const double* const* bra_ptrs_host, const double* const* ket_ptrs_host,
                             double* rdm1_host, double* rdm2_host, void* stream);
extern "C" int sq_rdm12_dist_sym(sq_space* sp, const double* const* bra_ptrs_host, const double* const* ket_ptrs_host, double lambda,
                                 double* rdm1_host, double* rdm2_host, void* stream) {
  if (lambda != 0.0 && lambda != 1.0 && lambda != -1.0) return SQ_ERR_INVALID;
  t_rdm_dist_lambda = lambda;
  const int rc = sq_rdm12_dist(sp, bra_ptrs_host, ket_ptrs_host, rdm1_host, rdm2_host, stream);
  t_rdm_dist_lambda = 0.0;
  return rc;
}

extern "C" int sq_rdm12_dist(sq_space* sp, const double* const* bra_ptrs_host, const double* const* ket_ptrs_host,
                             double* rdm1_host, double* rdm2_host, void* stream) {
  if (!sp || !bra_ptrs_host || !ket_ptrs_host || !rdm1_host) return SQ_ERR_INVALID;
  if (sp->device < 0) return SQ_ERR_INVALID;
  if (sp->world < 1 || sp->world > SQ_MAX_WORLD || (int)sp->row_starts.size() != sp->world + 1) {
    sq_set_error("sq_rdm12_dist: the space has no row partition (sq_space_set_partition)");
    return SQ_ERR_INVALID;
  }
  PeerView pb, pk;
  pb.world = pk.world = sp->world;
  for (int r = 0; r < SQ_MAX_WORLD; ++r) {
    pb.p[r] = r < sp->world ? bra_ptrs_host[r] : nullptr;
    pk.p[r] = r < sp->world ? ket_ptrs_host[r] : nullptr;
  }
  for (int r = 0; r <= SQ_MAX_WORLD; ++r) pb.row_starts[r] = pk.row_starts[r] = r <= sp->world ? sp->row_starts[r] : sp->row_starts[sp->world];
  const int n = sp->n_orb, n2 = n * n;
  if (sp->local_len() == 0) {   // a rank without rows contributes nothing
    for (int i = 0; i < n2; ++i) rdm1_host[i] = 0.0;
    if (rdm2_host) for (size_t i = 0; i < (size_t)n2 * n2; ++i) rdm2_host[i] = 0.0;
    return SQ_OK;
  }
  return rdm12_impl(sp, pb.p[sp->rank], pk.p[sp->rank], &pb, &pk, rdm1_host, rdm2_host, stream);
}

// <E_pq E_rs> for all p, q, r, s and the 1-RDM from the two symmetric Gram matrices of the S / A generators (bra == ket, real):
//   E_pq = (S_pq + A_pq) / 2, E_qp = (S_pq - A_pq) / 2 (p > q), E_pp = S_pp;
//   <S_x S_y> = G_SS, <A_x A_y> = -G_AA (A is anti-Hermitian), <S_x A_y> = <[S_x, A_y]> / 2, <A_x S_y> = <[A_x, S_y]> / 2 with
//   <[E_pq, E_rs]> = delta_qr G1_ps - delta_ps G1_rq;  G1_pq = (1 / N) sum_r <E_pq E_rr> = c_pq / N sum_r G_SS[pq][rr].
// G2h[(q, p)][(r, s)] = <E_pq E_rs> (the layout the plain Gram path produces); ldg = leading dimension of GSS / GAA.
static void assemble_g2_from_sym(int n, const double* GSS, const double* GAA, int ldg, int n_elec, std::vector<double>* G2h,
                                 std::vector<double>* g1h) {
  const int n2 = n * n;
  auto sslot = [](int p, int q) { return p * (p + 1) / 2 + q; };   // p >= q
  auto aslot = [](int p, int q) { return p * (p - 1) / 2 + q; };   // p > q
  std::vector<double>& g1 = *g1h;
  g1.assign((size_t)n2, 0.0);
  for (int p = 0; p < n; ++p)
    for (int q = 0; q <= p; ++q) {
      double v = 0.0;
      for (int r = 0; r < n; ++r) v += GSS[(size_t)sslot(p, q) * ldg + sslot(r, r)];
      v *= (p == q ? 1.0 : 0.5) / n_elec;
      g1[(size_t)p * n + q] = g1[(size_t)q * n + p] = v;
    }
  auto C4 = [&](int p, int q, int r, int s) { return (q == r ? g1[(size_t)p * n + s] : 0.0) - (p == s ? g1[(size_t)r * n + q] : 0.0); };
  struct Term { double c; int kind, slot, p, q; };   // kind 0: S, 1: A; (p, q) with p >= q
  auto expand = [&](int p, int q, Term* out) -> int {   // E_pq in terms of S / A generators
    if (p == q) { out[0] = {1.0, 0, sslot(p, p), p, p}; return 1; }
    const int hi = p > q ? p : q, lo = p > q ? q : p;
    out[0] = {0.5, 0, sslot(hi, lo), hi, lo};
    out[1] = {p > q ? 0.5 : -0.5, 1, aslot(hi, lo), hi, lo};
    return 2;
  };
  // <[Z_x, Z_y]> with Z = S or A of the pair (p >= q): Z = E_pq + eps E_qp (eps = +1 S, -1 A; S_pp = E_pp)
  auto comm = [&](const Term& x, const Term& y) {
    double v = 0.0;
    const int nx = x.p == x.q ? 1 : 2, ny = y.p == y.q ? 1 : 2;
    for (int i = 0; i < nx; ++i)
      for (int j = 0; j < ny; ++j) {
        const double ci = (i == 0 ? 1.0 : (x.kind ? -1.0 : 1.0)), cj = (j == 0 ? 1.0 : (y.kind ? -1.0 : 1.0));
        const int p = i == 0 ? x.p : x.q, q = i == 0 ? x.q : x.p, r = j == 0 ? y.p : y.q, s = j == 0 ? y.q : y.p;
        v += ci * cj * C4(p, q, r, s);
      }
    return v;
  };
  G2h->assign((size_t)n2 * n2, 0.0);
  Term ta[2], tb[2];
  for (int p = 0; p < n; ++p)
    for (int q = 0; q < n; ++q) {
      const int na = expand(p, q, ta);
      for (int r = 0; r < n; ++r)
        for (int s = 0; s < n; ++s) {
          const int nb = expand(r, s, tb);
          double v = 0.0;
          for (int i = 0; i < na; ++i)
            for (int j = 0; j < nb; ++j) {
              double m;
              if (ta[i].kind == 0 && tb[j].kind == 0) m = GSS[(size_t)ta[i].slot * ldg + tb[j].slot];
              else if (ta[i].kind == 1 && tb[j].kind == 1) m = -GAA[(size_t)ta[i].slot * ldg + tb[j].slot];
              else m = 0.5 * comm(ta[i], tb[j]);
              v += ta[i].c * tb[j].c * m;
            }
          (*G2h)[((size_t)(q * n + p)) * n2 + (r * n + s)] = v;
        }
    }
}

static int g_rdm_sym = 1;   // sq_set_option("rdm_sym", "0"): the plain <E_pq E_rs> Gram matrix also for bra == ket
void sq_hamiltonian_set_rdm_sym(int on) { g_rdm_sym = on ? 1 : 0; }

static int rdm12_impl(sq_space* sp, const double* bra_dev, const double* ket_dev, const PeerView* pv_bra,
                      const PeerView* pv_ket, double* rdm1_host, double* rdm2_host, void* stream) {
  SqRange nvtx_range("sq_rdm12");
  cudaStream_t st = (cudaStream_t)stream;
  SQ_CUDA(cudaSetDevice(sp->device));
  const bool same = (bra_dev == ket_dev);
  HamWork* w = nullptr;
  SQ_CHECK(get_work(sp, !same && rdm2_host, false, &w));
  const int n = sp->n_orb, n2 = n * n;
  double* d_G2 = w->d_small;                       // [n2][n2] row-major [(q,p)][(r,s)]
  double* d_g1 = w->d_small + (size_t)n2 * n2;     // [n2]
  SQ_CUDA(cudaMemsetAsync(w->d_small, 0, sizeof(double) * ((size_t)n2 * n2 + 2 * (size_t)n2), st));
  bool use_const = false;
  SQ_CHECK(bind_etab(sp, w, st, &use_const));
  const int64_t len = sp->local_len();
  const int n_elec = sp->n_alpha + sp->n_beta;
  // With the 2-RDM accumulator at hand, rdm1 needs no pass of its own: sum_r E_rr = N on this space, so
  // <bra|E_pq|ket> = (1/N) sum_r <bra|E_pq E_rr|ket>  (saves one GEMV sweep over every panel).
  const bool rdm1_from_G2 = rdm2_host && n_elec > 0;
  GramTiles tiles;
  int n_split = 1;
  // bra == ket (real): two symmetric Gram matrices of the S / A generators instead of the n^2 x n^2 one (a quarter of the products)
  const int nS = n * (n + 1) / 2, nA = n2 - nS, GR = sq_gram_sym_rows();
  const bool sym_route = g_rdm_sym && same && rdm2_host && n_elec > 0 && nS <= GR;   // sharded vectors too (peer gathers)
  if (sym_route) {
    if (!w->d_gsym) SQ_CUDA(cudaMalloc(&w->d_gsym, sizeof(double) * 2 * (size_t)GR * GR));
    SQ_CHECK(sq_gram_sym_begin(w->n_sm, &w->d_gram, &w->gram_doubles, &n_split, st));
  } else if (rdm2_host) {
    SQ_CHECK(sq_gram_begin(n2, same, w->n_sm, &w->d_gram, &w->gram_doubles, &tiles, &n_split, st));
  }
  // spin-flip symmetric vector (measured): the S / A panels of the determinants above / on the diagonal only, weighted
  bool tri = false;
  double tri_lambda = 0.0;
  if (sym_route && g_sigma_spinsym && !pv_ket) {
    SQ_CHECK(spinsym_measure(sp, w, ket_dev, st, &tri_lambda));
    tri = tri_lambda != 0.0;
  }
  // sharded vector: the caller measured the symmetry over all ranks (sq_rdm12_dist_sym); the kept cyclic band of the local rows
  const bool half_band = sym_route && pv_ket && t_rdm_dist_lambda != 0.0 && sp->n_alpha == sp->n_beta && sp->NA == sp->NB;
  const bool blk = tri && g_spinsym_blk && w->W >= 1024 && sp->NA < ((int64_t)1 << 24) && n < 256 && blk_smem_bytes(n) <= 200 * 1024 && blk_tables_ok(w->h_etab, n);   // blocked panels (build_blk_kernel)
  const int64_t nbg = (sp->NA + 31) / 32, nblk = nbg * (nbg + 1) / 2, bpp = w->W / 1024;
  const size_t blk_smem = blk_smem_bytes(n);
  const int64_t len_eff = blk ? ((nblk + bpp - 1) / bpp) * w->W : (tri ? sp->NA * (sp->NA + 1) / 2 : (half_band ? half_len_host(sp) : len));
  if (blk) allow_smem(build_blk_kernel<true>, blk_smem);
  // Two-stage pipeline: the gather of panel k+1 runs beside the DGEMM of panel k (two panels per vector in flight).
  const bool piped = g_panel_pipeline && w->d_D[1] && (same || !rdm2_host || w->d_D[3]);
  cudaStream_t s_build = piped ? w->s_build : st, s_gemm = piped ? w->s_gemm : st;
  if (piped) {
    SQ_CUDA(cudaEventRecord(w->ev_start, st));
    SQ_CUDA(cudaStreamWaitEvent(s_build, w->ev_start, 0));
    SQ_CUDA(cudaStreamWaitEvent(s_gemm, w->ev_start, 0));
  }
  int64_t k = 0;
  for (int64_t j0 = 0; j0 < len_eff; j0 += w->W, ++k) {
    const int64_t wl = (len_eff - j0 < w->W) ? len_eff - j0 : w->W;
    const int b = piped ? (int)(k & 1) : 0;
    double* Dket = w->d_D[b];
    if (piped && k >= 2) SQ_CUDA(cudaStreamWaitEvent(s_build, w->ev_gemm[b], 0));   // panels b are free once GEMM k-2 is done
    if (blk) {
      if (w->W > bpp * 1024)
        SQ_CUDA(cudaMemset2DAsync(Dket + bpp * 1024, sizeof(double) * (size_t)w->W, 0, sizeof(double) * (size_t)(w->W - bpp * 1024), (size_t)n2, s_build));
      build_blk_kernel<true><<<(unsigned)bpp, 256, blk_smem, s_build>>>(ket_dev, Dket, w->W, k * bpp, nblk, nbg, w->d_etab, n, sp->d_strA,
                                                                       sp->d_rankA, sp->NA, (uint32_t)(sp->NA * 8), tri_lambda);
      SQ_CHECK(launch_error("build_blk_kernel"));
    } else if (tri) {
      build_DSA_tri_kernel<<<(unsigned)(w->W / 256), 256, 0, s_build>>>(ket_dev, Dket, w->W, j0, len_eff, n, sp->d_strA, sp->d_strB,
                                                                      sp->d_rankA, sp->d_rankB, sp->NB);
      SQ_CHECK(launch_error("build_DSA_tri_kernel"));
    } else if (sym_route && pv_ket) {
      SQ_CHECK(launch_build_DSA_peer(sp, pv_ket, ket_dev, Dket, w->W, j0, half_band, s_build));
    } else if (sym_route) {
      build_DSA_kernel<<<(unsigned)(w->W / 256), 256, 0, s_build>>>(ket_dev, Dket, w->W, j0, len, n, sp->d_strA, sp->d_strB, sp->d_rankA,
                                                                  sp->d_rankB, sp->NB, sp->row_begin);
      SQ_CHECK(launch_error("build_DSA_kernel"));
    } else {
      SQ_CHECK(launch_build_D(sp, w, ket_dev, Dket, j0, s_build, use_const, false, pv_ket));
    }
    const double* Dbra = Dket;
    if (rdm2_host && !same) {
      SQ_CHECK(launch_build_D(sp, w, bra_dev, w->d_D[2 + b], j0, s_build, use_const, false, pv_bra));
      Dbra = w->d_D[2 + b];
    }
    if (piped) {
      SQ_CUDA(cudaEventRecord(w->ev_built[b], s_build));
      SQ_CUDA(cudaStreamWaitEvent(s_gemm, w->ev_built[b], 0));
    }
    if (!rdm1_from_G2) {
      // rdm1[pq] += sum_t bra[j0+t] * Dket[pq][t]   (one CTA per row, fixed summation order)
      SQ_CHECK(sq_panel_gemv(Dket, w->W, n2, bra_dev + j0, wl, d_g1, s_gemm));
    }
    if (sym_route) {
      SQ_CHECK(sq_gram_sym_panel(Dket, w->W, nS, nA, w->W, n_split, w->d_gram, s_gemm));
    } else if (rdm2_host) {
      // G2 row-major [a][b] = sum_t Dbra[a][t] Dket[b][t]: hand-written DMMA kernel, 128 x 128 output tiles, split-K partial
      // sums accumulated across the panels in per-(tile, split) slots (sqsv_dmma.cu); for bra == ket only the upper-triangular
      // tiles are computed (the Gram matrix is symmetric).  The panel is zero-padded beyond the last determinant.
      SQ_CHECK(sq_gram_panel(Dbra, Dket, w->W, n2, w->W, tiles, n_split, w->d_gram, s_gemm));
    }
    if (piped) SQ_CUDA(cudaEventRecord(w->ev_gemm[b], s_gemm));
  }
  if (piped) {
    SQ_CUDA(cudaEventRecord(w->ev_start, s_gemm));
    SQ_CUDA(cudaStreamWaitEvent(st, w->ev_start, 0));
    SQ_CUDA(cudaEventRecord(w->ev_start, s_build));
    SQ_CUDA(cudaStreamWaitEvent(st, w->ev_start, 0));
  }
  std::vector<double> G2h(rdm2_host ? (size_t)n2 * n2 : 0), g1h((size_t)n2);
  if (sym_route) {
    SQ_CHECK(sq_gram_sym_end(n_split, w->d_gram, w->d_gsym, st));
    std::vector<double> gs(2 * (size_t)GR * GR);
    SQ_CUDA(cudaMemcpyAsync(gs.data(), w->d_gsym, sizeof(double) * gs.size(), cudaMemcpyDeviceToHost, st));
    SQ_CUDA(cudaStreamSynchronize(st));
    assemble_g2_from_sym(n, gs.data(), gs.data() + (size_t)GR * GR, GR, n_elec, &G2h, &g1h);
  } else {
    if (rdm2_host) SQ_CHECK(sq_gram_end(tiles, n_split, w->d_gram, n2, same, d_G2, st));
    SQ_CUDA(cudaMemcpyAsync(g1h.data(), d_g1, sizeof(double) * n2, cudaMemcpyDeviceToHost, st));
    if (rdm2_host)
      SQ_CUDA(cudaMemcpyAsync(G2h.data(), d_G2, sizeof(double) * (size_t)n2 * n2, cudaMemcpyDeviceToHost, st));
    SQ_CUDA(cudaStreamSynchronize(st));
  }
  if (rdm1_from_G2 && !sym_route) {
    for (int p = 0; p < n; ++p)
      for (int q = 0; q < n; ++q) {
        double v = 0.0;
        for (int r = 0; r < n; ++r) v += G2h[((size_t)(q * n + p)) * n2 + (r * n + r)];
        g1h[(size_t)p * n + q] = v / n_elec;
      }
  }
  for (int i = 0; i < n2; ++i) rdm1_host[i] = g1h[i];
  if (rdm2_host) {
    // rdm2[p][q][r][s] = <bra|E_pq E_rs|ket> - delta_qr rdm1[p][s];  <..> = G2[(q,p)][(r,s)]
    for (int p = 0; p < n; ++p)
      for (int q = 0; q < n; ++q)
        for (int r = 0; r < n; ++r)
          for (int s = 0; s < n; ++s) {
            double v = G2h[((size_t)(q * n + p)) * n2 + (r * n + s)];
            if (q == r) v -= g1h[p * n + s];
            rdm2_host[(((size_t)p * n + q) * n + r) * n + s] = v;
          }
  }
  return SQ_OK;
}

// Symmetrised panel (see build_Dsym_kernel) of an alpha-sharded vector: alpha partners through the peer mappings.
__global__ void __launch_bounds__(256)
build_Dsym_peer_kernel(PeerView pv, const double* __restrict__ IN, double* __restrict__ D, int64_t W, int64_t j0, int64_t len,
                       const ERec* __restrict__ etab, int n, const uint32_t* __restrict__ strA,
                       const uint32_t* __restrict__ strB, const int32_t* __restrict__ rankA,
                       const int32_t* __restrict__ rankB, int64_t NB, int64_t row_begin) {
  const int n2 = n * n;
  const ERec* sm = stage_etab<false>(etab, n2);
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= W) return;
  const int64_t j = j0 + t;
  const int nS = n * (n + 1) / 2;
  if (j >= len) {
    for (int slot = 0; slot < nS; ++slot) D[(int64_t)slot * W + t] = 0.0;
    return;
  }
  const int64_t ia_loc = j / NB, ib = j - ia_loc * NB;
  const uint32_t a = __ldg(strA + row_begin + ia_loc), b = __ldg(strB + ib);
  auto elem = [&](int slot) -> double {
    double v = 0.0;
    const ERec ra = sm[2 * slot], rb = sm[2 * slot + 1];
    if ((a & ra.tocc) == ra.tocc && (a & ra.temp) == 0u) {
      const uint32_t sa = a ^ ra.flip;
      const int par = (__popc(sa & ra.parS) + __popc(b & ra.parO)) & 1;
      const int64_t gr = __ldg(rankA + sa);          // global row of the alpha partner
      int o = 0;
      while (o + 1 < pv.world && gr >= pv.row_starts[o + 1]) ++o;
      v += (par ? -ra.s0 : ra.s0) * pv.p[o][(gr - pv.row_starts[o]) * NB + ib];
    }
    if ((b & rb.tocc) == rb.tocc && (b & rb.temp) == 0u) {
      const uint32_t sb = b ^ rb.flip;
      const int par = (__popc(sb & rb.parS) + __popc(a & rb.parO)) & 1;
      v += (par ? -rb.s0 : rb.s0) * IN[ia_loc * NB + __ldg(rankB + sb)];
    }
    return v;
  };
  int slot = 0;
  for (int r = 0; r < n; ++r)
    for (int q = 0; q <= r; ++q, ++slot) D[(int64_t)slot * W + t] = (r == q) ? elem(r * n + r) : elem(r * n + q) + elem(q * n + r);
}

// ---- sigma of an alpha-sharded vector ----------------------------------------------------------------------------------
// Same Knowles-Handy panels as sq_sigma (symmetrised generators for real-orbital integrals, general n^2 panel otherwise), one rank = the determinants of its
// own rows as SOURCES: D is gathered with alpha partners read over NVLink (build_D_peer_kernel, as in sq_rdm12_dist),
// F = 1/2 g D is a local DGEMM, and E_pq-images whose alpha row lives on another GPU are accumulated in place in the
// owner's shard with system-scope fp64 atomics through the peer mapping -- the exchange step is inside the scatter
// kernel, there is no transpose.  Several GPUs (and the owner itself) may add to one element concurrently, hence
// atomicAdd_system for EVERY accumulation into OUT, local ones included.
struct PeerViewRW {
  double* p[SQ_MAX_WORLD];
  int64_t row_starts[SQ_MAX_WORLD + 1];
  int world;
};
__global__ void __launch_bounds__(256)
scatter_E_peer_kernel(PeerViewRW pvo, const double* __restrict__ IN, double* OUT /* == pvo.p[own rank]: no restrict */, const double* __restrict__ F,
                      const double* __restrict__ kmat, const int* __restrict__ frow, int64_t W, int64_t j0, int64_t len,
                      const ERec* __restrict__ etab, int n2,
                      const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB, const int32_t* __restrict__ rankA,
                      const int32_t* __restrict__ rankB, int64_t NB, int64_t row_begin) {
  const ERec* sm = stage_etab<false>(etab, n2);
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t j = j0 + t;
  if (t >= W || j >= len) return;
  const int64_t ia_loc = j / NB, ib = j - ia_loc * NB;
  const uint32_t a = __ldg(strA + row_begin + ia_loc), b = __ldg(strB + ib);
  const double cj = IN[j];
  double diag = 0.0;
  for (int slot = 0; slot < n2; ++slot) {
    const ERec ra = sm[2 * slot], rb = sm[2 * slot + 1];
    const bool va = (a & ra.occ) == ra.occ && (a & ra.emp) == 0u;
    const bool vb = (b & rb.occ) == rb.occ && (b & rb.emp) == 0u;
    if (!va && !vb) continue;
    const double val = F[(int64_t)__ldg(frow + slot) * W + t] + __ldg(kmat + slot) * cj;   // frow: row of F that holds (p,q)
    if (va) {
      const int par = (__popc(a & ra.parS) + __popc(b & ra.parO)) & 1;
      const double sv = (par ? -ra.s0 : ra.s0) * val;
      if (ra.flip == 0u) {
        diag += sv;
      } else {
        const int64_t gr = __ldg(rankA + (a ^ ra.flip));   // global row of the image; its owner may be another GPU
        int o = 0;
        while (o + 1 < pvo.world && gr >= pvo.row_starts[o + 1]) ++o;
        atomicAdd_system(pvo.p[o] + (gr - pvo.row_starts[o]) * NB + ib, sv);
      }
    }
    if (vb) {
      const int par = (__popc(b & rb.parS) + __popc(a & rb.parO)) & 1;
      const double sv = (par ? -rb.s0 : rb.s0) * val;
      if (rb.flip == 0u) diag += sv;
      else atomicAdd_system(OUT + ia_loc * NB + __ldg(rankB + (b ^ rb.flip)), sv);
    }
  }
  atomicAdd_system(OUT + j, diag);
}

// out (all shards) += sum_pq k_pq E_pq |in> + 1/2 sum_pqrs g_pqrs E_pq E_rs |in> restricted to THIS rank's source rows.
// Protocol (the caller's, slowquant_b200/distributed.py::sigma_sharded): every rank sets its out shard to e_core * in,
// barrier, every rank calls sq_sigma_dist, stream synchronise, barrier -- then out holds H|in>.
// ---------------------------------------------------------------------------------------------------------------------------
// Spin-flip symmetric ALPHA-SHARDED vectors: half of the sigma build on every rank.  The upper triangle of the single-device
// route would leave the ranks with the first rows nearly all of the work, so the kept half is the cyclic band
//     K = { (ia, ib) : (ib - ia) mod N <= (N - 1) / 2,  ties (N even, distance N / 2) kept for ia < ib },
// which holds the diagonal, exactly one of every pair (J, J^T), and the same number of columns (+-1) in every row.  Sources are the
// kept determinants of the local rows, scattered to ALL their targets exactly as in the full build (S' = sum over kept sources,
// diagonal sources with weight 1/2; folding the targets onto the kept half instead would turn the contiguous atomics of a warp into
// column-strided ones -- measured: 8 x SLOWER than the full build over NVLink).  Since c[J^T] H|J^T> = lambda U (c[J] H|J>),
// sigma = S' + lambda U S': after a device-wide barrier sq_spinsym_mirror_dist symmetrises the vector in place, pair by pair, 32 x 32
// tiles through shared memory so that remote reads and writes are 256-byte row segments.  The caller starts from
// out = 1/2 e_core in (the symmetrisation doubles it).
// ---------------------------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ bool half_kept(int64_t ia, int64_t ib, int64_t N) {
  int64_t d = ib - ia;
  if (d < 0) d += N;
  if (N & 1) return d <= (N - 1) / 2;
  return d < N / 2 || (d == N / 2 && ia < ib);
}
__host__ __device__ __forceinline__ int64_t half_count(int64_t ia, int64_t N) {   // kept columns of row ia: distances 0 .. count - 1
  return (N & 1) ? (N + 1) / 2 : N / 2 + (ia < N / 2 ? 1 : 0);
}
// kept determinants of rows [r0, r1) in row-major order of (row, distance)
__host__ __device__ __forceinline__ int64_t half_len(int64_t r0, int64_t r1, int64_t N) {
  if (N & 1) return (r1 - r0) * ((N + 1) / 2);
  const int64_t mid = N / 2, n1 = (r0 < mid ? (r1 < mid ? r1 : mid) - r0 : 0);
  return n1 * (N / 2 + 1) + (r1 - r0 - n1) * (N / 2);
}
__device__ __forceinline__ void half_unrank(int64_t j, int64_t r0, int64_t r1, int64_t N, int64_t* ia, int64_t* ib) {
  int64_t row, d;
  if (N & 1) {
    const int64_t c = (N + 1) / 2;
    row = r0 + j / c;
    d = j % c;
  } else {
    const int64_t mid = N / 2, n1 = (r0 < mid ? (r1 < mid ? r1 : mid) - r0 : 0), c1 = N / 2 + 1, c0 = N / 2;
    if (j < n1 * c1) {
      row = r0 + j / c1;
      d = j % c1;
    } else {
      const int64_t jj = j - n1 * c1;
      row = r0 + n1 + jj / c0;
      d = jj % c0;
    }
  }
  *ia = row;
  int64_t c = row + d;
  if (c >= N) c -= N;
  *ib = c;
}
__device__ __forceinline__ int peer_owner(const int64_t* row_starts, int world, int64_t gr) {
  int o = 0;
  while (o + 1 < world && gr >= row_starts[o + 1]) ++o;
  return o;
}

// per-rank maxima of the symmetry test over the local rows (all columns): res as in spinsym_check_kernel
__global__ void __launch_bounds__(256)
spinsym_check_peer_kernel(PeerView pv, const double* __restrict__ IN, int64_t N, int64_t row_begin, int64_t n_rows,
                          const uint32_t* __restrict__ str, unsigned long long* __restrict__ res) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t bi = blockIdx.y, bj = blockIdx.x;
  for (int k = ty; k < 32; k += 8) {   // mirror tile: rows of column block bj (their owners' shards), columns = the local rows of block bi
    const int64_t gr = bj * 32 + k, lc = bi * 32 + tx;
    double v = 0.0;
    if (gr < N && lc < n_rows) {
      const int o = peer_owner(pv.row_starts, pv.world, gr);
      v = pv.p[o][(gr - pv.row_starts[o]) * N + row_begin + lc];
    }
    tile[k][tx] = v;
  }
  __syncthreads();
  double m0 = 0.0, m1 = 0.0, m2 = 0.0;
  for (int k = ty; k < 32; k += 8) {
    const int64_t lr = bi * 32 + k, c = bj * 32 + tx;
    if (lr < n_rows && c < N) {
      const double x = IN[lr * N + c], y = tile[tx][k];
      const double ph = (__popc(__ldg(str + row_begin + lr) & __ldg(str + c)) & 1) ? -x : x;
      m0 = fmax(m0, fmax(fabs(x), fabs(y)));
      m1 = fmax(m1, fabs(y - ph));
      m2 = fmax(m2, fabs(y + ph));
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    m0 = fmax(m0, __shfl_down_sync(0xffffffffu, m0, off));
    m1 = fmax(m1, __shfl_down_sync(0xffffffffu, m1, off));
    m2 = fmax(m2, __shfl_down_sync(0xffffffffu, m2, off));
  }
  if (tx == 0) {
    atomicMax(res + 0, (unsigned long long)__double_as_longlong(m0));
    atomicMax(res + 1, (unsigned long long)__double_as_longlong(m1));
    atomicMax(res + 2, (unsigned long long)__double_as_longlong(m2));
  }
}

__global__ void __launch_bounds__(256)
build_Dsym_half_peer_kernel(PeerView pv, const double* __restrict__ IN, double* __restrict__ D, int64_t W, int64_t j0, int64_t len,
                            const ERec* __restrict__ etab, int n, const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB,
                            const int32_t* __restrict__ rankA, const int32_t* __restrict__ rankB, int64_t NB, int64_t row_begin,
                            int64_t row_end) {
  const int n2 = n * n;
  const ERec* sm = stage_etab<false>(etab, n2);
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= W) return;
  const int64_t j = j0 + t;
  const int nS = n * (n + 1) / 2;
  if (j >= len) {
    for (int slot = 0; slot < nS; ++slot) D[(int64_t)slot * W + t] = 0.0;
    return;
  }
  int64_t ia, ib;
  half_unrank(j, row_begin, row_end, NB, &ia, &ib);
  const int64_t ia_loc = ia - row_begin;
  const uint32_t a = __ldg(strA + ia), b = __ldg(strB + ib);
  auto elem = [&](int slot) -> double {
    double v = 0.0;
    const ERec ra = sm[2 * slot], rb = sm[2 * slot + 1];
    if ((a & ra.tocc) == ra.tocc && (a & ra.temp) == 0u) {
      const uint32_t sa = a ^ ra.flip;
      const int par = (__popc(sa & ra.parS) + __popc(b & ra.parO)) & 1;
      const int64_t gr = __ldg(rankA + sa);
      const int o = peer_owner(pv.row_starts, pv.world, gr);
      v += (par ? -ra.s0 : ra.s0) * pv.p[o][(gr - pv.row_starts[o]) * NB + ib];
    }
    if ((b & rb.tocc) == rb.tocc && (b & rb.temp) == 0u) {
      const uint32_t sb = b ^ rb.flip;
      const int par = (__popc(sb & rb.parS) + __popc(a & rb.parO)) & 1;
      v += (par ? -rb.s0 : rb.s0) * IN[ia_loc * NB + __ldg(rankB + sb)];
    }
    return v;
  };
  int slot = 0;
  for (int r = 0; r < n; ++r)
    for (int q = 0; q <= r; ++q, ++slot) D[(int64_t)slot * W + t] = (r == q) ? elem(r * n + r) : elem(r * n + q) + elem(q * n + r);
}

__global__ void __launch_bounds__(256)
scatter_E_half_peer_kernel(PeerViewRW pvo, const double* __restrict__ IN, const double* __restrict__ F, const double* __restrict__ kmat,
                           const int* __restrict__ frow, int64_t W, int64_t j0, int64_t len, const ERec* __restrict__ etab, int n2,
                           const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB, const int32_t* __restrict__ rankA,
                           const int32_t* __restrict__ rankB, int64_t NB, int64_t row_begin, int64_t row_end) {
  // the scatter of scatter_E_peer_kernel from the KEPT sources only, to ALL targets (no folding: the atomics of a warp stay
  // contiguous, also the remote ones); a source on the diagonal enters with weight 1/2.  With S' this partial sum,
  // sigma = S' + lambda U S' (spinsym_symmetrize_peer_kernel).
  const ERec* sm = stage_etab<false>(etab, n2);
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t j = j0 + t;
  if (t >= W || j >= len) return;
  int64_t ia, ib;
  half_unrank(j, row_begin, row_end, NB, &ia, &ib);
  const uint32_t a = __ldg(strA + ia), b = __ldg(strB + ib);
  const double wsrc = (ia == ib) ? 0.5 : 1.0;
  const double cj = IN[(ia - row_begin) * NB + ib];
  double* orow = pvo.p[peer_owner(pvo.row_starts, pvo.world, ia)] + (ia - pvo.row_starts[peer_owner(pvo.row_starts, pvo.world, ia)]) * NB;
  double diag = 0.0;
  for (int slot = 0; slot < n2; ++slot) {
    const ERec ra = sm[2 * slot], rb = sm[2 * slot + 1];
    const bool va = (a & ra.occ) == ra.occ && (a & ra.emp) == 0u;
    const bool vb = (b & rb.occ) == rb.occ && (b & rb.emp) == 0u;
    if (!va && !vb) continue;
    const double val = wsrc * (F[(int64_t)__ldg(frow + slot) * W + t] + __ldg(kmat + slot) * cj);
    if (va) {
      const int par = (__popc(a & ra.parS) + __popc(b & ra.parO)) & 1;
      const double sv = (par ? -ra.s0 : ra.s0) * val;
      if (ra.flip == 0u) {
        diag += sv;
      } else {
        const int64_t gr = __ldg(rankA + (a ^ ra.flip));
        const int o = peer_owner(pvo.row_starts, pvo.world, gr);
        atomicAdd_system(pvo.p[o] + (gr - pvo.row_starts[o]) * NB + ib, sv);
      }
    }
    if (vb) {
      const int par = (__popc(b & rb.parS) + __popc(a & rb.parO)) & 1;
      const double sv = (par ? -rb.s0 : rb.s0) * val;
      if (rb.flip == 0u) diag += sv;
      else atomicAdd_system(orow + __ldg(rankB + (b ^ rb.flip)), sv);
    }
  }
  atomicAdd_system(orow + ib, diag);
}

// In place: X <- X + lambda U X, i.e. x[I] <- x[I] + lambda phi(I) x[I^T] for every determinant.  Every unordered pair {I, I^T}
// is handled by exactly one thread -- the one that owns the KEPT member (I = (ia, ib), ia a local row) -- which reads both
// values (the mirror may sit in another rank's shard), and writes both; 32 x 32 tiles go through shared memory so that the
// remote reads and writes are 256-byte row segments.  A diagonal element becomes (1 + lambda phi) x.
__global__ void __launch_bounds__(256)
spinsym_symmetrize_peer_kernel(PeerViewRW pvo, double* OUT /* == pvo.p[own rank] */, int64_t N, int64_t row_begin, int64_t n_rows,
                               const uint32_t* __restrict__ str, double lambda) {
  __shared__ double mir[32][33];    // mirror tile: rows of column block bj, columns = the local rows of block bi
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t bi = blockIdx.y, bj = blockIdx.x;
  for (int k = ty; k < 32; k += 8) {
    const int64_t gr = bj * 32 + k, lc = bi * 32 + tx;
    double v = 0.0;
    if (gr < N && lc < n_rows) {
      const int o = peer_owner(pvo.row_starts, pvo.world, gr);
      v = pvo.p[o][(gr - pvo.row_starts[o]) * N + row_begin + lc];
    }
    mir[k][tx] = v;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int64_t lr = bi * 32 + k, c = bj * 32 + tx, gr = row_begin + lr;
    double nm = 0.0;
    bool wr = false;
    if (lr < n_rows && c < N && half_kept(gr, c, N)) {
      const double x = OUT[lr * N + c], y = mir[tx][k];
      const double s = ((__popc(__ldg(str + gr) & __ldg(str + c)) & 1) ? -lambda : lambda);
      OUT[lr * N + c] = x + s * y;
      nm = y + s * x;
      wr = gr != c;
    }
    mir[tx][k] = wr ? nm : __longlong_as_double(0x7ff8dead00000000LL);   // tag: not to be written back
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int64_t gr = bj * 32 + k, lc = bi * 32 + tx;
    const double v = mir[k][tx];
    if (gr < N && lc < n_rows && __double_as_longlong(v) != 0x7ff8dead00000000LL) {
      const int o = peer_owner(pvo.row_starts, pvo.world, gr);
      pvo.p[o][(gr - pvo.row_starts[o]) * N + row_begin + lc] = v;
    }
  }
}

// S / A generator panel of an alpha-sharded vector (alpha partners through the peer mappings).  HALF: only the kept cyclic band of the
// local rows, off-diagonal columns weighted by sqrt(2) (spin-flip symmetric vector: the Gram sums over all determinants are
// 2 sum_kept-off-diagonal + sum_diagonal).
template <bool HALF>
__global__ void __launch_bounds__(256)
build_DSA_peer_kernel(PeerView pv, const double* __restrict__ IN, double* __restrict__ D, int64_t W, int64_t j0, int64_t len, int n,
                      const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB, const int32_t* __restrict__ rankA,
                      const int32_t* __restrict__ rankB, int64_t NB, int64_t row_begin, int64_t row_end) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= W) return;
  const int64_t j = j0 + t;
  const int n2 = n * n, nS = n * (n + 1) / 2;
  if (j >= len) {
    for (int slot = 0; slot < n2; ++slot) D[(int64_t)slot * W + t] = 0.0;
    return;
  }
  int64_t ia, ib;
  if (HALF) {
    half_unrank(j, row_begin, row_end, NB, &ia, &ib);
  } else {
    const int64_t l = j / NB;
    ia = row_begin + l;
    ib = j - l * NB;
  }
  const int64_t ia_loc = ia - row_begin;
  const uint32_t a = __ldg(strA + ia), b = __ldg(strB + ib);
  const double wgt = (HALF && ia != ib) ? 1.4142135623730951 : 1.0;
  auto elem = [&](int p, int q) -> double {
    double v = 0.0;
    const ERec ra = erec_closed(p, q, 0), rb = erec_closed(p, q, 1);
    if ((a & ra.tocc) == ra.tocc && (a & ra.temp) == 0u) {
      const uint32_t sa = a ^ ra.flip;
      const int par = (__popc(sa & ra.parS) + __popc(b & ra.parO)) & 1;
      const int64_t gr = __ldg(rankA + sa);
      const int o = peer_owner(pv.row_starts, pv.world, gr);
      v += (par ? -ra.s0 : ra.s0) * pv.p[o][(gr - pv.row_starts[o]) * NB + ib];
    }
    if ((b & rb.tocc) == rb.tocc && (b & rb.temp) == 0u) {
      const uint32_t sb = b ^ rb.flip;
      const int par = (__popc(sb & rb.parS) + __popc(a & rb.parO)) & 1;
      v += (par ? -rb.s0 : rb.s0) * IN[ia_loc * NB + __ldg(rankB + sb)];
    }
    return v;
  };
  int ss = 0, as = nS;
  for (int p = 0; p < n; ++p)
    for (int q = 0; q <= p; ++q, ++ss) {
      if (p == q) {
        D[(int64_t)ss * W + t] = wgt * elem(p, p);
      } else {
        const double x = elem(p, q), y = elem(q, p);
        D[(int64_t)ss * W + t] = wgt * (x + y);
        D[(int64_t)as * W + t] = wgt * (x - y);
        ++as;
      }
    }
}
// launcher used by rdm12_impl (defined above these kernels); *len_eff = number of panel columns of this rank
static int launch_build_DSA_peer(sq_space* sp, const PeerView* pv, const double* in, double* D, int64_t W, int64_t j0, bool half,
                                 cudaStream_t st) {
  const int64_t len = half ? half_len(sp->row_begin, sp->row_end, sp->NB) : sp->local_len();
  if (half)
    build_DSA_peer_kernel<true><<<(unsigned)(W / 256), 256, 0, st>>>(*pv, in, D, W, j0, len, sp->n_orb, sp->d_strA, sp->d_strB, sp->d_rankA,
                                                                     sp->d_rankB, sp->NB, sp->row_begin, sp->row_end);
  else
    build_DSA_peer_kernel<false><<<(unsigned)(W / 256), 256, 0, st>>>(*pv, in, D, W, j0, len, sp->n_orb, sp->d_strA, sp->d_strB, sp->d_rankA,
                                                                      sp->d_rankB, sp->NB, sp->row_begin, sp->row_end);
  return launch_error("build_DSA_peer_kernel");
}
static int64_t half_len_host(const sq_space* sp) { return half_len(sp->row_begin, sp->row_end, sp->NB); }

static int sigma_dist_impl(sq_space* sp, const double* h_act_host, const double* g_act_host, const double* const* in_ptrs_host,
                           double* const* out_ptrs_host, double lambda, int* used_half, void* stream);

extern "C" int sq_sigma_dist(sq_space* sp, const double* h_act_host, const double* g_act_host, const double* const* in_ptrs_host,
                             double* const* out_ptrs_host, void* stream) {
  return sigma_dist_impl(sp, h_act_host, g_act_host, in_ptrs_host, out_ptrs_host, 0.0, nullptr, stream);
}

// lambda = +-1: the input is spin-flip symmetric (c[B,A] = lambda phi c[A,B], measured by sq_spinsym_measure_dist + a MAX all-reduce):
// build sigma on the kept half only.  *used_half = 1 if that happened (real-orbital integrals, n_alpha = n_beta): the caller then
// puts a device-wide barrier and calls sq_spinsym_mirror_dist; 0: the full build ran (nothing else to do).
extern "C" int sq_sigma_dist_sym(sq_space* sp, const double* h_act_host, const double* g_act_host, const double* const* in_ptrs_host,
                                 double* const* out_ptrs_host, double lambda, int* used_half, void* stream) {
  if (!used_half || (lambda != 1.0 && lambda != -1.0 && lambda != 0.0)) return SQ_ERR_INVALID;
  return sigma_dist_impl(sp, h_act_host, g_act_host, in_ptrs_host, out_ptrs_host, lambda, used_half, stream);
}

static int fill_peer_views(sq_space* sp, const double* const* in_ptrs_host, double* const* out_ptrs_host, PeerView* pin, PeerViewRW* pout) {
  if (sp->world < 1 || sp->world > SQ_MAX_WORLD || (int)sp->row_starts.size() != sp->world + 1) {
    sq_set_error("the space has no row partition (sq_space_set_partition)");
    return SQ_ERR_INVALID;
  }
  for (int r = 0; r < SQ_MAX_WORLD; ++r) {
    if (pin) pin->p[r] = (in_ptrs_host && r < sp->world) ? in_ptrs_host[r] : nullptr;
    if (pout) pout->p[r] = (out_ptrs_host && r < sp->world) ? out_ptrs_host[r] : nullptr;
    if (r < sp->world && ((pin && !pin->p[r]) || (pout && !pout->p[r]))) {
      sq_set_error("missing shard pointer for rank %d", r);
      return SQ_ERR_INVALID;
    }
  }
  for (int r = 0; r <= SQ_MAX_WORLD; ++r) {
    const int64_t v = r <= sp->world ? sp->row_starts[r] : sp->row_starts[sp->world];
    if (pin) pin->row_starts[r] = v;
    if (pout) pout->row_starts[r] = v;
  }
  if (pin) pin->world = sp->world;
  if (pout) pout->world = sp->world;
  return SQ_OK;
}

// res3_host = this rank's {max |c|, max |c[B,A] - phi c[A,B]|, max |c[B,A] + phi c[A,B]|} over its rows (the caller takes the MAX over
// the ranks: lambda = +1 if the second, -1 if the third is <= 1e-12 of the first).  n_alpha = n_beta only; reads remote shards.
extern "C" int sq_spinsym_measure_dist(sq_space* sp, const double* const* in_ptrs_host, double* res3_host, void* stream) {
  if (!sp || !in_ptrs_host || !res3_host || sp->device < 0) return SQ_ERR_INVALID;
  res3_host[0] = res3_host[1] = res3_host[2] = 0.0;
  if (sp->n_alpha != sp->n_beta || sp->NA != sp->NB) {
    res3_host[1] = res3_host[2] = 1.0;   // never symmetric
    return SQ_OK;
  }
  PeerView pin;
  SQ_CHECK(fill_peer_views(sp, in_ptrs_host, nullptr, &pin, nullptr));
  const int64_t n_rows = sp->row_end - sp->row_begin;
  if (n_rows == 0) return SQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  SQ_CUDA(cudaSetDevice(sp->device));
  HamWork* w = nullptr;
  SQ_CHECK(get_work(sp, false, true, &w));
  if (!w->d_symres) SQ_CUDA(cudaMalloc(&w->d_symres, 3 * sizeof(unsigned long long)));
  SQ_CUDA(cudaMemsetAsync(w->d_symres, 0, 3 * sizeof(unsigned long long), st));
  spinsym_check_peer_kernel<<<dim3((unsigned)((sp->NB + 31) / 32), (unsigned)((n_rows + 31) / 32)), 256, 0, st>>>(
      pin, pin.p[sp->rank], sp->NB, sp->row_begin, n_rows, sp->d_strA, w->d_symres);
  SQ_CHECK(launch_error("spinsym_check_peer_kernel"));
  SQ_CUDA(cudaMemcpyAsync(res3_host, w->d_symres, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  SQ_CUDA(cudaStreamSynchronize(st));
  return SQ_OK;
}

// second half of sq_sigma_dist_sym (after a device-wide barrier): the elements outside the kept band from their mirrors
extern "C" int sq_spinsym_mirror_dist(sq_space* sp, double* const* out_ptrs_host, double lambda, void* stream) {
  if (!sp || !out_ptrs_host || sp->device < 0 || (lambda != 1.0 && lambda != -1.0)) return SQ_ERR_INVALID;
  PeerViewRW pout;
  SQ_CHECK(fill_peer_views(sp, nullptr, out_ptrs_host, nullptr, &pout));
  const int64_t n_rows = sp->row_end - sp->row_begin;
  if (n_rows == 0) return SQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  SQ_CUDA(cudaSetDevice(sp->device));
  spinsym_symmetrize_peer_kernel<<<dim3((unsigned)((sp->NB + 31) / 32), (unsigned)((n_rows + 31) / 32)), 256, 0, st>>>(
      pout, pout.p[sp->rank], sp->NB, sp->row_begin, n_rows, sp->d_strA, lambda);
  return launch_error("spinsym_symmetrize_peer_kernel");
}

static int sigma_dist_impl(sq_space* sp, const double* h_act_host, const double* g_act_host, const double* const* in_ptrs_host,
                           double* const* out_ptrs_host, double lambda, int* used_half, void* stream) {
  SqRange nvtx_range("sq_sigma_dist");
  if (used_half) *used_half = 0;
  if (!sp || !h_act_host || !g_act_host || !in_ptrs_host || !out_ptrs_host) return SQ_ERR_INVALID;
  if (sp->device < 0) return SQ_ERR_INVALID;
  if (sp->world < 1 || sp->world > SQ_MAX_WORLD || (int)sp->row_starts.size() != sp->world + 1) {
    sq_set_error("sq_sigma_dist: the space has no row partition (sq_space_set_partition)");
    return SQ_ERR_INVALID;
  }
  PeerView pin;
  PeerViewRW pout;
  pin.world = pout.world = sp->world;
  for (int r = 0; r < SQ_MAX_WORLD; ++r) {
    pin.p[r] = r < sp->world ? in_ptrs_host[r] : nullptr;
    pout.p[r] = r < sp->world ? out_ptrs_host[r] : nullptr;
    if (r < sp->world && (!pin.p[r] || !pout.p[r] || pin.p[r] == pout.p[r])) {
      sq_set_error("sq_sigma_dist: missing or aliased shard pointer for rank %d", r);
      return SQ_ERR_INVALID;
    }
  }
  for (int r = 0; r <= SQ_MAX_WORLD; ++r)
    pin.row_starts[r] = pout.row_starts[r] = r <= sp->world ? sp->row_starts[r] : sp->row_starts[sp->world];
  const int64_t len = sp->local_len();
  if (len == 0) return SQ_OK;   // a rank without rows has no sources
  cudaStream_t st = (cudaStream_t)stream;
  SQ_CUDA(cudaSetDevice(sp->device));
  if (sp->world > 1) {
    // the images that land in another rank's rows are added with system-scope fp64 atomics through the peer mapping: only
    // correct when the device pair has native P2P atomics (NVLink); refuse anything else instead of returning a wrong sigma
    int n_dev = 0;
    SQ_CUDA(cudaGetDeviceCount(&n_dev));
    for (int d = 0; d < n_dev; ++d) {
      if (d == sp->device) continue;
      int can = 0, native = 0;
      SQ_CUDA(cudaDeviceCanAccessPeer(&can, sp->device, d));
      if (!can) continue;   // not a peer of this process's device: its shard cannot be one of the mapped pointers
      SQ_CUDA(cudaDeviceGetP2PAttribute(&native, cudaDevP2PAttrNativeAtomicSupported, sp->device, d));
      if (!native) {
        sq_set_error("sq_sigma_dist: devices %d and %d have no native peer-to-peer atomics (not NVLink-connected); use the RDM route "
                     "(sq_rdm12_dist) for the energy of a sharded vector on this topology", sp->device, d);
        return SQ_ERR_UNSUPPORTED;
      }
    }
  }
  HamWork* w = nullptr;
  SQ_CHECK(get_work(sp, false, true, &w));
  const int n = sp->n_orb, n2 = n * n;
  std::vector<double> k((size_t)n2);
  double gmax = 0.0;
  for (size_t i = 0; i < (size_t)n2 * n2; ++i) gmax = std::max(gmax, std::fabs(g_act_host[i]));
  for (int p = 0; p < n; ++p)
    for (int q = 0; q < n; ++q) {
      double v = h_act_host[p * n + q];
      for (int r = 0; r < n; ++r) v -= 0.5 * g_act_host[(((size_t)p * n + r) * n + r) * n + q];
      k[(size_t)p * n + q] = v;
    }
  // same switch as sq_sigma: integrals with g_pqrs = g_qprs = g_pqsr (and k_pq = k_qp) take the n (n + 1) / 2 symmetrised
  // generators (3.6 x fewer DGEMM flops at n = 20), anything else the general n^2 panel
  bool sym = true;
  const double tol = 1e-13 * (gmax > 0 ? gmax : 1.0);
  auto G = [&](int p, int q, int r, int t) { return g_act_host[(((size_t)p * n + q) * n + r) * n + t]; };
  for (int p = 0; p < n && sym; ++p)
    for (int q = 0; q < n && sym; ++q) {
      if (std::fabs(k[(size_t)p * n + q] - k[(size_t)q * n + p]) > 1e-13 * (1.0 + std::fabs(k[(size_t)p * n + q]))) sym = false;
      for (int r = 0; r < n && sym; ++r)
        for (int t = 0; t < n; ++t)
          if (std::fabs(G(p, q, r, t) - G(q, p, r, t)) > tol || std::fabs(G(p, q, r, t) - G(p, q, t, r)) > tol) {
            sym = false;
            break;
          }
    }
  const int nS = n * (n + 1) / 2;
  const int nrow = sym ? nS : n2;   // rows of the D and F panels
  const int ldg = (nrow + 1) & ~1;  // even leading dimension of the integral matrix (16-byte rows for the async copies)
  std::vector<double> Gm((size_t)nrow * ldg, 0.0);
  std::vector<int> frow((size_t)n2);
  if (sym) {
    auto slot = [](int r, int t) { return r >= t ? r * (r + 1) / 2 + t : t * (t + 1) / 2 + r; };
    for (int p = 0; p < n; ++p)
      for (int q = 0; q <= p; ++q)
        for (int r = 0; r < n; ++r)
          for (int t = 0; t <= r; ++t) Gm[(size_t)slot(p, q) * ldg + slot(r, t)] = 0.5 * G(p, q, r, t);
    for (int p = 0; p < n; ++p)
      for (int q = 0; q < n; ++q) frow[(size_t)p * n + q] = slot(p, q);
  } else {
    for (int a = 0; a < n2; ++a)
      for (int b = 0; b < n2; ++b) Gm[(size_t)a * ldg + b] = 0.5 * g_act_host[(size_t)a * n2 + b];
    for (int i = 0; i < n2; ++i) frow[i] = i;
  }
  double* d_G = w->d_small;
  double* d_k = w->d_small + (size_t)n2 * (n2 + 1);
  SQ_CUDA(cudaMemcpyAsync(d_G, Gm.data(), sizeof(double) * Gm.size(), cudaMemcpyHostToDevice, st));
  SQ_CUDA(cudaMemcpyAsync(d_k, k.data(), sizeof(double) * k.size(), cudaMemcpyHostToDevice, st));
  SQ_CUDA(cudaMemcpyAsync(w->d_frow, frow.data(), sizeof(int) * frow.size(), cudaMemcpyHostToDevice, st));
  SQ_CUDA(cudaStreamSynchronize(st));   // host vectors go out of scope below
  const double* in_dev = pin.p[sp->rank];
  double* out_dev = pout.p[sp->rank];
  const size_t smem = sizeof(ERec) * 2 * (size_t)n2;
  allow_smem(scatter_E_peer_kernel, smem);
  if (lambda != 0.0 && !(sym && sp->n_alpha == sp->n_beta && sp->NA == sp->NB)) {
    sq_set_error("sq_sigma_dist_sym: the half build needs real-orbital integrals (g_pqrs = g_qprs = g_pqsr) and n_alpha = n_beta");
    return SQ_ERR_UNSUPPORTED;
  }
  if (lambda != 0.0) {
    // spin-flip symmetric input: the kept half of the local rows (cyclic band, see above)
    const int64_t len_half = half_len(sp->row_begin, sp->row_end, sp->NB);
    allow_smem(build_Dsym_half_peer_kernel, smem);
    allow_smem(scatter_E_half_peer_kernel, smem);
    for (int64_t j0 = 0; j0 < len_half; j0 += w->W) {
      build_Dsym_half_peer_kernel<<<(unsigned)(w->W / 256), 256, smem, st>>>(pin, in_dev, w->d_D[0], w->W, j0, len_half, w->d_etab, n,
                                                                             sp->d_strA, sp->d_strB, sp->d_rankA, sp->d_rankB, sp->NB,
                                                                             sp->row_begin, sp->row_end);
      SQ_CHECK(launch_error("build_Dsym_half_peer_kernel"));
      SQ_CHECK(sq_sigma_gemm(d_G, ldg, w->d_D[0], w->d_F[0], nrow, w->W, st));
      scatter_E_half_peer_kernel<<<(unsigned)(w->W / 256), 256, smem, st>>>(pout, in_dev, w->d_F[0], d_k, w->d_frow, w->W, j0, len_half,
                                                                            w->d_etab, n2, sp->d_strA, sp->d_strB, sp->d_rankA, sp->d_rankB,
                                                                            sp->NB, sp->row_begin, sp->row_end);
      SQ_CHECK(launch_error("scatter_E_half_peer_kernel"));
    }
    if (used_half) *used_half = 1;
    return SQ_OK;
  }
  for (int64_t j0 = 0; j0 < len; j0 += w->W) {   // one stream: gather -> DGEMM -> scatter per panel
    if (sym) {
      allow_smem(build_Dsym_peer_kernel, smem);
      build_Dsym_peer_kernel<<<(unsigned)(w->W / 256), 256, smem, st>>>(pin, in_dev, w->d_D[0], w->W, j0, len, w->d_etab, n, sp->d_strA,
                                                                        sp->d_strB, sp->d_rankA, sp->d_rankB, sp->NB, sp->row_begin);
      SQ_CHECK(launch_error("build_Dsym_peer_kernel"));
    } else {
      SQ_CHECK(launch_build_D(sp, w, in_dev, w->d_D[0], j0, st, false, false, &pin));
    }
    SQ_CHECK(sq_sigma_gemm(d_G, ldg, w->d_D[0], w->d_F[0], nrow, w->W, st));
    scatter_E_peer_kernel<<<(unsigned)(w->W / 256), 256, smem, st>>>(pout, in_dev, out_dev, w->d_F[0], d_k, w->d_frow, w->W, j0, len,
                                                                     w->d_etab, n2, sp->d_strA, sp->d_strB, sp->d_rankA, sp->d_rankB, sp->NB,
                                                                     sp->row_begin);
    SQ_CHECK(launch_error("scatter_E_peer_kernel"));
  }
  return SQ_OK;
}

// Host check of erec_closed against the table get_work builds from sq_make_string_action (works on host-only spaces).
extern "C" int sq_debug_etab_closed_form(const sq_space* sp, int* n_mismatch) {
  if (!sp || !n_mismatch) return SQ_ERR_INVALID;
  const int n = sp->n_orb;
  int bad = 0;
  for (int p = 0; p < n; ++p)
    for (int q = 0; q < n; ++q)
      for (int spin = 0; spin < 2; ++spin) {
        int32_t label[2] = {2 * (2 * p + spin) + 1, 2 * (2 * q + spin)};
        StringAction a;
        SQ_CHECK(sq_make_string_action(sp, label, 2, &a));
        ERec r;
        if (spin == 0)
          r = {a.toccA, a.tempA, a.occA, a.empA, a.flipA, a.parA, a.parB, a.s0};
        else
          r = {a.toccB, a.tempB, a.occB, a.empB, a.flipB, a.parB, a.parA, a.s0};
        const ERec c = erec_closed(p, q, spin);
        if (r.tocc != c.tocc || r.temp != c.temp || r.occ != c.occ || r.emp != c.emp || r.flip != c.flip || r.parS != c.parS ||
            r.parO != c.parO || r.s0 != c.s0)
          ++bad;
      }
  *n_mismatch = bad;
  return SQ_OK;
}
