// Internal structures of libsqsv (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <array>
#include <atomic>
#include <map>
#include <string>
#include <vector>

#include "sqsv.h"

#define SQ_MAX_ORB 32          // string occupation masks are uint32_t
#define SQ_MAX_STRING_OPS 32   // ladder operators per normal-ordered string
#define SQ_MAX_PROGRAM 8       // fused rotation steps per tile launch

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost nothing unless a profiler is attached

// NVTX range per kernel family / API call (tracing): visible in nsys / ncu --nvtx timelines
struct SqRange {
  explicit SqRange(const char* name) { nvtxRangePushA(name); }
  ~SqRange() { nvtxRangePop(); }
  SqRange(const SqRange&) = delete;
  SqRange& operator=(const SqRange&) = delete;
};

extern std::atomic<int64_t> g_sq_launches;
void sq_set_error(const char* fmt, ...);

#define SQ_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      sq_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
      return SQ_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

#define SQ_CHECK(st)            \
  do {                          \
    int s_ = (st);              \
    if (s_ != SQ_OK) return s_; \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Action of ONE normal-ordered ladder-operator string on a determinant |A,B> (A/B = alpha/beta
// occupation masks, bit o = spatial orbital o), in *source* form:
//   valid(src)  <=>  (A & occA)==occA && (A & empA)==0 && (B & occB)==occB && (B & empB)==0
//   tgt         =    (A ^ flipA, B ^ flipB)
//   sign        =    s0 * (-1)^{popc(A & parA) + popc(B & parB)}
// This closed form is equivalent to the reference's sequential bit-flip/popcount loop
// (operator_state_algebra.py:118-135): the phase contributed by ladder operator k is
// popc((src ^ F_k) & below_k), and parities of popcounts add as XOR of masks.
// ---------------------------------------------------------------------------------------------
struct StringAction {
  uint32_t occA, empA, flipA, parA;
  uint32_t occB, empB, flipB, parB;
  int s0;        // +1 / -1
  bool conserving;  // keeps N_alpha and N_beta (otherwise the string leaves the CI space)
  // target-form screens (for the gather kernel): tgt must have tocc* set and temp* clear
  uint32_t toccA, tempA, toccB, tempB;
};

// packed per-string code used by the tile kernels (one uint32 per alpha / beta string):
//   bits 1:0  class: 0 inert, 1 src (i occupied, a empty), 2 tgt
//   bit  2    same-spin single-excitation sign negative         (src only)
//   bit  3    cross sign negative: factor this string gives the other spin's single (all strings)
//   bit  4    pair-double sign factor negative                  (src only)
//   bits 31:5 partner string index                              (src only)
#define SQ_CLS_INERT 0u
#define SQ_CLS_SRC 1u
#define SQ_CLS_TGT 2u

struct PairTables {       // tables for one spatial orbital pair (i,a)
  int i, a;
  uint32_t* d_codeA = nullptr;   // [NA]
  uint32_t* d_codeB = nullptr;   // [NB]
  std::vector<uint32_t> h_codeA, h_codeB;   // host copies (quad work lists are derived from them)
  int32_t* d_rowsA = nullptr;    // alpha rows that are src or inert (tgt rows ride with their src)
  int64_t n_rows = 0;            // local row items
  int64_t n_src_rows = 0, n_src_cols = 0;
  int64_t touched = 0;           // amplitudes a full sa_single(+double) block touches (local rows)
  // ---- compacted, class-homogeneous work lists for tile_kernel_v2 ----
  // column items {ib, ibp | flags<<27}: [src columns, padded to a CTA multiple][inert columns, padded];
  // pad entries have ib = -1.  flags: bit0 same-spin sign neg, bit1 cross neg, bit2 partner cross neg,
  // bit3 pair-double factor neg.
  int2* d_colItems = nullptr;
  int n_colblk_src = 0, n_colblk_inert = 0;
  // row items {ia, iap, flags, 0}: [src rows padded to a TILE_ROWS multiple][inert rows padded]; pad ia = -1.
  int4* d_rowItems = nullptr;
  int n_rowchunk_src = 0, n_rowchunk_inert = 0;
  int sigma = 0;                 // gauge-invariant pair-double sign (+1/-1) if uniform, 0 otherwise
  bool blocked = false;          // the pair touches a constrained alpha orbital of the space: no tables, never launched
  bool cross_global = false;     // some row pair of this orbital pair spans two devices (same answer on every rank)
  int64_t n_cross_items = 0;     // cross-device row pairs this rank works on (half of the columns each)
};

// Work lists for TWO commuting bricks (disjoint orbital pairs P1, P2) applied in one sweep.  A string is
// active (src/tgt) or inert in each pair; its group holds 1, 2 or 4 strings (index a = a1*2 + a2, a_k = 0 src /
// 1 tgt in pair k).  Lists are sorted by group type t = 2*active(P1) + active(P2) in the order 3,2,1,0 and
// padded per type, so every CTA is type-homogeneous in rows and in columns.
struct QuadTables {
  bool ok = false;               // false: flags are not pair-local for this combination -> no quad fusion
  int pair1 = -1, pair2 = -1;
  int4* d_colIdx = nullptr;      // 4 string indices per column group (-1 = absent / pad)
  int* d_colFlags = nullptr;     // bits 0-2: sSb1, crb1(b1=0 or inert), crbp1 ; bits 4-6: same for pair 2
  int4* d_rowIdx = nullptr;      // row indices relative to the shard start
  int* d_rowFlags = nullptr;     // bits 0-2: sSa1, cra1, crap1 ; bits 4-6: pair 2
  int colblk_end[4] = {0, 0, 0, 0};    // cumulative CTA counts after types 3,2,1,0
  int rowchunk_end[4] = {0, 0, 0, 0};
  int64_t touched = 0;
};

#define SQ_MAX_WORLD 16
struct PeerPtrs { double* p[SQ_MAX_WORLD]; };   // base pointer of the vector shard on every rank (peer-mapped)

struct GenTables {        // tables for one generic excitation generator G (single string)
  StringAction act;
  int32_t* d_srcRows = nullptr;  // alpha strings valid as source (local rows)
  int32_t* d_tgtRows = nullptr;  // their targets
  int8_t* d_sgnRows = nullptr;   // alpha sign factor (includes s0)
  int64_t n_rows = 0;
  int32_t* d_colCode = nullptr;  // [NB]: (partner<<1 | neg) for valid beta sources, -1 otherwise
  int64_t n_cols_valid = 0;
};

struct sq_space {
  int n_orb, n_alpha, n_beta, device;
  int64_t NA, NB, ndet;
  int64_t row_begin, row_end;       // local alpha rows
  // constrained alpha list (sq_space_create_constrained): only the strings with (mask & alpha_cmask) == alpha_cpat, in the
  // order they have in the full list.  Operators that move an alpha electron on a constrained orbital are "blocked".
  uint32_t alpha_cmask = 0, alpha_cpat = 0;
  int world = 1, rank = 0;          // alpha-row partition over devices (sq_space_set_partition)
  std::vector<int64_t> row_starts;  // [world + 1]; rank r owns rows [row_starts[r], row_starts[r+1])
  unsigned long long* d_peer_tab = nullptr;             // device table of shard base pointers (cross-device tiles)
  unsigned long long peer_shadow[16] = {0};             // last uploaded table
  uint64_t binom[SQ_MAX_ORB + 2][SQ_MAX_ORB + 2];
  std::vector<uint32_t> strA, strB; // occupation masks in itertools.combinations order
  std::vector<int32_t> rankA, rankB;  // mask -> string index (-1 if wrong electron count); 2^n entries
  uint32_t *d_strA = nullptr, *d_strB = nullptr;
  uint32_t* d_gwordB = nullptr;     // gauge words of the beta strings (sqsv_win.cu), built lazily
  int32_t *d_rankA = nullptr, *d_rankB = nullptr;
  // reduction scratch
  double* d_partial = nullptr;
  int64_t n_partial = 0;
  double* h_pinned = nullptr;       // small pinned staging buffer
  // lazily allocated full-vector work buffers (sa_double polynomial, sigma)
  double* d_work[3] = {nullptr, nullptr, nullptr};
  int64_t local_len() const { return (row_end - row_begin) * NB; }
};

struct GenOp {   // generic multi-string generator (sa_double_*)
  std::vector<StringAction> strings;
  std::vector<double> coeffs;
};

struct LayoutOp {
  int type;
  std::vector<int> idx;
  int pair = -1;      // index into sq_layout::pairs (sa_single / pair-double)
  bool pair_double = false;
  int gen = -1;       // index into sq_layout::gens (generic single-string generator)
  bool null_op = false;  // generator vanishes identically
  bool blocked = false;  // moves an alpha electron on a constrained orbital of the space (sq_space_create_constrained)
  GenOp* multi = nullptr;
};

// one kernel launch of the plan of sq_ups_apply (sqsv_api.cu)
struct Launch {
  int kind = 0;                 // 0 single run (tile / generic / sa_double / null), 1 quad (2 runs), 2 window sweep
  std::vector<int> runs;        // run indices in execution order
  const struct WinTables* wt = nullptr;
  const struct QuadTables* qt = nullptr;
};
struct PlanCache {
  std::vector<std::vector<int>> runs;
  std::vector<Launch> launches;
};

struct sq_layout {
  sq_space* sp;
  std::vector<LayoutOp> ops;
  std::vector<PairTables> pairs;
  std::map<std::pair<int, int>, int> pair_index;
  std::vector<GenTables> gens;
  std::map<std::vector<int>, int> gen_index;
  std::map<std::pair<int, int>, QuadTables> quads;   // built lazily per (pair1, pair2)
  std::vector<PlanCache> plans;   // launch plans by run structure (sqsv_api.cu)
  std::vector<PlanCache> grad_plans;   // the same for the gradient sweep (its own window configuration)
  int plan_version = 0;
  std::map<std::array<int, 5>, struct WinTables*> wins;   // built lazily per (w0, H) (sqsv_win.cu)
};

static inline int sq_row_owner(const sq_space* sp, int64_t row) {
  if (sp->world <= 1) return sp->rank;
  int r = 0;
  while (r + 1 < sp->world && row >= sp->row_starts[r + 1]) ++r;
  return r;
}
static inline int64_t sq_rank_start(const sq_space* sp, int r) {
  return sp->world <= 1 ? sp->row_begin : sp->row_starts[r];
}

// host helpers (sqsv_space.cu)
int sq_rank_mask(const sq_space* sp, int spin, uint32_t mask);   // -1 if not in list
int sq_make_string_action(const sq_space* sp, const int32_t* ops, int n_ops, StringAction* out);
void sq_hamiltonian_release(const sq_space* sp);   // frees sigma / RDM panels (sqsv_hamiltonian.cu)
void sq_hamiltonian_set_panel_width(long long w);  // determinants per panel for spaces that build their panels later (0: 1 GiB)
void sq_hamiltonian_set_rows_cfg(int threads, int ch);  // CTA size / column chunks per row of the row kernels
void sq_hamiltonian_set_rows_mode(int on);         // row-per-CTA panel kernels (default on) or determinant-per-thread
void sq_hamiltonian_set_pipeline(int on);          // sigma / RDM panel pipeline over internal streams (default on)
void sq_hamiltonian_set_etab_alu(int on);          // panel kernels without an E_pq table (records computed from (p,q); default off)
void sq_hamiltonian_set_sigma_spinsym(int on);     // half sigma build for spin-flip symmetric vectors (default on)
void sq_hamiltonian_set_spinsym_blk(int on);       // 32 x 32 blocked panels of the half build (default on); off: determinant-per-thread kernels
void sq_hamiltonian_set_etab_tab(int on);          // panel kernels on per-string partner tables
void sq_hamiltonian_set_rdm_tri(int on);
void sq_hamiltonian_set_rdm_sym(int on);           // 2-RDM of one vector from the symmetric S / A Gram matrices (default on)
void sq_hamiltonian_set_sigma_fused(int on);       // sigma through the fused gather -> DMMA -> scatter kernel (default on)           // RDMs with bra == ket: three half-size DGEMMs instead of one (default off)
void sq_hamiltonian_set_etab_mode(int use_const);  // E_pq table in constant (1) or shared (0) memory
void sq_reshard_set_mode(int lsu);                 // re-shard kernel: 0 bulk-copy engine (default), 1 vector load/store
#ifdef __CUDACC__
// fp64 tensor-core MMA and async-copy primitives shared by sqsv_dmma.cu and the fused sigma kernel (sqsv_hamiltonian.cu).
// Fragment layout of mma.m8n8k4.f64: A (8x4, row): a0 = A[lane >> 2][lane & 3]; B (4x8, col): b0 = B[lane & 3][lane >> 2];
// C/D (8x8): c{0,1} = C[lane >> 2][2 * (lane & 3) + {0,1}].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, int src_bytes) {
  // 16-byte async copy; src_bytes = 0 zero-fills the destination (rows / columns beyond the matrix)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// hand-written fp64 tensor-core contractions (sqsv_dmma.cu)
#define SQ_GRAM_MAXT 40
struct GramTiles {          // 128 x 128 output tiles of one Gram matrix (upper triangle for bra == ket)
  int n;
  int ta[SQ_GRAM_MAXT], tb[SQ_GRAM_MAXT];
};
int sq_gram_begin(int nrows, bool symmetric, int n_sm, double** d_partial, size_t* partial_doubles, GramTiles* tiles, int* n_split,
                  cudaStream_t st);
int sq_gram_panel(const double* X, const double* Y, int64_t ld, int nrows, int64_t K, const GramTiles& tiles, int n_split,
                  double* d_partial, cudaStream_t st);
int sq_gram_end(const GramTiles& tiles, int n_split, const double* d_partial, int nrows, bool symmetric, double* d_G2, cudaStream_t st);
int sq_gram_sym_rows();
int sq_gram_sym_begin(int n_sm, double** d_partial, size_t* partial_doubles, int* n_split, cudaStream_t st);
int sq_gram_sym_panel(const double* Z, int64_t ld, int rows0, int rows1, int64_t K, int n_split, double* d_partial, cudaStream_t st);
int sq_gram_sym_end(int n_split, const double* d_partial, double* d_G, cudaStream_t st);
void sq_sigma_gemm_set_residency(int n);
void sq_sigma_gemm_set_row_parts(int n);   // 2 (default): four warps per sigma GEMM CTA, 3: six
int sq_sigma_gemm(const double* Gm, int ldg, const double* D, double* F, int nrow, int64_t W, cudaStream_t st);
int sq_panel_gemv(const double* D, int64_t W, int nrows, const double* x, int64_t K, double* g1, cudaStream_t st);
int sq_ensure_work(sq_space* sp, int which);
int sq_ensure_partial(sq_space* sp, int64_t n);

// kernel launchers (sqsv_kernels.cu)
struct TileStep { int kind; double c, s; };   // kind: 0 alpha-rot, 1 beta-rot, 2 pair-double rot
// one brick (fused program on one orbital pair) in the gauge-fixed tile basis (x00, x01, x10, x11)
struct TileMatrices {
  double m[16];        // 4x4, row-major
  double ca, sa;       // total alpha-single rotation (src row x inert column)
  double cb, sb;       // total beta-single rotation (inert row x src column)
};
void sq_build_tile_matrices(const TileStep* steps, int n_steps, int sigma, TileMatrices* tm);
// same with explicit signs of the alpha single, beta single and pair double generators
void sq_build_tile_matrices3(const TileStep* steps, int n_steps, int ea, int eb, int ed, TileMatrices* tm);

// Tables of the window kernel (sqsv_win.cu): one orbital window [w0, w0+H), see the header of that file.
#define SQ_WIN_MAX_BRICKS 16
struct WinTables {
  bool ok = false;
  int w0 = 0, H = 0;
  int LTA = 0, LTB = 0, max_a = 0, max_b = 0, lanes_j = 0, gp = 16;
  int n_groups_a = 0, n_chunks_b = 0, n_ranges_b = 0, nbuf = 1;
  int maxQ = 0, maxS = 0, tile_doubles = 0;
  int2 *d_groupsA = nullptr, *d_clsA = nullptr, *d_chunksB = nullptr, *d_clsB = nullptr, *d_rangesB = nullptr;
  int *d_deltaA = nullptr, *d_deltaB = nullptr, *d_gbaseB = nullptr, *d_rchunks = nullptr;
  uint32_t* d_lists = nullptr;
  int4* d_listidx = nullptr;
  std::vector<int> pair_local;     // layout pair index -> pair id inside the window tables, -1 if unusable
  std::vector<int8_t> eps;         // [3 * local pair] constant signs of Ta, Tb, pair double in the window gauge
  std::vector<int> pair_lo;        // [local pair] lower orbital of the pair, relative to w0
  std::vector<char> pair_flip;     // [local pair] 1 if the pair's source orbital i is the upper one
  size_t smem = 0;                 // dynamic shared memory per CTA with SQ_WIN_MAX_BRICKS bricks
  int64_t touched = 0;             // amplitudes per launch (the whole local vector)
  struct Win3Tables* w3 = nullptr; // tables of win3_kernel (sqsv_win3.cu); nullptr / !ok: the launch uses win_kernel
};
// host tables of one spin of a window (built by build_side, sqsv_win.cu)
struct SideHost {
  std::vector<int2> groups;          // alpha: {first row, class}; beta: {chunk id, class | tiles << 16}
  std::vector<int> gbase;            // beta: [chunk][WIN_G]
  std::vector<int2> cls;             // {count, e_w}
  std::vector<int> delta;
  std::vector<std::vector<uint32_t>> wl;   // window parts per electron count, combination order
  int ncls = 0, LT = 0, max_cnt = 0;
};

// ---- window kernel, version 3 (sqsv_win3.cu): register blocks over THREE orbitals, merged small tiles ----
#define W3_MAXLISTS 6      // distinct orbital triples per launch (H - 2 for windows of H <= 8 orbitals)
#define W3_MAXSTEPS 16
struct WinBrick {
  double m[16];            // 4x4 on (x[r][c], x[r][c'], x[r'][c], x[r'][c']), row-major
  double ca, sa, cb, sb;   // alpha single on (x[r][c], x[r'][c]); beta single on (x[r][c], x[r][c'])
};
struct Win3Program {
  int n_steps, n_lists;
  int list_t0[W3_MAXLISTS];                  // first window orbital of the triple of list l
  unsigned char step_list[W3_MAXSTEPS];      // list of step s
  unsigned char step_first[W3_MAXSTEPS + 1]; // bricks [step_first[s], step_first[s+1]) of `br` run in step s
  unsigned char brick_lp[SQ_WIN_MAX_BRICKS]; // 0: the brick sits on the lower two orbitals of its triple, 1: on the upper two
  WinBrick brv[SQ_WIN_MAX_BRICKS][8];        // per brick: one matrix set per item type (orientation of hole groups, row / column items)
};
struct Win3Tables {
  bool ok = false;
  int H = 0, LTA = 0, LTB = 0, gp = 16, lanes_j = 0;
  int lmax = 0;          // largest expanded item list of one (work item, triple)
  int max_rows = 0, max_cols = 0, max_chunks = 0, tile_doubles = 0;
  // host mirrors (launch bookkeeping and the host emulation used by the CPU tests)
  std::vector<int2> agroups;   // class-major: {first row (shard-relative), class}
  std::vector<int2> acls;      // {rows of a tile, e_w}
  std::vector<int> adelta;     // [class][LTA]
  std::vector<int2> bchunks;   // class-major: {chunk id, class | tiles << 16}
  std::vector<int> bgbase;     // [chunk id][16] first column of every tile
  std::vector<int2> bcls;
  std::vector<int> bdelta;
  std::vector<int4> work;      // one CTA: {a_first, a_cnt | Ka << 16, b_first, n_chunks | Kb << 16}
  std::vector<uint2> items;    // x: r0 | r1 << 8 | r2 << 16 | type << 24, y: c0 | c1 << 8 | c2 << 16 (tile-local string ranks)
  std::vector<int4> itemidx;   // [t0][e_wa][e_wb] {offset, items, counts of types 0..3 (bytes), counts of types 4..7}
  int2 *d_agroups = nullptr, *d_acls = nullptr, *d_bchunks = nullptr, *d_bcls = nullptr;
  int *d_adelta = nullptr, *d_bgbase = nullptr, *d_bdelta = nullptr;
  int4 *d_work = nullptr, *d_itemidx = nullptr;
  uint2* d_items = nullptr;
};
int sq_build_win3(sq_space* sp, struct WinTables* wt, const SideHost& hA, const SideHost& hB);
void sq_free_win3(Win3Tables* w3);
int sq_win3_program(const struct WinTables& wt, const int* pair_idx, const struct TileStep* const* steps, const int* n_steps, int n_bricks,
                    Win3Program* P);
int sq_launch_win3(sq_space* sp, const struct WinTables& wt, const Win3Program& P, double* state, cudaStream_t st, int n_states,
                   int64_t state_stride);
int sq_win3_emulate_host(const sq_space* sp, const struct WinTables& wt, const Win3Program& P, double* host_state);
void sq_gauge_host(const sq_space* sp, double* host_state);
void sq_win3_print_stats(const struct WinTables& wt, const int* pair_idx, int n_bricks);
void sq_win3_set_enabled(int on);
bool sq_win3_enabled();

int sq_win_max_class(int n, int ne, int w0, int H);
size_t sq_win_smem_bytes(int max_a, int max_b, int gp, int nbuf, int lta, int ltb, int maxQ, int maxS, int n_bricks);
int sq_launch_gauge(sq_space* sp, double* state, cudaStream_t st, int n_states = 1, int64_t state_stride = 0);
bool sq_win_pair_ok(const sq_layout* lay, int pair, int w0, int H);
int sq_get_win(sq_space* sp, sq_layout* lay, int w0, int H, const WinTables** out);
void sq_free_win_tables(WinTables* wt);
int sq_win_grad_replicas();
int sq_launch_win_grad(sq_space* sp, const WinTables& wt, const int* pair_idx, const TileStep* const* steps, const int* n_steps,
                       const int* slot0, int n_bricks, double* bra, double* ket, double* d_out, int n_out, cudaStream_t st);
int sq_launch_win(sq_space* sp, const WinTables& wt, const int* pair_idx, const TileStep* const* steps, const int* n_steps,
                  int n_bricks, double* state, cudaStream_t st, int n_states = 1, int64_t state_stride = 0);
int sq_launch_tile(sq_space* sp, const PairTables& pt, const TileStep* steps, int n_steps,
                   double* state, const PeerPtrs* peers, cudaStream_t st);
int sq_launch_quad(sq_space* sp, const QuadTables& qt, const TileStep* steps1, int n1, int sigma1,
                   const TileStep* steps2, int n2, int sigma2, double* state, cudaStream_t st);
int sq_launch_quad_grad(sq_space* sp, const QuadTables& qt, const TileStep* steps1, int n1, int sigma1, const TileStep* steps2,
                        int n2, int sigma2, double* bra, double* ket, double* d_out1, double* d_out2, cudaStream_t st);
int sq_reduce_partials(const double* partial, int64_t nblocks, int ns, int n_out, double* out, cudaStream_t st);
int sq_launch_tile_grad(sq_space* sp, const PairTables& pt, const TileStep* steps, int n_steps,
                        double* bra, double* ket, double* grad_out_host, cudaStream_t st);
int sq_launch_tile_grad_peer(sq_space* sp, const PairTables& pt, const TileStep* steps, int n_steps, double* bra, double* ket,
                             const unsigned long long* d_tab_bra, const unsigned long long* d_tab_ket, double* d_out,
                             cudaStream_t st);
int sq_launch_gen_rot(sq_space* sp, const GenTables& gt, double c, double s, double* state,
                      cudaStream_t st);
int sq_launch_gen_apply(sq_space* sp, const GenTables& gt, const double* in, double* out,
                        cudaStream_t st);
int sq_launch_gen_grad(sq_space* sp, const GenTables& gt, double c, double s, double* bra, double* ket,
                       double* grad_out_host, cudaStream_t st);
int sq_launch_gather(sq_space* sp, const std::vector<StringAction>& strings,
                     const std::vector<double>& coeffs, const double* in, double* out, int accumulate,
                     cudaStream_t st);
int sq_launch_dot(sq_space* sp, const double* a, const double* b, double* out_host, cudaStream_t st);
int sq_launch_axpy(sq_space* sp, double alpha, const double* x, double* y, cudaStream_t st);
int sq_launch_scale_copy(sq_space* sp, double alpha, const double* x, double* y, cudaStream_t st);
