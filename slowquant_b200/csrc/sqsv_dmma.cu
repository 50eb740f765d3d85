// Hand-written fp64 tensor-core (DMMA) contractions of the sigma / RDM path -- the only GEMM-shaped work of the engine
// (north_star: "the strided-to-dense 2-RDM contraction ... is the only step allowed onto fp64 tensor cores").  tcgen05 has no
// fp64, so on sm_100a the fp64 tensor path is mma.sync.aligned.m8n8k4.f64 (DMMA in the SASS); operands are staged in shared
// memory by multi-stage cp.async pipelines, accumulators live in registers.  These kernels replace the cuBLAS DGEMM calls of
// round 1 (cuBLAS dispatched an sm_80 CUTLASS kernel for these shapes).
//
//   gram_dmma_kernel   G2[a][b] += sum_t X[a][t] Y[b][t]        (a, b < n^2 = 256, t over the W determinants of a panel)
//                      reference: the <E_pq E_rs> loops of ups_wavefunction.py:432-476.  128 x 128 output tiles, split-K over
//                      the CTAs of the grid; every (tile, split) owns one slot of a persistent partial-sum buffer that it
//                      updates in panel order, and one reduction kernel adds the slots in a fixed order at the end: the
//                      summation order is deterministic (no atomics).  For bra == ket only the upper-triangular tiles run.
//   sigma_dmma_kernel  F[pq][t] = sum_rs Gm[pq][rs] D[rs][t]    (pq, rs < 136 symmetrised generators, or n^2)
//                      reference: the two-body part of hamiltonian_0i_0a applied string by string (operators.py:476-529,
//                      operator_state_algebra.py:596-628).  One CTA (4 warps) = all rows x 64 determinants; the integral matrix
//                      streams through shared memory in chunks of 16 columns.
//
// Fragment layout of mma.m8n8k4.f64 (PTX ISA): A (8x4, row): a0 = A[lane >> 2][lane & 3]; B (4x8, col): b0 = B[lane & 3][lane >> 2];
// C/D (8x8): c{0,1} = C[lane >> 2][2 * (lane & 3) + {0,1}].
#include <cstdio>
#include <cstdlib>

#include "sqsv_internal.h"

// ------------------------------------------------------------------------------------------------------------------
// Gram matrix: 128 x 128 tile of G2 per CTA, K range [k_begin, k_end) of the panel (multiples of GR_KC).
// ------------------------------------------------------------------------------------------------------------------
#define GR_BM 128
#define GR_KC 16
#define GR_LD 20            // shared-memory row stride in doubles: = 4 (mod 16) -> conflict-free fragment loads
#define GR_STAGES 4
#define GR_THREADS 256


__global__ void __launch_bounds__(GR_THREADS, 1)
gram_dmma_kernel(const double* __restrict__ X, const double* __restrict__ Y, int64_t ld, int nrows, int64_t K, int n_split,
                 const GramTiles tiles, double* __restrict__ partial) {
  extern __shared__ __align__(16) double gsm[];
  const int tile = blockIdx.x / n_split, split = blockIdx.x % n_split;
  const int a0 = tiles.ta[tile] * GR_BM, b0 = tiles.tb[tile] * GR_BM;
  const int64_t n_chunks = K / GR_KC;
  const int64_t c_begin = n_chunks * split / n_split, c_end = n_chunks * (split + 1) / n_split;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 2, wn = warp & 3;          // 2 x 4 warps: 64 x 32 outputs each
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(gsm);
  constexpr int STAGE_D = 2 * GR_BM * GR_LD;         // doubles per stage (X tile, then Y tile)
  // one stage: 2 x 128 rows x 16 doubles = 2 x 1024 16-byte chunks; thread t copies chunks t, t + 256, ...
  auto issue = [&](int64_t c, int slot) {
    const int64_t k0 = c * GR_KC;
    const uint32_t sb = sbase + (uint32_t)(slot * STAGE_D) * 8u;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int id = threadIdx.x + q * GR_THREADS;   // 0 .. 2047
      const int which = id >> 10, r = (id >> 3) & 127, ch = id & 7;
      const int row = (which ? b0 : a0) + r;
      const double* src = (which ? Y : X) + (int64_t)(row < nrows ? row : 0) * ld + k0 + ch * 2;
      cp16(sb + (uint32_t)(which * GR_BM * GR_LD + r * GR_LD + ch * 2) * 8u, src, row < nrows ? 16 : 0);
    }
  };
  const int64_t n_it = c_end - c_begin;
  for (int s = 0; s < GR_STAGES - 1; ++s) {
    if (s < n_it) issue(c_begin + s, s);
    cp_commit();
  }
  for (int64_t it = 0; it < n_it; ++it) {
    cp_wait<GR_STAGES - 2>();
    __syncthreads();                                  // stage `it` has landed for everybody; slot (it - 1) is free
    if (it + GR_STAGES - 1 < n_it) issue(c_begin + it + GR_STAGES - 1, (int)((it + GR_STAGES - 1) % GR_STAGES));
    cp_commit();
    const double* xs = gsm + (it % GR_STAGES) * STAGE_D;
    const double* ys = xs + GR_BM * GR_LD;
#pragma unroll
    for (int k4 = 0; k4 < GR_KC / 4; ++k4) {
      double af[8], bf[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) af[i] = xs[(wm * 64 + i * 8 + (lane >> 2)) * GR_LD + k4 * 4 + (lane & 3)];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = ys[(wn * 32 + j * 8 + (lane >> 2)) * GR_LD + k4 * 4 + (lane & 3)];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  // this (tile, split) slot is owned by exactly one CTA per launch and launches are stream-ordered: plain read-modify-write
  double* out = partial + ((size_t)split * SQ_GRAM_MAXT + tile) * (GR_BM * GR_BM);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = wm * 64 + i * 8 + (lane >> 2), c = wn * 32 + j * 8 + 2 * (lane & 3);
      double2* p = reinterpret_cast<double2*>(out + r * GR_BM + c);
      double2 v = *p;
      v.x += acc[i][j][0];
      v.y += acc[i][j][1];
      *p = v;
    }
}

// G2[a][b] (row-major, leading dimension nrows) = sum over the splits, in split order; mirrored tiles are filled from their
// transposes (bra == ket).
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const double* __restrict__ partial, int n_split, const GramTiles tiles, int nrows, int mirror,
                   double* __restrict__ G2) {
  const int tile = blockIdx.y;
  const int e = blockIdx.x * 256 + threadIdx.x;       // element of the 128 x 128 tile
  const int r = e / GR_BM, c = e % GR_BM;
  const int a = tiles.ta[tile] * GR_BM + r, b = tiles.tb[tile] * GR_BM + c;
  if (a >= nrows || b >= nrows) return;
  double s = 0.0;
  for (int k = 0; k < n_split; ++k) s += partial[((size_t)k * SQ_GRAM_MAXT + tile) * (GR_BM * GR_BM) + e];
  G2[(size_t)a * nrows + b] = s;
  if (mirror && tiles.ta[tile] != tiles.tb[tile]) G2[(size_t)b * nrows + a] = s;
}

static size_t gram_smem() { return sizeof(double) * GR_STAGES * 2 * GR_BM * GR_LD; }

// begin an accumulation: tiles of the nrows x nrows Gram matrix, zeroed partial sums
int sq_gram_begin(int nrows, bool symmetric, int n_sm, double** d_partial, size_t* partial_doubles, GramTiles* tiles, int* n_split,
                  cudaStream_t st) {
  const int nt = (nrows + GR_BM - 1) / GR_BM;
  tiles->n = 0;
  for (int a = 0; a < nt; ++a)
    for (int b = symmetric ? a : 0; b < nt; ++b) {
      if (tiles->n >= SQ_GRAM_MAXT) {
        sq_set_error("Gram matrix of %d rows needs more than %d output tiles", nrows, SQ_GRAM_MAXT);
        return SQ_ERR_UNSUPPORTED;
      }
      tiles->ta[tiles->n] = a;
      tiles->tb[tiles->n] = b;
      ++tiles->n;
    }
  *n_split = std::max(1, n_sm / tiles->n);             // one CTA per SM (one wave)
  const size_t need = (size_t)(*n_split) * SQ_GRAM_MAXT * GR_BM * GR_BM;
  if (*partial_doubles < need) {
    if (*d_partial) cudaFree(*d_partial);
    *d_partial = nullptr;
    *partial_doubles = 0;
    SQ_CUDA(cudaMalloc(d_partial, sizeof(double) * need));
    *partial_doubles = need;
  }
  SQ_CUDA(cudaMemsetAsync(*d_partial, 0, sizeof(double) * need, st));
  return SQ_OK;
}

// one panel: partial[tile, split] += X[a-tile rows][k range of the split] . Y[b-tile rows][same]^T
int sq_gram_panel(const double* X, const double* Y, int64_t ld, int nrows, int64_t K, const GramTiles& tiles, int n_split,
                  double* d_partial, cudaStream_t st) {
  if (K % GR_KC != 0 || ld % 2 != 0) {
    sq_set_error("Gram panel: K = %lld must be a multiple of %d and the leading dimension even", (long long)K, GR_KC);
    return SQ_ERR_INVALID;
  }
  static bool attr = false;
  if (!attr) {
    SQ_CUDA(cudaFuncSetAttribute(gram_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gram_smem()));
    attr = true;
  }
  gram_dmma_kernel<<<(unsigned)(tiles.n * n_split), GR_THREADS, gram_smem(), st>>>(X, Y, ld, nrows, K, n_split, tiles, d_partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sq_set_error("gram_dmma_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}

int sq_gram_end(const GramTiles& tiles, int n_split, const double* d_partial, int nrows, bool symmetric, double* d_G2, cudaStream_t st) {
  const dim3 grid(GR_BM * GR_BM / 256, (unsigned)tiles.n);
  gram_reduce_kernel<<<grid, 256, 0, st>>>(d_partial, n_split, tiles, nrows, symmetric ? 1 : 0, d_G2);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sq_set_error("gram_reduce_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Symmetric Gram matrices for the 2-RDM of one real vector (bra == ket): with S_pq = E_pq + E_qp (p >= q; S_pp = E_pp) and
// A_pq = E_pq - E_qp (p > q) the products <S A> and <A S> are commutators, i.e. 1-RDM elements, so
//     <E_pq E_rs> = sum T T' ( <S_x S_y> | -<A_x A_y> | 1/2 <[Z_x, Z_y]> )
// needs only the two SYMMETRIC Gram matrices G_SS = Ds Ds^T (136 x 136 at n = 16) and G_AA = Da Da^T (120 x 120): a quarter of
// the n^4 products of the plain <E_pq E_rs> Gram matrix (the reference's s_lim loops, ups_wavefunction.py:450-475, exploit the
// same 4-fold symmetry).  One CTA = one matrix (<= 144 rows) and one K range; 12 warps own the 48 x 24 blocks of its upper
// triangle, operands staged once per stage for rows and columns alike (X = Y).  Split-K partial sums live in per-(matrix, split)
// slots updated in panel order and are added in a fixed order at the end (deterministic).
// ------------------------------------------------------------------------------------------------------------------
#define GS_R 144
#define GS_KC 16
#define GS_LD 20
#define GS_STAGES 4
#define GS_THREADS 384

__global__ void __launch_bounds__(GS_THREADS, 1)
gram_sym_kernel(const double* __restrict__ Z, int64_t ld, int rows0, int rows1, int64_t K, int n_split, double* __restrict__ partial) {
  extern __shared__ __align__(16) double gsm2[];
  const int mat = blockIdx.x / n_split, split = blockIdx.x % n_split;
  const int row0 = mat ? rows0 : 0, nrows = mat ? rows1 : rows0;
  const int64_t n_chunks = K / GS_KC;
  const int64_t c_begin = n_chunks * split / n_split, c_end = n_chunks * (split + 1) / n_split;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // warp -> (row block of 48, column block of 24) of the upper triangle: (0, 0..5), (1, 2..5), (2, 4..5)
  const int bi = warp < 6 ? 0 : (warp < 10 ? 1 : 2);
  const int bj = warp < 6 ? warp : (warp < 10 ? warp - 4 : warp - 6);
  double acc[6][3][2];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(gsm2);
  constexpr int STAGE_D = GS_R * GS_LD;
  auto issue = [&](int64_t c, int slot) {
    const int64_t k0 = c * GS_KC;
    const uint32_t sb = sbase + (uint32_t)(slot * STAGE_D) * 8u;
#pragma unroll
    for (int q = 0; q < 3; ++q) {                      // 144 rows x 8 chunks of 16 bytes = 1152 = 3 x 384
      const int id = threadIdx.x + q * GS_THREADS;
      const int r = id >> 3, ch = id & 7;
      const bool ok = r < nrows;
      cp16(sb + (uint32_t)(r * GS_LD + ch * 2) * 8u, Z + (int64_t)(row0 + (ok ? r : 0)) * ld + k0 + ch * 2, ok ? 16 : 0);
    }
  };
  const int64_t n_it = c_end - c_begin;
  for (int s = 0; s < GS_STAGES - 1; ++s) {
    if (s < n_it) issue(c_begin + s, s);
    cp_commit();
  }
  for (int64_t it = 0; it < n_it; ++it) {
    cp_wait<GS_STAGES - 2>();
    __syncthreads();
    if (it + GS_STAGES - 1 < n_it) issue(c_begin + it + GS_STAGES - 1, (int)((it + GS_STAGES - 1) % GS_STAGES));
    cp_commit();
    const double* xs = gsm2 + (it % GS_STAGES) * STAGE_D;
#pragma unroll
    for (int k4 = 0; k4 < GS_KC / 4; ++k4) {
      double af[6], bf[3];
#pragma unroll
      for (int i = 0; i < 6; ++i) af[i] = xs[(bi * 48 + i * 8 + (lane >> 2)) * GS_LD + k4 * 4 + (lane & 3)];
#pragma unroll
      for (int j = 0; j < 3; ++j) bf[j] = xs[(bj * 24 + j * 8 + (lane >> 2)) * GS_LD + k4 * 4 + (lane & 3)];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  double* out = partial + ((size_t)split * 2 + mat) * (GS_R * GS_R);   // owned by this CTA: plain read-modify-write
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int r = bi * 48 + i * 8 + (lane >> 2), c = bj * 24 + j * 8 + 2 * (lane & 3);
      double2* p = reinterpret_cast<double2*>(out + r * GS_R + c);
      double2 v = *p;
      v.x += acc[i][j][0];
      v.y += acc[i][j][1];
      *p = v;
    }
}

// G[mat][r][c] (two 144 x 144 matrices, symmetric) = sum over the splits in split order; only c >= r was computed
__global__ void __launch_bounds__(256)
gram_sym_reduce_kernel(const double* __restrict__ partial, int n_split, double* __restrict__ G) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= 2 * GS_R * GS_R) return;
  const int mat = e / (GS_R * GS_R), rc = e % (GS_R * GS_R), r = rc / GS_R, c = rc % GS_R;
  if (c < r) return;
  double s = 0.0;
  for (int k = 0; k < n_split; ++k) s += partial[((size_t)k * 2 + mat) * (GS_R * GS_R) + rc];
  G[(size_t)mat * GS_R * GS_R + r * GS_R + c] = s;
  G[(size_t)mat * GS_R * GS_R + c * GS_R + r] = s;
}

static size_t gram_sym_smem() { return sizeof(double) * GS_STAGES * GS_R * GS_LD; }

int sq_gram_sym_rows() { return GS_R; }

int sq_gram_sym_begin(int n_sm, double** d_partial, size_t* partial_doubles, int* n_split, cudaStream_t st) {
  *n_split = std::max(1, n_sm / 2);                  // two matrices: one CTA per SM
  const size_t need = (size_t)(*n_split) * 2 * GS_R * GS_R;
  if (*partial_doubles < need) {
    if (*d_partial) cudaFree(*d_partial);
    *d_partial = nullptr;
    *partial_doubles = 0;
    SQ_CUDA(cudaMalloc(d_partial, sizeof(double) * need));
    *partial_doubles = need;
  }
  SQ_CUDA(cudaMemsetAsync(*d_partial, 0, sizeof(double) * need, st));
  return SQ_OK;
}

// one panel Z (rows0 S-rows followed by rows1 A-rows, leading dimension ld, K columns)
int sq_gram_sym_panel(const double* Z, int64_t ld, int rows0, int rows1, int64_t K, int n_split, double* d_partial, cudaStream_t st) {
  if (K % GS_KC != 0 || ld % 2 != 0 || rows0 > GS_R || rows1 > GS_R || rows0 < 1) {
    sq_set_error("symmetric Gram panel: K = %lld must be a multiple of %d, at most %d rows per matrix", (long long)K, GS_KC, GS_R);
    return SQ_ERR_INVALID;
  }
  static bool attr = false;
  if (!attr) {
    SQ_CUDA(cudaFuncSetAttribute(gram_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gram_sym_smem()));
    attr = true;
  }
  const int n_mat = rows1 > 0 ? 2 : 1;
  gram_sym_kernel<<<(unsigned)(n_mat * n_split), GS_THREADS, gram_sym_smem(), st>>>(Z, ld, rows0, rows1, K, n_split, d_partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sq_set_error("gram_sym_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}

// d_G: two 144 x 144 row-major matrices (G_SS, G_AA)
int sq_gram_sym_end(int n_split, const double* d_partial, double* d_G, cudaStream_t st) {
  gram_sym_reduce_kernel<<<(2 * GS_R * GS_R + 255) / 256, 256, 0, st>>>(d_partial, n_split, d_G);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sq_set_error("gram_sym_reduce_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// sigma: F[m][t] = sum_k Gm[m][k] D[k][t], m, k < nrow, t in a tile of 64 determinants.
// "Thin" CTAs: 4 warps = 2 (row halves) x 2 (32 determinants each), ~190 registers per thread, one or two CTAs per SM -- the
// tensor pipe is saturated by one warp per SM sub-partition, and most of the SM (registers, warp slots, the LSU pipe) stays
// free for the gather / scatter kernels of the neighbouring panels that run beside it on other streams (sq_sigma's pipeline).
// ------------------------------------------------------------------------------------------------------------------
#define SG_BN 64
#define SG_KC 16
#define SG_LDA 20           // Gm chunk [m][16] row stride: = 4 (mod 16) -> conflict-free A fragments
#define SG_LDB 68           // D chunk [16][64] row stride: = 4 (mod 16) -> conflict-free B fragments
#define SG_STAGES 3
#define SG_THREADS 128

// WM = row parts of a CTA (2 warps each: 32 determinants per warp), MH = row fragments per warp: WM * MH * 8 >= nrow rows per CTA
// (136 rows: 2 x 9 -- four warps, the default; 3 x 6 -- six warps per CTA, three per SM sub-partition with two CTAs per SM: slower)
template <int WM, int MH>
__global__ void __launch_bounds__(WM * 64, 2)
sigma_dmma_kernel(const double* __restrict__ Gm, int ldg, const double* __restrict__ D, double* __restrict__ F, int nrow, int64_t W) {
  extern __shared__ __align__(16) double ssm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // row part and column half of this warp; with two row parts the assignment alternates with the CTA index, so that the two CTAs
  // of an SM put one warp of each row part on every sub-partition (the parts differ in their number of live row fragments)
  const int wm = WM == 2 ? ((warp >> 1) ^ (int)(blockIdx.x & 1u)) : warp >> 1, wn = warp & 1;
  const int64_t t0 = (int64_t)blockIdx.x * SG_BN;
  constexpr int NT = WM * 64;                             // threads
  constexpr int MP = WM * MH * 8;                         // padded rows held in shared memory
  const int m0 = blockIdx.y * MP;                         // first output row of this CTA
  constexpr int stage_d = MP * SG_LDA + SG_KC * SG_LDB;   // doubles per stage
  const int n_chunks = (nrow + SG_KC - 1) / SG_KC;
  const int nfrag = min(MH, (nrow - m0 - wm * MH * 8 + 7) / 8);   // live row fragments of this warp (136 rows: 9 and 8); the others stay zero, unstored
  double acc[MH][4][2];
#pragma unroll
  for (int i = 0; i < MH; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(ssm);
  auto issue = [&](int c, int slot) {
    const int k0 = c * SG_KC;
    const uint32_t sa = sbase + (uint32_t)(slot * stage_d) * 8u, sb = sa + (uint32_t)(MP * SG_LDA) * 8u;
    // Gm chunk: MP rows x 16 doubles = MP * 8 16-byte chunks (rows / columns beyond the matrix are zero-filled)
    for (int id = threadIdx.x; id < MP * 8; id += NT) {
      const int m = id >> 3, ch = id & 7;
      const bool ok = m0 + m < nrow && k0 + ch * 2 < ldg; // ldg is even; the pad column of an odd nrow holds zeros
      cp16(sa + (uint32_t)(m * SG_LDA + ch * 2) * 8u, Gm + (size_t)(ok ? m0 + m : 0) * ldg + (ok ? k0 + ch * 2 : 0), ok ? 16 : 0);
    }
    // D chunk: 16 rows x 64 doubles = 512 16-byte chunks
    for (int id = threadIdx.x; id < SG_KC * (SG_BN / 2); id += NT) {
      const int k = id >> 5, ch = id & 31;
      const bool ok = k0 + k < nrow;
      cp16(sb + (uint32_t)(k * SG_LDB + ch * 2) * 8u, D + (size_t)(ok ? k0 + k : 0) * W + t0 + ch * 2, ok ? 16 : 0);
    }
  };
  for (int s = 0; s < SG_STAGES - 1; ++s) {
    if (s < n_chunks) issue(s, s);
    cp_commit();
  }
  for (int it = 0; it < n_chunks; ++it) {
    cp_wait<SG_STAGES - 2>();
    __syncthreads();
    if (it + SG_STAGES - 1 < n_chunks) issue(it + SG_STAGES - 1, (it + SG_STAGES - 1) % SG_STAGES);
    cp_commit();
    const double* as = ssm + (it % SG_STAGES) * stage_d + (wm * MH * 8) * SG_LDA;
    const double* bs = ssm + (it % SG_STAGES) * stage_d + MP * SG_LDA + wn * 32;
    const int kleft = nrow - it * SG_KC;   // rows of D in this chunk; the zero-filled rest of the last chunk is skipped
#pragma unroll
    for (int k4 = 0; k4 < SG_KC / 4; ++k4) {
      if (k4 * 4 >= kleft) break;
      double bf[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = bs[(k4 * 4 + (lane & 3)) * SG_LDB + j * 8 + (lane >> 2)];
#pragma unroll
      for (int i = 0; i < MH; ++i) {
        if (i < nfrag) {
          const double a = as[(i * 8 + (lane >> 2)) * SG_LDA + k4 * 4 + (lane & 3)];
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a, bf[j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MH; ++i) {
    const int m = m0 + (wm * MH + i) * 8 + (lane >> 2);
    if (m < nrow) {
      double* p = F + (size_t)m * W + t0 + wn * 32 + 2 * (lane & 3);
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<double2*>(p + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

static int g_sigma_row_parts = 2;    // sq_set_option("sgemm_wm", "2" | "3"): row parts (pairs of warps) of a sigma GEMM CTA.  Measured at CAS(16,16): sigma 174 ms
                                     // with 2 x 9 fragments, 187 ms with 3 x 6 (more A-fragment loads per DMMA: the kernel is bound by the shared-memory
                                     // data pipe, not by warps per sub-partition; profiles/r2_visit38_ab_sgemm_wm.txt)
void sq_sigma_gemm_set_row_parts(int n) { g_sigma_row_parts = n == 3 ? 3 : 2; }
static int g_sigma_cta_per_sm = 2;   // measured at CAS(16,16): sigma 445 ms with two CTAs per SM, 524 ms with one (profiles/r2_visit5_ab_sgemm_cta.txt); sq_set_option("sgemm_cta", "1" | "2"): GEMM CTAs per SM (the rest of the SM runs gathers / scatters)
void sq_sigma_gemm_set_residency(int n) { g_sigma_cta_per_sm = n == 1 ? 1 : 2; }

template <int WM, int MH>
static int launch_sigma_mh(const double* Gm, int ldg, const double* D, double* F, int nrow, int64_t W, int m_tiles, cudaStream_t st) {
  size_t smem = sizeof(double) * SG_STAGES * ((size_t)WM * MH * 8 * SG_LDA + SG_KC * SG_LDB);
  if (g_sigma_cta_per_sm == 1 && smem < 116 * 1024) smem = 116 * 1024;   // more than half of the SM's shared memory: one CTA per SM
  static size_t attr = 0;
  if (smem > attr) {
    SQ_CUDA(cudaFuncSetAttribute(sigma_dmma_kernel<WM, MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  sigma_dmma_kernel<WM, MH><<<dim3((unsigned)(W / SG_BN), (unsigned)m_tiles), WM * 64, smem, st>>>(Gm, ldg, D, F, nrow, W);
  return SQ_OK;
}

// F (nrow x W, row-major, ld W) = Gm (nrow x nrow, row-major with leading dimension ldg: even, pad column zero) * D (nrow x W,
// row-major)
int sq_sigma_gemm(const double* Gm, int ldg, const double* D, double* F, int nrow, int64_t W, cudaStream_t st) {
  if (W % SG_BN != 0 || nrow < 1 || ldg % 2 != 0 || ldg < nrow) {
    sq_set_error("sigma GEMM: panel width %lld must be a multiple of %d and the leading dimension %d even", (long long)W, SG_BN, ldg);
    return SQ_ERR_INVALID;
  }
  // accumulators: 8 * MH doubles per thread, MH <= 9: up to 144 rows per CTA; more rows (n = 16 with unsymmetric integrals: 256,
  // n = 20 symmetrised: 210) are split into equal row tiles along gridDim.y
  const int m_tiles = (nrow + 143) / 144;
  const int rows_cta = (nrow + m_tiles - 1) / m_tiles;
  int rc = SQ_ERR_UNSUPPORTED;
  if (g_sigma_row_parts == 3 && rows_cta > 48) {   // six warps per CTA: 24 * MH rows
    switch ((rows_cta + 23) / 24) {
#define SG_CASE(M) case M: rc = launch_sigma_mh<3, M>(Gm, ldg, D, F, nrow, W, m_tiles, st); break;
      SG_CASE(3) SG_CASE(4) SG_CASE(5) SG_CASE(6)
#undef SG_CASE
      default: break;
    }
  } else {
    switch ((rows_cta + 15) / 16) {
#define SG_CASE(M) case M: rc = launch_sigma_mh<2, M>(Gm, ldg, D, F, nrow, W, m_tiles, st); break;
      SG_CASE(1) SG_CASE(2) SG_CASE(3) SG_CASE(4) SG_CASE(5) SG_CASE(6) SG_CASE(7) SG_CASE(8) SG_CASE(9)
#undef SG_CASE
      default: break;
    }
  }
  if (rc == SQ_ERR_UNSUPPORTED) {
    sq_set_error("sigma GEMM: no kernel instantiation for %d generator rows", nrow);
    return rc;
  }
  SQ_CHECK(rc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sq_set_error("sigma_dmma_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// rdm1 without the 2-RDM: g1[row] += sum_t D[row][t] * x[t]   (one CTA per row: fixed summation order, no atomics)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
panel_gemv_kernel(const double* __restrict__ D, int64_t W, const double* __restrict__ x, int64_t K, double* __restrict__ g1) {
  __shared__ double red[32];
  const double* row = D + (size_t)blockIdx.x * W;
  double s = 0.0;
  for (int64_t t = threadIdx.x; t < K; t += 1024) s += row[t] * x[t];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = red[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) g1[blockIdx.x] += v;
  }
}

int sq_panel_gemv(const double* D, int64_t W, int nrows, const double* x, int64_t K, double* g1, cudaStream_t st) {
  panel_gemv_kernel<<<(unsigned)nrows, 1024, 0, st>>>(D, W, x, K, g1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sq_set_error("panel_gemv_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}
