// sm_100a kernels of the state-vector engine.  All of them are fp64 streaming kernels bounded by HBM
// bandwidth: amplitudes are read and written exactly once per launch with fully coalesced 8-byte lanes
// along the beta-string (column) axis; sign / partner tables are 4-byte-per-string codes that stay in
// L1/L2.  No tensor-core shapes exist on this path (it is integer address work + 2x2 rotations).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sqsv_internal.h"

#define TILE_THREADS 256
#define TILE_ROWS 8

struct TileProgram {
  int n;
  int kind[SQ_MAX_PROGRAM];   // 0 alpha single, 1 beta single, 2 pair double
  int slot[SQ_MAX_PROGRAM];   // gradient slot (grad kernel only)
  double c[SQ_MAX_PROGRAM];
  double s[SQ_MAX_PROGRAM];
};

__device__ __forceinline__ void rot(double& xs, double& xt, double c, double s) {
  // exp(theta T) on a (src,tgt) pair with T[tgt,src] = +1 (sign folded into s):
  //   src' = c src - s tgt ; tgt' = c tgt + s src          (SURVEY 8a "tUPS primitive spec")
  double a = xs, b = xt;
  xs = c * a - s * b;
  xt = c * b + s * a;
}

__device__ __forceinline__ double sgnbit(uint32_t code, int bit) { return (code >> bit) & 1u ? -1.0 : 1.0; }

// ---------------------------------------------------------------------------------------------
// Tile kernel: applies a fused program of alpha-single / beta-single / pair-double rotations that all
// act on the same spatial orbital pair (i,a) -- e.g. one whole tUPS brick [sa_single, double, sa_single]
// (reference util.py:694-745, operator_state_algebra.py:1002-1085) -- in ONE sweep.  A thread owns the
// tile {Ia, Ia'} x {Ib, Ib'} (or the degenerate 2x1 / 1x2 tile when one of the strings is inert) in
// registers; tiles partition the vector, so the update is in place and race free.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_THREADS)
tile_kernel(double* __restrict__ C, const uint32_t* __restrict__ codeA, const uint32_t* __restrict__ codeB,
            const int32_t* __restrict__ rows, int64_t n_rows, int64_t NB, int64_t row_begin,
            const TileProgram prog) {
  const int64_t ib = (int64_t)blockIdx.x * TILE_THREADS + threadIdx.x;
  if (ib >= NB) return;
  const uint32_t cb = __ldg(codeB + ib);
  const uint32_t clsb = cb & 3u;
  if (clsb == SQ_CLS_TGT) return;
  const int64_t ibp = (int64_t)(cb >> 5);
  const uint32_t cbp = (clsb == SQ_CLS_SRC) ? __ldg(codeB + ibp) : 0u;
  const double sSb = sgnbit(cb, 2), crb = sgnbit(cb, 3), crbp = sgnbit(cbp, 3), db = sgnbit(cb, 4);

  const int64_t r0 = (int64_t)blockIdx.y * TILE_ROWS;
  const int64_t r1 = (r0 + TILE_ROWS < n_rows) ? r0 + TILE_ROWS : n_rows;
  for (int64_t r = r0; r < r1; ++r) {
    const int64_t ia = rows[r];
    const uint32_t ca = __ldg(codeA + ia);
    const uint32_t clsa = ca & 3u;
    double* row0 = C + (ia - row_begin) * NB;
    if (clsa == SQ_CLS_INERT) {
      if (clsb != SQ_CLS_SRC) continue;
      double x0 = row0[ib], x1 = row0[ibp];
      const double sg = sSb * sgnbit(ca, 3);
#pragma unroll
      for (int k = 0; k < SQ_MAX_PROGRAM; ++k) {
        if (k >= prog.n) break;
        if (prog.kind[k] == 1) rot(x0, x1, prog.c[k], sg * prog.s[k]);
      }
      row0[ib] = x0;
      row0[ibp] = x1;
    } else {  // src row; its partner row rides along
      const int64_t iap = (int64_t)(ca >> 5);
      const uint32_t cap = __ldg(codeA + iap);
      double* row1 = C + (iap - row_begin) * NB;
      const double sSa = sgnbit(ca, 2), cra = sgnbit(ca, 3), crap = sgnbit(cap, 3), da = sgnbit(ca, 4);
      if (clsb == SQ_CLS_INERT) {
        double x0 = row0[ib], x1 = row1[ib];
        const double sg = sSa * crb;
#pragma unroll
        for (int k = 0; k < SQ_MAX_PROGRAM; ++k) {
          if (k >= prog.n) break;
          if (prog.kind[k] == 0) rot(x0, x1, prog.c[k], sg * prog.s[k]);
        }
        row0[ib] = x0;
        row1[ib] = x1;
      } else {
        double x00 = row0[ib], x01 = row0[ibp], x10 = row1[ib], x11 = row1[ibp];
        const double sgA0 = sSa * crb, sgA1 = sSa * crbp;   // alpha rotation in column Ib / Ib'
        const double sgB0 = sSb * cra, sgB1 = sSb * crap;   // beta rotation in row Ia / Ia'
        const double sgD = da * db;
#pragma unroll
        for (int k = 0; k < SQ_MAX_PROGRAM; ++k) {
          if (k >= prog.n) break;
          const double c = prog.c[k], s = prog.s[k];
          if (prog.kind[k] == 0) {
            rot(x00, x10, c, sgA0 * s);
            rot(x01, x11, c, sgA1 * s);
          } else if (prog.kind[k] == 1) {
            rot(x00, x01, c, sgB0 * s);
            rot(x10, x11, c, sgB1 * s);
          } else {
            rot(x00, x11, c, sgD * s);
          }
        }
        row0[ib] = x00;
        row0[ibp] = x01;
        row1[ib] = x10;
        row1[ibp] = x11;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// tile_kernel_v2: same tiles, restructured for the memory roofline.
//   * the whole fused program is pre-multiplied on the host into ONE 4x4 matrix (src x src tiles) and two
//     2x2 rotations (src x inert, inert x src).  Per-determinant signs are removed by a diagonal +-1 gauge
//     (x10 -> sgA0 x10, x01 -> sgB0 x01, x11 -> sgA0 sgB1 x11), which turns every alpha/beta rotation sign
//     into +1 and leaves one gauge-invariant sign sigma on the pair double; sigma is uniform over the space
//     (checked on the host), so the matrix is a kernel constant.  16 DFMA per 4 amplitudes instead of a
//     branchy step loop.
//   * columns and rows come as compacted, class-homogeneous work lists: every CTA is src-or-inert uniform,
//     no lane idles on a "tgt" string and inert x inert CTAs exit at once.
//   * all loads of a CTA's row batch are issued before the first store (ROWS_PER_ITER x 4 independent
//     8-byte loads per thread in flight).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double flip(double x, int neg) {
  // multiply by +-1 through the sign bit
  return __hiloint2double(__double2hiint(x) ^ (neg << 31), __double2loint(x));
}

#define V2_ROWS_PER_ITER 4

// Row pointers: a row item names the owner rank of each of its two rows; `bases` holds the peer-mapped
// base pointer of every rank's shard (NVLink P2P), so a tile whose rows live on two GPUs is rotated in
// place by plain loads/stores on local + remote memory.  On one GPU both owners are rank 0.
// PEER = false (one GPU, or a sharded operator whose pairs are all local): rows are offsets from the local
// shard.  PEER = true: the owner's base pointer comes from a 16-entry device table (an indexed kernel-parameter
// array would be spilled to local memory).
template <bool PEER>
struct Bases {
  double* C;
  const unsigned long long* tab;
};
template <bool PEER>
__device__ __forceinline__ double* rowp0(const Bases<PEER>& b, const int4& ri, int64_t NB) {
  if (PEER) return reinterpret_cast<double*>(__ldg(b.tab + ((ri.z >> 8) & 0xff))) + (int64_t)ri.x * NB;
  return b.C + (int64_t)ri.x * NB;
}
template <bool PEER>
__device__ __forceinline__ double* rowp1(const Bases<PEER>& b, const int4& ri, int64_t NB) {
  if (PEER) return reinterpret_cast<double*>(__ldg(b.tab + ((ri.z >> 16) & 0xff))) + (int64_t)ri.y * NB;
  return b.C + (int64_t)ri.y * NB;
}
// pad entries have x < 0; cross-device pairs are split between the two owners by column-CTA parity (w = 1 / 2)
template <bool PEER>
__device__ __forceinline__ bool item_on(const int4& ri) {
  if (PEER) return ri.x >= 0 && (ri.w == 0 || (int)(blockIdx.x & 1u) == ri.w - 1);
  return ri.x >= 0;
}

template <bool PEER>
__global__ void __launch_bounds__(TILE_THREADS)
tile_kernel_v2(const Bases<PEER> bases, const int2* __restrict__ colItems, int n_colblk_src,
               const int4* __restrict__ rowItems, int n_rowchunk_src, int64_t NB, const TileMatrices tm) {
  const bool col_src = (int)blockIdx.x < n_colblk_src;
  const bool row_src = (int)blockIdx.y < n_rowchunk_src;
  if (!col_src && !row_src) return;
  const int2 ci = __ldg(colItems + (int64_t)blockIdx.x * TILE_THREADS + threadIdx.x);
  if (ci.x < 0) return;
  const int64_t ib = ci.x;
  const int64_t ibp = ci.y & 0x07ffffff;
  const int cf = (int)((uint32_t)ci.y >> 27);
  const int sSb = cf & 1, crb = (cf >> 1) & 1;
  const int4* rit = rowItems + (int64_t)blockIdx.y * TILE_ROWS;

  if (row_src && col_src) {
#pragma unroll 1
    for (int it = 0; it < TILE_ROWS / V2_ROWS_PER_ITER; ++it) {
      int4 ri[V2_ROWS_PER_ITER];
      double x00[V2_ROWS_PER_ITER], x01[V2_ROWS_PER_ITER], x10[V2_ROWS_PER_ITER], x11[V2_ROWS_PER_ITER];
#pragma unroll
      for (int j = 0; j < V2_ROWS_PER_ITER; ++j) ri[j] = __ldg(rit + it * V2_ROWS_PER_ITER + j);
#pragma unroll
      for (int j = 0; j < V2_ROWS_PER_ITER; ++j) {
        if (item_on<PEER>(ri[j])) {
          const double* r0 = rowp0(bases, ri[j], NB);
          const double* r1 = rowp1(bases, ri[j], NB);
          x00[j] = r0[ib];
          x01[j] = r0[ibp];
          x10[j] = r1[ib];
          x11[j] = r1[ibp];
        }
      }
#pragma unroll
      for (int j = 0; j < V2_ROWS_PER_ITER; ++j) {
        if (item_on<PEER>(ri[j])) {
          const int rf = ri[j].z;
          const int sSa = rf & 1, cra = (rf >> 1) & 1, crap = (rf >> 2) & 1;
          const int g10 = sSa ^ crb;            // sgA0
          const int g01 = sSb ^ cra;            // sgB0
          const int g11 = g10 ^ sSb ^ crap;     // sgA0 * sgB1
          const double y0 = x00[j], y1 = flip(x01[j], g01), y2 = flip(x10[j], g10), y3 = flip(x11[j], g11);
          const double z0 = tm.m[0] * y0 + tm.m[1] * y1 + tm.m[2] * y2 + tm.m[3] * y3;
          const double z1 = tm.m[4] * y0 + tm.m[5] * y1 + tm.m[6] * y2 + tm.m[7] * y3;
          const double z2 = tm.m[8] * y0 + tm.m[9] * y1 + tm.m[10] * y2 + tm.m[11] * y3;
          const double z3 = tm.m[12] * y0 + tm.m[13] * y1 + tm.m[14] * y2 + tm.m[15] * y3;
          double* r0 = rowp0(bases, ri[j], NB);
          double* r1 = rowp1(bases, ri[j], NB);
          r0[ib] = z0;
          r0[ibp] = flip(z1, g01);
          r1[ib] = flip(z2, g10);
          r1[ibp] = flip(z3, g11);
        }
      }
    }
  } else if (row_src) {   // src row pair x inert column: alpha rotation only
#pragma unroll 1
    for (int it = 0; it < TILE_ROWS / V2_ROWS_PER_ITER; ++it) {
      int4 ri[V2_ROWS_PER_ITER];
      double x0[V2_ROWS_PER_ITER], x1[V2_ROWS_PER_ITER];
#pragma unroll
      for (int j = 0; j < V2_ROWS_PER_ITER; ++j) ri[j] = __ldg(rit + it * V2_ROWS_PER_ITER + j);
#pragma unroll
      for (int j = 0; j < V2_ROWS_PER_ITER; ++j) {
        if (item_on<PEER>(ri[j])) {
          x0[j] = rowp0(bases, ri[j], NB)[ib];
          x1[j] = rowp1(bases, ri[j], NB)[ib];
        }
      }
#pragma unroll
      for (int j = 0; j < V2_ROWS_PER_ITER; ++j) {
        if (item_on<PEER>(ri[j])) {
          const int g = (ri[j].z & 1) ^ crb;   // sSa * crossB(column)
          const double a = x0[j], b = flip(x1[j], g);
          rowp0(bases, ri[j], NB)[ib] = tm.ca * a - tm.sa * b;
          rowp1(bases, ri[j], NB)[ib] = flip(tm.ca * b + tm.sa * a, g);
        }
      }
    }
  } else {                // inert row x src column pair: beta rotation only
#pragma unroll 1
    for (int it = 0; it < TILE_ROWS / V2_ROWS_PER_ITER; ++it) {
      int4 ri[V2_ROWS_PER_ITER];
      double x0[V2_ROWS_PER_ITER], x1[V2_ROWS_PER_ITER];
#pragma unroll
      for (int j = 0; j < V2_ROWS_PER_ITER; ++j) ri[j] = __ldg(rit + it * V2_ROWS_PER_ITER + j);
#pragma unroll
      for (int j = 0; j < V2_ROWS_PER_ITER; ++j) {
        if (item_on<PEER>(ri[j])) {
          const double* r0 = rowp0(bases, ri[j], NB);
          x0[j] = r0[ib];
          x1[j] = r0[ibp];
        }
      }
#pragma unroll
      for (int j = 0; j < V2_ROWS_PER_ITER; ++j) {
        if (item_on<PEER>(ri[j])) {
          const int g = sSb ^ ((ri[j].z >> 1) & 1);   // sSb * crossA(row)
          const double a = x0[j], b = flip(x1[j], g);
          double* r0 = rowp0(bases, ri[j], NB);
          r0[ib] = tm.cb * a - tm.sb * b;
          r0[ibp] = flip(tm.cb * b + tm.sb * a, g);
        }
      }
    }
  }
}

// block-wide sum of NS per-thread values -> partial[blockLinear*NS + k]; deterministic order.
template <int NS>
__device__ __forceinline__ void block_reduce_store(double* vals, double* __restrict__ partial, int64_t blockLinear) {
  __shared__ double sm[NS][TILE_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    double v = vals[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) sm[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < NS) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < TILE_THREADS / 32; ++w) v += sm[threadIdx.x][w];
    partial[blockLinear * NS + threadIdx.x] = v;
  }
}

// Same tiles on two vectors (bra, ket): before each rotation step accumulate <bra|T_step|ket>
// (reference ups_wavefunction.py:1114-1138: g_i = 2 <bra|T_i|ket>, then bra <- U_i bra, ket <- U_i ket).
__global__ void __launch_bounds__(TILE_THREADS)
tile_grad_kernel(double* __restrict__ BRA, double* __restrict__ KET, const uint32_t* __restrict__ codeA,
                 const uint32_t* __restrict__ codeB, const int32_t* __restrict__ rows, int64_t n_rows,
                 int64_t NB, int64_t row_begin, const TileProgram prog, double* __restrict__ partial) {
  double acc[SQ_MAX_PROGRAM];
#pragma unroll
  for (int k = 0; k < SQ_MAX_PROGRAM; ++k) acc[k] = 0.0;
  const int64_t ib = (int64_t)blockIdx.x * TILE_THREADS + threadIdx.x;
  uint32_t cb = 0, clsb = SQ_CLS_TGT;
  if (ib < NB) {
    cb = __ldg(codeB + ib);
    clsb = cb & 3u;
  }
  if (clsb != SQ_CLS_TGT) {
    const int64_t ibp = (int64_t)(cb >> 5);
    const uint32_t cbp = (clsb == SQ_CLS_SRC) ? __ldg(codeB + ibp) : 0u;
    const double sSb = sgnbit(cb, 2), crb = sgnbit(cb, 3), crbp = sgnbit(cbp, 3), db = sgnbit(cb, 4);
    const int64_t r0 = (int64_t)blockIdx.y * TILE_ROWS;
    const int64_t r1 = (r0 + TILE_ROWS < n_rows) ? r0 + TILE_ROWS : n_rows;
    for (int64_t r = r0; r < r1; ++r) {
      const int64_t ia = rows[r];
      const uint32_t ca = __ldg(codeA + ia);
      const uint32_t clsa = ca & 3u;
      const int64_t o0 = (ia - row_begin) * NB;
      if (clsa == SQ_CLS_INERT) {
        if (clsb != SQ_CLS_SRC) continue;
        double b0 = BRA[o0 + ib], b1 = BRA[o0 + ibp], k0 = KET[o0 + ib], k1 = KET[o0 + ibp];
        const double sg = sSb * sgnbit(ca, 3);
#pragma unroll
        for (int k = 0; k < SQ_MAX_PROGRAM; ++k) {
          if (k >= prog.n) break;
          if (prog.kind[k] == 1) {
            acc[k] += sg * (b1 * k0 - b0 * k1);
            rot(b0, b1, prog.c[k], sg * prog.s[k]);
            rot(k0, k1, prog.c[k], sg * prog.s[k]);
          }
        }
        BRA[o0 + ib] = b0; BRA[o0 + ibp] = b1; KET[o0 + ib] = k0; KET[o0 + ibp] = k1;
      } else {
        const int64_t iap = (int64_t)(ca >> 5);
        const uint32_t cap = __ldg(codeA + iap);
        const int64_t o1 = (iap - row_begin) * NB;
        const double sSa = sgnbit(ca, 2), cra = sgnbit(ca, 3), crap = sgnbit(cap, 3), da = sgnbit(ca, 4);
        if (clsb == SQ_CLS_INERT) {
          double b0 = BRA[o0 + ib], b1 = BRA[o1 + ib], k0 = KET[o0 + ib], k1 = KET[o1 + ib];
          const double sg = sSa * crb;
#pragma unroll
          for (int k = 0; k < SQ_MAX_PROGRAM; ++k) {
            if (k >= prog.n) break;
            if (prog.kind[k] == 0) {
              acc[k] += sg * (b1 * k0 - b0 * k1);
              rot(b0, b1, prog.c[k], sg * prog.s[k]);
              rot(k0, k1, prog.c[k], sg * prog.s[k]);
            }
          }
          BRA[o0 + ib] = b0; BRA[o1 + ib] = b1; KET[o0 + ib] = k0; KET[o1 + ib] = k1;
        } else {
          double b00 = BRA[o0 + ib], b01 = BRA[o0 + ibp], b10 = BRA[o1 + ib], b11 = BRA[o1 + ibp];
          double k00 = KET[o0 + ib], k01 = KET[o0 + ibp], k10 = KET[o1 + ib], k11 = KET[o1 + ibp];
          const double sgA0 = sSa * crb, sgA1 = sSa * crbp;
          const double sgB0 = sSb * cra, sgB1 = sSb * crap;
          const double sgD = da * db;
#pragma unroll
          for (int k = 0; k < SQ_MAX_PROGRAM; ++k) {
            if (k >= prog.n) break;
            const double c = prog.c[k], s = prog.s[k];
            if (prog.kind[k] == 0) {
              acc[k] += sgA0 * (b10 * k00 - b00 * k10) + sgA1 * (b11 * k01 - b01 * k11);
              rot(b00, b10, c, sgA0 * s); rot(b01, b11, c, sgA1 * s);
              rot(k00, k10, c, sgA0 * s); rot(k01, k11, c, sgA1 * s);
            } else if (prog.kind[k] == 1) {
              acc[k] += sgB0 * (b01 * k00 - b00 * k01) + sgB1 * (b11 * k10 - b10 * k11);
              rot(b00, b01, c, sgB0 * s); rot(b10, b11, c, sgB1 * s);
              rot(k00, k01, c, sgB0 * s); rot(k10, k11, c, sgB1 * s);
            } else {
              acc[k] += sgD * (b11 * k00 - b00 * k11);
              rot(b00, b11, c, sgD * s);
              rot(k00, k11, c, sgD * s);
            }
          }
          BRA[o0 + ib] = b00; BRA[o0 + ibp] = b01; BRA[o1 + ib] = b10; BRA[o1 + ibp] = b11;
          KET[o0 + ib] = k00; KET[o0 + ibp] = k01; KET[o1 + ib] = k10; KET[o1 + ibp] = k11;
        }
      }
    }
  }
  block_reduce_store<SQ_MAX_PROGRAM>(acc, partial, (int64_t)blockIdx.y * gridDim.x + blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// tile_grad_kernel_v2: the fused gradient step on the compacted work lists of tile_kernel_v2.
// In the gauge-fixed basis every alpha/beta generator element is +1 (sigma on the pair double), so
//   <bra|T_step|ket> = sum_tiles sum_pairs (b_tgt k_src - b_src k_tgt)
// without per-element sign arithmetic; bra and ket tiles are rotated in registers right after.
// ---------------------------------------------------------------------------------------------
struct GradProgram {
  int n;
  int kind[SQ_MAX_PROGRAM];
  double c[SQ_MAX_PROGRAM];
  double s[SQ_MAX_PROGRAM];   // pair-double entries already carry sigma
  double sig[SQ_MAX_PROGRAM]; // generator element of the step in the gauge-fixed basis (+1, or sigma)
};

#define G2_ROWS_PER_ITER 2

__device__ __forceinline__ void grad_pair(double& bs, double& bt, double& ks, double& kt, double c, double s, double sig,
                                          double& acc) {
  acc += sig * (bt * ks - bs * kt);
  rot(bs, bt, c, s);
  rot(ks, kt, c, s);
}

__global__ void __launch_bounds__(TILE_THREADS)
tile_grad_kernel_v2(double* __restrict__ BRA, double* __restrict__ KET, const int2* __restrict__ colItems,
                    int n_colblk_src, const int4* __restrict__ rowItems, int n_rowchunk_src, int64_t NB,
                    const GradProgram gp, double* __restrict__ partial) {
  double acc[SQ_MAX_PROGRAM];
#pragma unroll
  for (int k = 0; k < SQ_MAX_PROGRAM; ++k) acc[k] = 0.0;
  const bool col_src = (int)blockIdx.x < n_colblk_src;
  const bool row_src = (int)blockIdx.y < n_rowchunk_src;
  const int2 ci = __ldg(colItems + (int64_t)blockIdx.x * TILE_THREADS + threadIdx.x);
  if ((col_src || row_src) && ci.x >= 0) {
    const int64_t ib = ci.x;
    const int64_t ibp = ci.y & 0x07ffffff;
    const int cf = (int)((uint32_t)ci.y >> 27);
    const int sSb = cf & 1, crb = (cf >> 1) & 1;
    const int4* rit = rowItems + (int64_t)blockIdx.y * TILE_ROWS;
    if (row_src && col_src) {
#pragma unroll 1
      for (int it = 0; it < TILE_ROWS / G2_ROWS_PER_ITER; ++it) {
        int4 ri[G2_ROWS_PER_ITER];
        double b[G2_ROWS_PER_ITER][4], k[G2_ROWS_PER_ITER][4];
#pragma unroll
        for (int j = 0; j < G2_ROWS_PER_ITER; ++j) ri[j] = __ldg(rit + it * G2_ROWS_PER_ITER + j);
#pragma unroll
        for (int j = 0; j < G2_ROWS_PER_ITER; ++j) {
          if (ri[j].x >= 0) {
            const int64_t o0 = (int64_t)ri[j].x * NB, o1 = (int64_t)ri[j].y * NB;
            b[j][0] = BRA[o0 + ib]; b[j][1] = BRA[o0 + ibp]; b[j][2] = BRA[o1 + ib]; b[j][3] = BRA[o1 + ibp];
            k[j][0] = KET[o0 + ib]; k[j][1] = KET[o0 + ibp]; k[j][2] = KET[o1 + ib]; k[j][3] = KET[o1 + ibp];
          }
        }
#pragma unroll
        for (int j = 0; j < G2_ROWS_PER_ITER; ++j) {
          if (ri[j].x >= 0) {
            const int rf = ri[j].z;
            const int sSa = rf & 1, cra = (rf >> 1) & 1, crap = (rf >> 2) & 1;
            const int g10 = sSa ^ crb, g01 = sSb ^ cra, g11 = g10 ^ sSb ^ crap;
            b[j][1] = flip(b[j][1], g01); b[j][2] = flip(b[j][2], g10); b[j][3] = flip(b[j][3], g11);
            k[j][1] = flip(k[j][1], g01); k[j][2] = flip(k[j][2], g10); k[j][3] = flip(k[j][3], g11);
#pragma unroll
            for (int s = 0; s < SQ_MAX_PROGRAM; ++s) {
              if (s >= gp.n) break;
              const double c = gp.c[s], sn = gp.s[s], sg = gp.sig[s];
              if (gp.kind[s] == 0) {
                grad_pair(b[j][0], b[j][2], k[j][0], k[j][2], c, sn, sg, acc[s]);
                grad_pair(b[j][1], b[j][3], k[j][1], k[j][3], c, sn, sg, acc[s]);
              } else if (gp.kind[s] == 1) {
                grad_pair(b[j][0], b[j][1], k[j][0], k[j][1], c, sn, sg, acc[s]);
                grad_pair(b[j][2], b[j][3], k[j][2], k[j][3], c, sn, sg, acc[s]);
              } else {
                grad_pair(b[j][0], b[j][3], k[j][0], k[j][3], c, sn, sg, acc[s]);
              }
            }
            const int64_t o0 = (int64_t)ri[j].x * NB, o1 = (int64_t)ri[j].y * NB;
            BRA[o0 + ib] = b[j][0]; BRA[o0 + ibp] = flip(b[j][1], g01);
            BRA[o1 + ib] = flip(b[j][2], g10); BRA[o1 + ibp] = flip(b[j][3], g11);
            KET[o0 + ib] = k[j][0]; KET[o0 + ibp] = flip(k[j][1], g01);
            KET[o1 + ib] = flip(k[j][2], g10); KET[o1 + ibp] = flip(k[j][3], g11);
          }
        }
      }
    } else {
      // 2-amplitude tiles: (src row pair x inert column) feels only alpha steps, (inert row x src column pair) only beta
      const int want = row_src ? 0 : 1;
#pragma unroll 1
      for (int it = 0; it < TILE_ROWS / G2_ROWS_PER_ITER; ++it) {
        int4 ri[G2_ROWS_PER_ITER];
        double b0[G2_ROWS_PER_ITER], b1[G2_ROWS_PER_ITER], k0[G2_ROWS_PER_ITER], k1[G2_ROWS_PER_ITER];
        int64_t i0[G2_ROWS_PER_ITER], i1[G2_ROWS_PER_ITER];
#pragma unroll
        for (int j = 0; j < G2_ROWS_PER_ITER; ++j) ri[j] = __ldg(rit + it * G2_ROWS_PER_ITER + j);
#pragma unroll
        for (int j = 0; j < G2_ROWS_PER_ITER; ++j) {
          if (ri[j].x >= 0) {
            i0[j] = (int64_t)ri[j].x * NB + ib;
            i1[j] = row_src ? (int64_t)ri[j].y * NB + ib : (int64_t)ri[j].x * NB + ibp;
            b0[j] = BRA[i0[j]]; b1[j] = BRA[i1[j]]; k0[j] = KET[i0[j]]; k1[j] = KET[i1[j]];
          }
        }
#pragma unroll
        for (int j = 0; j < G2_ROWS_PER_ITER; ++j) {
          if (ri[j].x >= 0) {
            const int g = row_src ? ((ri[j].z & 1) ^ crb) : (sSb ^ ((ri[j].z >> 1) & 1));
            b1[j] = flip(b1[j], g);
            k1[j] = flip(k1[j], g);
#pragma unroll
            for (int s = 0; s < SQ_MAX_PROGRAM; ++s) {
              if (s >= gp.n) break;
              if (gp.kind[s] == want) grad_pair(b0[j], b1[j], k0[j], k1[j], gp.c[s], gp.s[s], 1.0, acc[s]);
            }
            BRA[i0[j]] = b0[j]; BRA[i1[j]] = flip(b1[j], g);
            KET[i0[j]] = k0[j]; KET[i1[j]] = flip(k1[j], g);
          }
        }
      }
    }
  }
  block_reduce_store<SQ_MAX_PROGRAM>(acc, partial, (int64_t)blockIdx.y * gridDim.x + blockIdx.x);
}

// tile_grad_peer_kernel: tile_grad_kernel_v2 for an exchange operator of an alpha-sharded (bra, ket) pair -- the two rows of a
// tile may live on two GPUs; both are read and written in place through the peer mappings (Bases<true>, one table per vector), and a
// cross-device row pair is split between its two owners by column-CTA parity exactly as in tile_kernel_v2<true> (item_on).  Every
// rank accumulates <bra|T_step|ket> over the tiles IT processes; the caller adds the ranks' partial sums (all-reduce).
// Written without GPU time (compiled, not yet run): reached only through sq_ups_grad_sweep_dist.
__global__ void __launch_bounds__(TILE_THREADS)
tile_grad_peer_kernel(const Bases<true> BB, const Bases<true> KB, const int2* __restrict__ colItems, int n_colblk_src,
                      const int4* __restrict__ rowItems, int n_rowchunk_src, int64_t NB, const GradProgram gp,
                      double* __restrict__ partial) {
  double acc[SQ_MAX_PROGRAM];
#pragma unroll
  for (int k = 0; k < SQ_MAX_PROGRAM; ++k) acc[k] = 0.0;
  const bool col_src = (int)blockIdx.x < n_colblk_src;
  const bool row_src = (int)blockIdx.y < n_rowchunk_src;
  const int2 ci = __ldg(colItems + (int64_t)blockIdx.x * TILE_THREADS + threadIdx.x);
  if ((col_src || row_src) && ci.x >= 0) {
    const int64_t ib = ci.x;
    const int64_t ibp = ci.y & 0x07ffffff;
    const int cf = (int)((uint32_t)ci.y >> 27);
    const int sSb = cf & 1, crb = (cf >> 1) & 1;
    const int4* rit = rowItems + (int64_t)blockIdx.y * TILE_ROWS;
    if (row_src && col_src) {
#pragma unroll 1
      for (int it = 0; it < TILE_ROWS / G2_ROWS_PER_ITER; ++it) {
        int4 ri[G2_ROWS_PER_ITER];
        double b[G2_ROWS_PER_ITER][4], k[G2_ROWS_PER_ITER][4];
#pragma unroll
        for (int j = 0; j < G2_ROWS_PER_ITER; ++j) ri[j] = __ldg(rit + it * G2_ROWS_PER_ITER + j);
#pragma unroll
        for (int j = 0; j < G2_ROWS_PER_ITER; ++j) {
          if (item_on<true>(ri[j])) {
            const double* b0 = rowp0(BB, ri[j], NB);
            const double* b1 = rowp1(BB, ri[j], NB);
            const double* k0 = rowp0(KB, ri[j], NB);
            const double* k1 = rowp1(KB, ri[j], NB);
            b[j][0] = b0[ib]; b[j][1] = b0[ibp]; b[j][2] = b1[ib]; b[j][3] = b1[ibp];
            k[j][0] = k0[ib]; k[j][1] = k0[ibp]; k[j][2] = k1[ib]; k[j][3] = k1[ibp];
          }
        }
#pragma unroll
        for (int j = 0; j < G2_ROWS_PER_ITER; ++j) {
          if (item_on<true>(ri[j])) {
            const int rf = ri[j].z;
            const int sSa = rf & 1, cra = (rf >> 1) & 1, crap = (rf >> 2) & 1;
            const int g10 = sSa ^ crb, g01 = sSb ^ cra, g11 = g10 ^ sSb ^ crap;
            b[j][1] = flip(b[j][1], g01); b[j][2] = flip(b[j][2], g10); b[j][3] = flip(b[j][3], g11);
            k[j][1] = flip(k[j][1], g01); k[j][2] = flip(k[j][2], g10); k[j][3] = flip(k[j][3], g11);
#pragma unroll
            for (int s = 0; s < SQ_MAX_PROGRAM; ++s) {
              if (s >= gp.n) break;
              const double c = gp.c[s], sn = gp.s[s], sg = gp.sig[s];
              if (gp.kind[s] == 0) {
                grad_pair(b[j][0], b[j][2], k[j][0], k[j][2], c, sn, sg, acc[s]);
                grad_pair(b[j][1], b[j][3], k[j][1], k[j][3], c, sn, sg, acc[s]);
              } else if (gp.kind[s] == 1) {
                grad_pair(b[j][0], b[j][1], k[j][0], k[j][1], c, sn, sg, acc[s]);
                grad_pair(b[j][2], b[j][3], k[j][2], k[j][3], c, sn, sg, acc[s]);
              } else {
                grad_pair(b[j][0], b[j][3], k[j][0], k[j][3], c, sn, sg, acc[s]);
              }
            }
            double* b0 = rowp0(BB, ri[j], NB);
            double* b1 = rowp1(BB, ri[j], NB);
            double* k0 = rowp0(KB, ri[j], NB);
            double* k1 = rowp1(KB, ri[j], NB);
            b0[ib] = b[j][0]; b0[ibp] = flip(b[j][1], g01);
            b1[ib] = flip(b[j][2], g10); b1[ibp] = flip(b[j][3], g11);
            k0[ib] = k[j][0]; k0[ibp] = flip(k[j][1], g01);
            k1[ib] = flip(k[j][2], g10); k1[ibp] = flip(k[j][3], g11);
          }
        }
      }
    } else {
      // 2-amplitude tiles: (src row pair x inert column) feels only alpha steps, (inert row x src column pair) only beta
      const int want = row_src ? 0 : 1;
#pragma unroll 1
      for (int it = 0; it < TILE_ROWS / G2_ROWS_PER_ITER; ++it) {
        int4 ri[G2_ROWS_PER_ITER];
        double b0[G2_ROWS_PER_ITER], b1[G2_ROWS_PER_ITER], k0[G2_ROWS_PER_ITER], k1[G2_ROWS_PER_ITER];
#pragma unroll
        for (int j = 0; j < G2_ROWS_PER_ITER; ++j) ri[j] = __ldg(rit + it * G2_ROWS_PER_ITER + j);
#pragma unroll
        for (int j = 0; j < G2_ROWS_PER_ITER; ++j) {
          if (item_on<true>(ri[j])) {
            const double* pb0 = rowp0(BB, ri[j], NB) + ib;
            const double* pb1 = row_src ? rowp1(BB, ri[j], NB) + ib : rowp0(BB, ri[j], NB) + ibp;
            const double* pk0 = rowp0(KB, ri[j], NB) + ib;
            const double* pk1 = row_src ? rowp1(KB, ri[j], NB) + ib : rowp0(KB, ri[j], NB) + ibp;
            b0[j] = *pb0; b1[j] = *pb1; k0[j] = *pk0; k1[j] = *pk1;
          }
        }
#pragma unroll
        for (int j = 0; j < G2_ROWS_PER_ITER; ++j) {
          if (item_on<true>(ri[j])) {
            const int g = row_src ? ((ri[j].z & 1) ^ crb) : (sSb ^ ((ri[j].z >> 1) & 1));
            b1[j] = flip(b1[j], g);
            k1[j] = flip(k1[j], g);
#pragma unroll
            for (int s = 0; s < SQ_MAX_PROGRAM; ++s) {
              if (s >= gp.n) break;
              if (gp.kind[s] == want) grad_pair(b0[j], b1[j], k0[j], k1[j], gp.c[s], gp.s[s], 1.0, acc[s]);
            }
            double* pb0 = rowp0(BB, ri[j], NB) + ib;
            double* pb1 = row_src ? rowp1(BB, ri[j], NB) + ib : rowp0(BB, ri[j], NB) + ibp;
            double* pk0 = rowp0(KB, ri[j], NB) + ib;
            double* pk1 = row_src ? rowp1(KB, ri[j], NB) + ib : rowp0(KB, ri[j], NB) + ibp;
            *pb0 = b0[j]; *pb1 = flip(b1[j], g);
            *pk0 = k0[j]; *pk1 = flip(k1[j], g);
          }
        }
      }
    }
  }
  block_reduce_store<SQ_MAX_PROGRAM>(acc, partial, (int64_t)blockIdx.y * gridDim.x + blockIdx.x);
}

// sum partial[b*NS + k] over b for each k (one block per k, fixed order -> deterministic)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ partial, int64_t nblocks,
                                                             int ns, double* __restrict__ out, double scale) {
  const int k = blockIdx.x;
  double v = 0.0;
  for (int64_t b = threadIdx.x; b < nblocks; b += 256) v += partial[b * ns + k];
  __shared__ double sm[8];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sm[w];
    out[k] = scale * t;
  }
}

// out[k] = sum over blocks of partial[b * ns + k], k < n_out (fixed order: deterministic)
int sq_reduce_partials(const double* partial, int64_t nblocks, int ns, int n_out, double* out, cudaStream_t st) {
  if (n_out < 1) return SQ_OK;
  reduce_partials_kernel<<<n_out, 256, 0, st>>>(partial, nblocks, ns, out, 1.0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sq_set_error("reduce_partials_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  return SQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Generic excitation generator G (one ladder string, disjoint annihilated / created sets; single ..
// sextuple of reference operators.py:145-359): exp(theta (G - G^dagger)) is a Givens rotation on every
// (src, tgt) determinant pair.  Rows = alpha strings valid as source (with their targets), columns carry
// (partner<<1 | negative) codes for beta strings.
// mode 0: rotate in place; mode 1: out = (G - G^dagger) in  (out pre-zeroed)
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(TILE_THREADS)
gen_kernel(double* __restrict__ C, const double* __restrict__ IN, const int32_t* __restrict__ srcRows,
           const int32_t* __restrict__ tgtRows, const int8_t* __restrict__ sgnRows, int64_t n_rows,
           const int32_t* __restrict__ colCode, int64_t NB, int64_t row_begin, double c, double s) {
  const int64_t ib = (int64_t)blockIdx.x * TILE_THREADS + threadIdx.x;
  if (ib >= NB) return;
  const int32_t code = __ldg(colCode + ib);
  if (code < 0) return;
  const int64_t ibp = code >> 1;
  const double sgc = (code & 1) ? -1.0 : 1.0;
  const int64_t r0 = (int64_t)blockIdx.y * TILE_ROWS;
  const int64_t r1 = (r0 + TILE_ROWS < n_rows) ? r0 + TILE_ROWS : n_rows;
  for (int64_t r = r0; r < r1; ++r) {
    const int64_t i0 = ((int64_t)srcRows[r] - row_begin) * NB + ib;
    const int64_t i1 = ((int64_t)tgtRows[r] - row_begin) * NB + ibp;
    const double sg = sgc * (double)sgnRows[r];
    if (MODE == 0) {
      double x0 = C[i0], x1 = C[i1];
      rot(x0, x1, c, sg * s);
      C[i0] = x0;
      C[i1] = x1;
    } else {
      const double x0 = IN[i0], x1 = IN[i1];
      C[i1] = sg * x0;
      C[i0] = -sg * x1;
    }
  }
}

__global__ void __launch_bounds__(TILE_THREADS)
gen_grad_kernel(double* __restrict__ BRA, double* __restrict__ KET, const int32_t* __restrict__ srcRows,
                const int32_t* __restrict__ tgtRows, const int8_t* __restrict__ sgnRows, int64_t n_rows,
                const int32_t* __restrict__ colCode, int64_t NB, int64_t row_begin, double c, double s,
                double* __restrict__ partial) {
  double acc[1] = {0.0};
  const int64_t ib = (int64_t)blockIdx.x * TILE_THREADS + threadIdx.x;
  int32_t code = -1;
  if (ib < NB) code = __ldg(colCode + ib);
  if (code >= 0) {
    const int64_t ibp = code >> 1;
    const double sgc = (code & 1) ? -1.0 : 1.0;
    const int64_t r0 = (int64_t)blockIdx.y * TILE_ROWS;
    const int64_t r1 = (r0 + TILE_ROWS < n_rows) ? r0 + TILE_ROWS : n_rows;
    for (int64_t r = r0; r < r1; ++r) {
      const int64_t i0 = ((int64_t)srcRows[r] - row_begin) * NB + ib;
      const int64_t i1 = ((int64_t)tgtRows[r] - row_begin) * NB + ibp;
      const double sg = sgc * (double)sgnRows[r];
      double b0 = BRA[i0], b1 = BRA[i1], k0 = KET[i0], k1 = KET[i1];
      acc[0] += sg * (b1 * k0 - b0 * k1);
      rot(b0, b1, c, sg * s);
      rot(k0, k1, c, sg * s);
      BRA[i0] = b0; BRA[i1] = b1; KET[i0] = k0; KET[i1] = k1;
    }
  }
  block_reduce_store<1>(acc, partial, (int64_t)blockIdx.y * gridDim.x + blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// Gather kernel for an arbitrary FermionicOperator (list of normal-ordered strings): one thread per
// TARGET determinant, strings visited in dictionary order, so the accumulation order per element is the
// one of the reference's loops (operator_state_algebra.py:596-628 with :118-135 / :204-218) and the sum
// is deterministic.  Products are formed without FMA contraction to stay bit-comparable with the CPU.
// ---------------------------------------------------------------------------------------------
struct StringRec {
  uint32_t toccA, tempA, flipA, parA;
  uint32_t toccB, tempB, flipB, parB;
  double coeff;   // includes s0
};

#define GATHER_CHUNK 128

__global__ void __launch_bounds__(TILE_THREADS)
gather_kernel(const double* __restrict__ IN, double* __restrict__ OUT, const StringRec* __restrict__ recs,
              int n_strings, const uint32_t* __restrict__ strA, const uint32_t* __restrict__ strB,
              const int32_t* __restrict__ rankA, const int32_t* __restrict__ rankB, int64_t NB,
              int64_t row_begin, int64_t row_first, int accumulate) {
  __shared__ StringRec sm[GATHER_CHUNK];
  const int64_t ia = row_first + blockIdx.y;
  const int64_t ib = (int64_t)blockIdx.x * TILE_THREADS + threadIdx.x;
  const bool active = ib < NB;
  const uint32_t a = __ldg(strA + ia);
  const uint32_t b = active ? __ldg(strB + ib) : 0u;
  const int64_t o = (ia - row_begin) * NB + ib;
  double acc = (accumulate && active) ? OUT[o] : 0.0;
  for (int base = 0; base < n_strings; base += GATHER_CHUNK) {
    const int cnt = (n_strings - base < GATHER_CHUNK) ? n_strings - base : GATHER_CHUNK;
    __syncthreads();
    // cooperative copy of the chunk (StringRec = 40 bytes = 10 words)
    {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(recs + base);
      uint32_t* dst = reinterpret_cast<uint32_t*>(sm);
      for (int w = threadIdx.x; w < cnt * (int)(sizeof(StringRec) / 4); w += TILE_THREADS) dst[w] = src[w];
    }
    __syncthreads();
    if (!active) continue;
    for (int k = 0; k < cnt; ++k) {
      const StringRec& rc = sm[k];
      if ((a & rc.toccA) != rc.toccA || (a & rc.tempA) != 0u) continue;   // uniform per block
      if ((b & rc.toccB) != rc.toccB || (b & rc.tempB) != 0u) continue;
      const uint32_t sa = a ^ rc.flipA, sb = b ^ rc.flipB;
      const int32_t ra = __ldg(rankA + sa), rb = __ldg(rankB + sb);
      // conserving strings always land inside the space; ranks are valid by construction
      const int par = (__popc(sa & rc.parA) + __popc(sb & rc.parB)) & 1;
      const double x = IN[((int64_t)ra - row_begin) * NB + rb];
      const double f = par ? -rc.coeff : rc.coeff;
      acc = __dadd_rn(acc, __dmul_rn(f, x));
    }
  }
  if (active) OUT[o] = acc;
}

// ---------------------------------------------------------------------------------------------
// BLAS-1
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dot_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                  int64_t n, double* __restrict__ partial) {
  double v = 0.0;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) v += a[i] * b[i];
  __shared__ double sm[8];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sm[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) axpy_kernel(double alpha, const double* __restrict__ x,
                                                   double* __restrict__ y, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride)
    y[i] = __dadd_rn(y[i], __dmul_rn(alpha, x[i]));
}

__global__ void __launch_bounds__(256) scale_copy_kernel(double alpha, const double* __restrict__ x,
                                                         double* __restrict__ y, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) y[i] = alpha * x[i];
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sq_set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}

static int fill_program(const TileStep* steps, int n_steps, TileProgram* p) {
  if (n_steps < 1 || n_steps > SQ_MAX_PROGRAM) {
    sq_set_error("tile program with %d steps (max %d)", n_steps, SQ_MAX_PROGRAM);
    return SQ_ERR_INVALID;
  }
  p->n = n_steps;
  for (int k = 0; k < SQ_MAX_PROGRAM; ++k) {
    p->kind[k] = (k < n_steps) ? steps[k].kind : -1;
    p->slot[k] = k;
    p->c[k] = (k < n_steps) ? steps[k].c : 1.0;
    p->s[k] = (k < n_steps) ? steps[k].s : 0.0;
  }
  return SQ_OK;
}

// host: multiply the step rotations into the gauge-fixed 4x4 matrix and the two total 2x2 rotations
void sq_build_tile_matrices(const TileStep* steps, int n_steps, int sigma, TileMatrices* tm) {
  sq_build_tile_matrices3(steps, n_steps, 1, 1, sigma, tm);
}

void sq_build_tile_matrices3(const TileStep* steps, int n_steps, int ea, int eb, int ed, TileMatrices* tm) {
  double M[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  auto apply = [&](int u, int v, double c, double s) {   // rows u (src) and v (tgt) of M <- rotation * M
    for (int k = 0; k < 4; ++k) {
      const double a = M[u][k], b = M[v][k];
      M[u][k] = c * a - s * b;
      M[v][k] = c * b + s * a;
    }
  };
  double tha_c = 1, tha_s = 0, thb_c = 1, thb_s = 0;
  auto compose = [](double& C0, double& S0, double c, double s) {   // angle addition
    const double nc = C0 * c - S0 * s, ns = S0 * c + C0 * s;
    C0 = nc;
    S0 = ns;
  };
  for (int k = 0; k < n_steps; ++k) {
    const double c = steps[k].c, s = steps[k].s;
    if (steps[k].kind == 0) {
      apply(0, 2, c, ea * s);
      apply(1, 3, c, ea * s);
      compose(tha_c, tha_s, c, ea * s);
    } else if (steps[k].kind == 1) {
      apply(0, 1, c, eb * s);
      apply(2, 3, c, eb * s);
      compose(thb_c, thb_s, c, eb * s);
    } else {
      apply(0, 3, c, ed * s);
    }
  }
  for (int r = 0; r < 4; ++r)
    for (int k = 0; k < 4; ++k) tm->m[4 * r + k] = M[r][k];
  tm->ca = tha_c;
  tm->sa = tha_s;
  tm->cb = thb_c;
  tm->sb = thb_s;
}

static int g_tile_variant = -1;   // SQ_TILE_KERNEL=1 forces the step-loop kernel (debug / A-B comparison)

int sq_launch_tile(sq_space* sp, const PairTables& pt, const TileStep* steps, int n_steps, double* state,
                   const PeerPtrs* peers, cudaStream_t st) {
  if (pt.n_rows == 0 && pt.n_cross_items == 0) return SQ_OK;
  if (g_tile_variant < 0) {
    const char* e = getenv("SQ_TILE_KERNEL");
    g_tile_variant = (e && e[0] == '1') ? 1 : 2;
  }
  if (pt.n_cross_items > 0 && (!peers || pt.sigma == 0)) {
    sq_set_error("orbital pair (%d,%d) pairs alpha rows on different devices: use sq_ups_apply_dist with peer-mapped "
                 "shards", pt.i, pt.a);
    return SQ_ERR_UNSUPPORTED;
  }
  if ((g_tile_variant == 2 || peers) && pt.sigma != 0) {
    const bool use_peers = peers && pt.n_cross_items > 0;
    if (use_peers) {
      // (re)upload the 16-entry base-pointer table only when it changes; the staging words are pinned, so the
      // stream is drained before they are overwritten
      unsigned long long tab[SQ_MAX_WORLD];
      for (int r = 0; r < SQ_MAX_WORLD; ++r) tab[r] = (unsigned long long)(uintptr_t)peers->p[r];
      tab[sp->rank] = (unsigned long long)(uintptr_t)state;
      if (!sp->d_peer_tab) SQ_CUDA(cudaMalloc(&sp->d_peer_tab, sizeof(tab)));
      if (memcmp(tab, sp->peer_shadow, sizeof(tab)) != 0) {
        SQ_CUDA(cudaStreamSynchronize(st));
        unsigned long long* stage = reinterpret_cast<unsigned long long*>(sp->h_pinned + 4096 - SQ_MAX_WORLD);
        memcpy(stage, tab, sizeof(tab));
        memcpy(sp->peer_shadow, tab, sizeof(tab));
        SQ_CUDA(cudaMemcpyAsync(sp->d_peer_tab, stage, sizeof(tab), cudaMemcpyHostToDevice, st));
      }
    }
    if (n_steps < 1 || n_steps > SQ_MAX_PROGRAM) {
      sq_set_error("tile program with %d steps (max %d)", n_steps, SQ_MAX_PROGRAM);
      return SQ_ERR_INVALID;
    }
    TileMatrices tm;
    sq_build_tile_matrices(steps, n_steps, pt.sigma, &tm);
    bool any_single = false;
    for (int k = 0; k < n_steps; ++k) any_single |= steps[k].kind != 2;
    // a pair-double-only program touches src x src tiles only
    const int gx = any_single ? pt.n_colblk_src + pt.n_colblk_inert : pt.n_colblk_src;
    const int gy = any_single ? pt.n_rowchunk_src + pt.n_rowchunk_inert : pt.n_rowchunk_src;
    if (gx == 0 || gy == 0) return SQ_OK;
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (use_peers) {
      Bases<true> bases{state, sp->d_peer_tab};
      tile_kernel_v2<true><<<grid, TILE_THREADS, 0, st>>>(bases, pt.d_colItems, pt.n_colblk_src, pt.d_rowItems,
                                                         pt.n_rowchunk_src, sp->NB, tm);
    } else {
      Bases<false> bases{state, nullptr};
      tile_kernel_v2<false><<<grid, TILE_THREADS, 0, st>>>(bases, pt.d_colItems, pt.n_colblk_src, pt.d_rowItems,
                                                          pt.n_rowchunk_src, sp->NB, tm);
    }
    return check_launch("tile_kernel_v2");
  }
  TileProgram prog;
  SQ_CHECK(fill_program(steps, n_steps, &prog));
  dim3 grid((unsigned)((sp->NB + TILE_THREADS - 1) / TILE_THREADS), (unsigned)((pt.n_rows + TILE_ROWS - 1) / TILE_ROWS));
  tile_kernel<<<grid, TILE_THREADS, 0, st>>>(state, pt.d_codeA, pt.d_codeB, pt.d_rowsA, pt.n_rows, sp->NB,
                                            sp->row_begin, prog);
  return check_launch("tile_kernel");
}

static int finish_partials(sq_space* sp, int64_t nblocks, int ns, double scale, double* out_host, cudaStream_t st) {
  // reduce into the tail of the partial buffer, then stage through pinned memory
  double* d_out = sp->d_partial + nblocks * ns;
  reduce_partials_kernel<<<ns, 256, 0, st>>>(sp->d_partial, nblocks, ns, d_out, scale);
  SQ_CHECK(check_launch("reduce_partials_kernel"));
  SQ_CUDA(cudaMemcpyAsync(sp->h_pinned, d_out, sizeof(double) * ns, cudaMemcpyDeviceToHost, st));
  SQ_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < ns; ++k) out_host[k] = sp->h_pinned[k];
  return SQ_OK;
}

// d_out: DEVICE array of n_steps doubles receiving <bra|T_step|ket> (no host synchronisation here; the
// sweep copies the whole gradient back once at its end)
int sq_launch_tile_grad(sq_space* sp, const PairTables& pt, const TileStep* steps, int n_steps, double* bra,
                        double* ket, double* d_out, cudaStream_t st) {
  SQ_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * n_steps, st));
  if (pt.n_cross_items > 0) {
    sq_set_error("gradient sweep: orbital pair (%d,%d) pairs rows on different devices (not supported yet)", pt.i, pt.a);
    return SQ_ERR_UNSUPPORTED;
  }
  if (pt.n_rows == 0) return SQ_OK;
  if (g_tile_variant < 0) {
    const char* e = getenv("SQ_TILE_KERNEL");
    g_tile_variant = (e && e[0] == '1') ? 1 : 2;
  }
  if (n_steps < 1 || n_steps > SQ_MAX_PROGRAM) {
    sq_set_error("tile program with %d steps (max %d)", n_steps, SQ_MAX_PROGRAM);
    return SQ_ERR_INVALID;
  }
  if (g_tile_variant == 2 && pt.sigma != 0) {
    GradProgram gp;
    gp.n = n_steps;
    for (int k = 0; k < SQ_MAX_PROGRAM; ++k) {
      const bool on = k < n_steps;
      gp.kind[k] = on ? steps[k].kind : -1;
      const double sig = (on && steps[k].kind == 2) ? (double)pt.sigma : 1.0;
      gp.c[k] = on ? steps[k].c : 1.0;
      gp.s[k] = on ? sig * steps[k].s : 0.0;
      gp.sig[k] = sig;
    }
    dim3 grid((unsigned)(pt.n_colblk_src + pt.n_colblk_inert), (unsigned)(pt.n_rowchunk_src + pt.n_rowchunk_inert));
    const int64_t nblocks = (int64_t)grid.x * grid.y;
    SQ_CHECK(sq_ensure_partial(sp, nblocks * SQ_MAX_PROGRAM + SQ_MAX_PROGRAM));
    // the work lists hold row indices relative to the shard start, so the kernel takes the shard base pointers
    tile_grad_kernel_v2<<<grid, TILE_THREADS, 0, st>>>(bra, ket, pt.d_colItems, pt.n_colblk_src, pt.d_rowItems,
                                                      pt.n_rowchunk_src, sp->NB, gp, sp->d_partial);
    SQ_CHECK(check_launch("tile_grad_kernel_v2"));
    reduce_partials_kernel<<<n_steps, 256, 0, st>>>(sp->d_partial, nblocks, SQ_MAX_PROGRAM, d_out, 1.0);
    return check_launch("reduce_partials_kernel");
  }
  TileProgram prog;
  SQ_CHECK(fill_program(steps, n_steps, &prog));
  dim3 grid((unsigned)((sp->NB + TILE_THREADS - 1) / TILE_THREADS), (unsigned)((pt.n_rows + TILE_ROWS - 1) / TILE_ROWS));
  const int64_t nblocks = (int64_t)grid.x * grid.y;
  SQ_CHECK(sq_ensure_partial(sp, nblocks * SQ_MAX_PROGRAM + SQ_MAX_PROGRAM));
  tile_grad_kernel<<<grid, TILE_THREADS, 0, st>>>(bra, ket, pt.d_codeA, pt.d_codeB, pt.d_rowsA, pt.n_rows, sp->NB,
                                                 sp->row_begin, prog, sp->d_partial);
  SQ_CHECK(check_launch("tile_grad_kernel"));
  reduce_partials_kernel<<<n_steps, 256, 0, st>>>(sp->d_partial, nblocks, SQ_MAX_PROGRAM, d_out, 1.0);
  return check_launch("reduce_partials_kernel");
}

// Exchange operator on a sharded (bra, ket) pair: d_tab_bra / d_tab_ket = DEVICE tables (SQ_MAX_WORLD entries) of the shard base
// pointers of every rank as mapped into this process (own rank = the local shard).  d_out as in sq_launch_tile_grad.
int sq_launch_tile_grad_peer(sq_space* sp, const PairTables& pt, const TileStep* steps, int n_steps, double* bra, double* ket,
                             const unsigned long long* d_tab_bra, const unsigned long long* d_tab_ket, double* d_out,
                             cudaStream_t st) {
  SQ_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * n_steps, st));
  if (pt.n_rows == 0 && pt.n_cross_items == 0) return SQ_OK;
  if (pt.sigma == 0) {
    sq_set_error("gradient sweep: orbital pair (%d,%d) has no gauge-fixed work lists", pt.i, pt.a);
    return SQ_ERR_UNSUPPORTED;
  }
  if (n_steps < 1 || n_steps > SQ_MAX_PROGRAM) {
    sq_set_error("tile program with %d steps (max %d)", n_steps, SQ_MAX_PROGRAM);
    return SQ_ERR_INVALID;
  }
  GradProgram gp;
  gp.n = n_steps;
  for (int k = 0; k < SQ_MAX_PROGRAM; ++k) {
    const bool on = k < n_steps;
    gp.kind[k] = on ? steps[k].kind : -1;
    const double sig = (on && steps[k].kind == 2) ? (double)pt.sigma : 1.0;
    gp.c[k] = on ? steps[k].c : 1.0;
    gp.s[k] = on ? sig * steps[k].s : 0.0;
    gp.sig[k] = sig;
  }
  const int gx = pt.n_colblk_src + pt.n_colblk_inert, gy = pt.n_rowchunk_src + pt.n_rowchunk_inert;
  if (gx == 0 || gy == 0) return SQ_OK;
  dim3 grid((unsigned)gx, (unsigned)gy);
  const int64_t nblocks = (int64_t)grid.x * grid.y;
  SQ_CHECK(sq_ensure_partial(sp, nblocks * SQ_MAX_PROGRAM + SQ_MAX_PROGRAM));
  Bases<true> BB{bra, d_tab_bra}, KB{ket, d_tab_ket};
  tile_grad_peer_kernel<<<grid, TILE_THREADS, 0, st>>>(BB, KB, pt.d_colItems, pt.n_colblk_src, pt.d_rowItems, pt.n_rowchunk_src,
                                                      sp->NB, gp, sp->d_partial);
  SQ_CHECK(check_launch("tile_grad_peer_kernel"));
  reduce_partials_kernel<<<n_steps, 256, 0, st>>>(sp->d_partial, nblocks, SQ_MAX_PROGRAM, d_out, 1.0);
  return check_launch("reduce_partials_kernel");
}

int sq_launch_gen_rot(sq_space* sp, const GenTables& gt, double c, double s, double* state, cudaStream_t st) {
  if (gt.n_rows == 0 || gt.n_cols_valid == 0) return SQ_OK;
  dim3 grid((unsigned)((sp->NB + TILE_THREADS - 1) / TILE_THREADS), (unsigned)((gt.n_rows + TILE_ROWS - 1) / TILE_ROWS));
  gen_kernel<0><<<grid, TILE_THREADS, 0, st>>>(state, nullptr, gt.d_srcRows, gt.d_tgtRows, gt.d_sgnRows, gt.n_rows,
                                              gt.d_colCode, sp->NB, sp->row_begin, c, s);
  return check_launch("gen_kernel<rot>");
}

int sq_launch_gen_apply(sq_space* sp, const GenTables& gt, const double* in, double* out, cudaStream_t st) {
  SQ_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)sp->local_len(), st));
  if (gt.n_rows == 0 || gt.n_cols_valid == 0) return SQ_OK;
  dim3 grid((unsigned)((sp->NB + TILE_THREADS - 1) / TILE_THREADS), (unsigned)((gt.n_rows + TILE_ROWS - 1) / TILE_ROWS));
  gen_kernel<1><<<grid, TILE_THREADS, 0, st>>>(out, in, gt.d_srcRows, gt.d_tgtRows, gt.d_sgnRows, gt.n_rows,
                                              gt.d_colCode, sp->NB, sp->row_begin, 1.0, 0.0);
  return check_launch("gen_kernel<apply>");
}

int sq_launch_gen_grad(sq_space* sp, const GenTables& gt, double c, double s, double* bra, double* ket,
                       double* d_out, cudaStream_t st) {
  SQ_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), st));
  if (gt.n_rows == 0 || gt.n_cols_valid == 0) return SQ_OK;
  dim3 grid((unsigned)((sp->NB + TILE_THREADS - 1) / TILE_THREADS), (unsigned)((gt.n_rows + TILE_ROWS - 1) / TILE_ROWS));
  const int64_t nblocks = (int64_t)grid.x * grid.y;
  SQ_CHECK(sq_ensure_partial(sp, nblocks + 1));
  gen_grad_kernel<<<grid, TILE_THREADS, 0, st>>>(bra, ket, gt.d_srcRows, gt.d_tgtRows, gt.d_sgnRows, gt.n_rows,
                                                gt.d_colCode, sp->NB, sp->row_begin, c, s, sp->d_partial);
  SQ_CHECK(check_launch("gen_grad_kernel"));
  reduce_partials_kernel<<<1, 256, 0, st>>>(sp->d_partial, nblocks, 1, d_out, 1.0);
  return check_launch("reduce_partials_kernel");
}

int sq_launch_gather(sq_space* sp, const std::vector<StringAction>& strings, const std::vector<double>& coeffs,
                     const double* in, double* out, int accumulate, cudaStream_t st) {
  const int64_t rows = sp->row_end - sp->row_begin;
  if (rows == 0) return SQ_OK;
  const int n = (int)strings.size();
  if (n == 0) {
    if (!accumulate) SQ_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)sp->local_len(), st));
    return SQ_OK;
  }
  std::vector<StringRec> recs(n);
  for (int k = 0; k < n; ++k) {
    const StringAction& a = strings[k];
    recs[k] = {a.toccA, a.tempA, a.flipA, a.parA, a.toccB, a.tempB, a.flipB, a.parB, coeffs[k] * a.s0};
  }
  StringRec* d_recs = nullptr;
  SQ_CUDA(cudaMallocAsync(&d_recs, sizeof(StringRec) * n, st));
  SQ_CUDA(cudaMemcpyAsync(d_recs, recs.data(), sizeof(StringRec) * n, cudaMemcpyHostToDevice, st));
  int status = SQ_OK;
  // gridDim.y is limited to 65535 rows per launch
  for (int64_t r = 0; r < rows && status == SQ_OK; r += 65535) {
    const int64_t nr = (rows - r < 65535) ? rows - r : 65535;
    dim3 grid((unsigned)((sp->NB + TILE_THREADS - 1) / TILE_THREADS), (unsigned)nr);
    gather_kernel<<<grid, TILE_THREADS, 0, st>>>(in, out, d_recs, n, sp->d_strA, sp->d_strB, sp->d_rankA, sp->d_rankB,
                                                sp->NB, sp->row_begin, sp->row_begin + r, accumulate);
    status = check_launch("gather_kernel");
  }
  // the host vector `recs` must outlive the async copy
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFreeAsync(d_recs, st);
  if (status == SQ_OK && e != cudaSuccess) {
    sq_set_error("gather_kernel failed: %s", cudaGetErrorString(e));
    status = SQ_ERR_CUDA;
  }
  return status;
}

int sq_launch_dot(sq_space* sp, const double* a, const double* b, double* out_host, cudaStream_t st) {
  const int64_t n = sp->local_len();
  int64_t nb = (n + 256 * 8 - 1) / (256 * 8);
  if (nb > 148 * 16) nb = 148 * 16;
  if (nb < 1) nb = 1;
  SQ_CHECK(sq_ensure_partial(sp, nb + 1));
  dot_kernel<<<(unsigned)nb, 256, 0, st>>>(a, b, n, sp->d_partial);
  SQ_CHECK(check_launch("dot_kernel"));
  return finish_partials(sp, nb, 1, 1.0, out_host, st);
}

int sq_launch_axpy(sq_space* sp, double alpha, const double* x, double* y, cudaStream_t st) {
  const int64_t n = sp->local_len();
  if (n == 0) return SQ_OK;
  int64_t nb = (n + 255) / 256;
  if (nb > 148 * 32) nb = 148 * 32;
  axpy_kernel<<<(unsigned)nb, 256, 0, st>>>(alpha, x, y, n);
  return check_launch("axpy_kernel");
}

int sq_launch_scale_copy(sq_space* sp, double alpha, const double* x, double* y, cudaStream_t st) {
  const int64_t n = sp->local_len();
  if (n == 0) return SQ_OK;
  int64_t nb = (n + 255) / 256;
  if (nb > 148 * 32) nb = 148 * 32;
  scale_copy_kernel<<<(unsigned)nb, 256, 0, st>>>(alpha, x, y, n);
  return check_launch("scale_copy_kernel");
}
