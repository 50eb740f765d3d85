// Two commuting tUPS bricks in ONE sweep ("quad" tiles).
//
// Bricks on disjoint orbital pairs P1 = (i1,a1) and P2 = (i2,a2) commute (reference util.py:694-745: every
// half layer of the brick wall is such a set, and so are the last brick of one half layer and the first of
// the next).  For fixed occupation of all other orbitals the amplitudes form a tile
// {alpha group} x {beta group}, each group = 1, 2 or 4 strings depending on whether the string has exactly
// one electron in P1 and/or in P2; a thread keeps the tile (up to 4 x 4 amplitudes) in registers, applies
// brick 1 (4x4 gauge-fixed matrix or 2x2 rotation on the P1 index, for every value of the P2 index) and then
// brick 2, and writes it back.  One read + one write of the vector for SIX ansatz operators; almost no
// amplitude is inert in both pairs, so the 32-byte sectors that an isolated high-orbital brick drags along
// without using are used here.
//
// Data movement: a batch always carries 16 amplitudes per thread (one 4x4 tile, two 2x4 tiles, ... eight
// 1x2 tiles), all loads of a batch are issued before the first store.  QUAD_ASYNC = 1 fetches through
// cp.async into per-thread shared-memory slots with a multi-stage pipeline instead of registers; on B200
// the 8-byte LDGSTS path turned out to be issue-limited (~8 cycles per warp instruction), so plain LDG is the
// default.
#include <cstdio>
#include <cstdlib>

#include <cuda_pipeline.h>

#include "sqsv_internal.h"

#ifndef QUAD_ASYNC
#define QUAD_ASYNC 0
#endif
#if QUAD_ASYNC
#define QUAD_THREADS 128
#define QUAD_MINBLOCKS 1
#else
#ifndef QUAD_THREADS
#define QUAD_THREADS 256
#endif
#ifndef QUAD_MINBLOCKS
#define QUAD_MINBLOCKS 3     // <= 85 registers per thread -> 24 resident warps per SM
#endif
#endif
#ifndef QUAD_ROWS
#define QUAD_ROWS 16   // row groups per CTA
#endif
#define QUAD_STAGES 3

struct QuadMats {
  double m1[16], m2[16];   // gauge-fixed 4x4 brick matrices, basis (x00, x01, x10, x11) = (row, column) index in the pair
  double ca1, sa1, cb1, sb1, ca2, sa2, cb2, sb2;   // total alpha / beta single rotations of each brick
};

__device__ __forceinline__ double qflip(double x, int neg) {
  return __hiloint2double(__double2hiint(x) ^ (neg << 31), __double2loint(x));
}
__device__ __forceinline__ int comp(const int4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }

__device__ __forceinline__ void mat4(const double* __restrict__ m, double& y0, double& y1, double& y2, double& y3) {
  const double z0 = m[0] * y0 + m[1] * y1 + m[2] * y2 + m[3] * y3;
  const double z1 = m[4] * y0 + m[5] * y1 + m[6] * y2 + m[7] * y3;
  const double z2 = m[8] * y0 + m[9] * y1 + m[10] * y2 + m[11] * y3;
  const double z3 = m[12] * y0 + m[13] * y1 + m[14] * y2 + m[15] * y3;
  y0 = z0; y1 = z1; y2 = z2; y3 = z3;
}
__device__ __forceinline__ void rot2(double& a, double& b, double c, double s, int g) {
  const double x = a, y = qflip(b, g);
  a = c * x - s * y;
  b = qflip(c * y + s * x, g);
}

// R1/R2: the row group is active in pair 1 / 2; C1/C2: same for the column group.
// Flat tile index: ((a1 * NR2 + a2) * NC1 + b1) * NC2 + b2.
template <bool R1, bool R2, bool C1, bool C2>
struct Shape {
  static constexpr int NR1 = R1 ? 2 : 1, NR2 = R2 ? 2 : 1, NC1 = C1 ? 2 : 1, NC2 = C2 ? 2 : 1;
  static constexpr int TILE = NR1 * NR2 * NC1 * NC2;
  static constexpr int G = (16 / TILE > 8) ? 8 : 16 / TILE;   // row groups per batch
  __device__ static constexpr int at(int a1, int a2, int b1, int b2) { return ((a1 * NR2 + a2) * NC1 + b1) * NC2 + b2; }
};

template <bool R1, bool R2, bool C1, bool C2>
__device__ __forceinline__ void quad_load(const double* __restrict__ C, int64_t NB, const int4 rw, const int4 cl,
                                          double* __restrict__ x) {
  using S = Shape<R1, R2, C1, C2>;
#pragma unroll
  for (int a1 = 0; a1 < S::NR1; ++a1)
#pragma unroll
    for (int a2 = 0; a2 < S::NR2; ++a2) {
      const double* row = C + (int64_t)comp(rw, a1 * 2 + a2) * NB;
#pragma unroll
      for (int b1 = 0; b1 < S::NC1; ++b1)
#pragma unroll
        for (int b2 = 0; b2 < S::NC2; ++b2) x[S::at(a1, a2, b1, b2)] = row[comp(cl, b1 * 2 + b2)];
    }
}

template <bool R1, bool R2, bool C1, bool C2>
__device__ __forceinline__ void quad_issue(const double* __restrict__ C, int64_t NB, const int4 rw, const int4 cl,
                                           double* __restrict__ sm) {
  using S = Shape<R1, R2, C1, C2>;
#pragma unroll
  for (int a1 = 0; a1 < S::NR1; ++a1)
#pragma unroll
    for (int a2 = 0; a2 < S::NR2; ++a2) {
      const double* row = C + (int64_t)comp(rw, a1 * 2 + a2) * NB;
#pragma unroll
      for (int b1 = 0; b1 < S::NC1; ++b1)
#pragma unroll
        for (int b2 = 0; b2 < S::NC2; ++b2)
          __pipeline_memcpy_async(sm + S::at(a1, a2, b1, b2) * QUAD_THREADS, row + comp(cl, b1 * 2 + b2), sizeof(double));
    }
}

template <bool R1, bool R2, bool C1, bool C2>
__device__ __forceinline__ void quad_store(double* __restrict__ C, int64_t NB, const int4 rw, const int4 cl,
                                           const double* __restrict__ x) {
  using S = Shape<R1, R2, C1, C2>;
#pragma unroll
  for (int a1 = 0; a1 < S::NR1; ++a1)
#pragma unroll
    for (int a2 = 0; a2 < S::NR2; ++a2) {
      double* row = C + (int64_t)comp(rw, a1 * 2 + a2) * NB;
#pragma unroll
      for (int b1 = 0; b1 < S::NC1; ++b1)
#pragma unroll
        for (int b2 = 0; b2 < S::NC2; ++b2) row[comp(cl, b1 * 2 + b2)] = x[S::at(a1, a2, b1, b2)];
    }
}

// both bricks on a tile held in registers
template <bool R1, bool R2, bool C1, bool C2>
__device__ __forceinline__ void quad_apply(double* __restrict__ x, const int rf, const int cf, const QuadMats& qm) {
  using S = Shape<R1, R2, C1, C2>;
  constexpr int H1 = S::NR1 - 1, H2 = S::NR2 - 1, K1 = S::NC1 - 1, K2 = S::NC2 - 1;   // index of the "tgt" member (0 if inert)
  // ---- brick 1: acts on (a1, b1) for every (a2, b2) ----
  {
    const int sSa = rf & 1, cra = (rf >> 1) & 1, crap = (rf >> 2) & 1;
    const int sSb = cf & 1, crb = (cf >> 1) & 1;
    if (R1 && C1) {
      const int g10 = sSa ^ crb, g01 = sSb ^ cra, g11 = g10 ^ sSb ^ crap;
#pragma unroll
      for (int a2 = 0; a2 < S::NR2; ++a2)
#pragma unroll
        for (int b2 = 0; b2 < S::NC2; ++b2) {
          double y0 = x[S::at(0, a2, 0, b2)], y1 = qflip(x[S::at(0, a2, K1, b2)], g01),
                 y2 = qflip(x[S::at(H1, a2, 0, b2)], g10), y3 = qflip(x[S::at(H1, a2, K1, b2)], g11);
          mat4(qm.m1, y0, y1, y2, y3);
          x[S::at(0, a2, 0, b2)] = y0;
          x[S::at(0, a2, K1, b2)] = qflip(y1, g01);
          x[S::at(H1, a2, 0, b2)] = qflip(y2, g10);
          x[S::at(H1, a2, K1, b2)] = qflip(y3, g11);
        }
    } else if (R1) {
      const int g = sSa ^ crb;   // alpha single on a column that is inert in pair 1
#pragma unroll
      for (int a2 = 0; a2 < S::NR2; ++a2)
#pragma unroll
        for (int b2 = 0; b2 < S::NC2; ++b2) rot2(x[S::at(0, a2, 0, b2)], x[S::at(H1, a2, 0, b2)], qm.ca1, qm.sa1, g);
    } else if (C1) {
      const int g = sSb ^ cra;   // beta single on a row that is inert in pair 1
#pragma unroll
      for (int a2 = 0; a2 < S::NR2; ++a2)
#pragma unroll
        for (int b2 = 0; b2 < S::NC2; ++b2) rot2(x[S::at(0, a2, 0, b2)], x[S::at(0, a2, K1, b2)], qm.cb1, qm.sb1, g);
    }
  }
  // ---- brick 2: acts on (a2, b2) for every (a1, b1) ----
  {
    const int sSa = (rf >> 4) & 1, cra = (rf >> 5) & 1, crap = (rf >> 6) & 1;
    const int sSb = (cf >> 4) & 1, crb = (cf >> 5) & 1;
    if (R2 && C2) {
      const int g10 = sSa ^ crb, g01 = sSb ^ cra, g11 = g10 ^ sSb ^ crap;
#pragma unroll
      for (int a1 = 0; a1 < S::NR1; ++a1)
#pragma unroll
        for (int b1 = 0; b1 < S::NC1; ++b1) {
          double y0 = x[S::at(a1, 0, b1, 0)], y1 = qflip(x[S::at(a1, 0, b1, K2)], g01),
                 y2 = qflip(x[S::at(a1, H2, b1, 0)], g10), y3 = qflip(x[S::at(a1, H2, b1, K2)], g11);
          mat4(qm.m2, y0, y1, y2, y3);
          x[S::at(a1, 0, b1, 0)] = y0;
          x[S::at(a1, 0, b1, K2)] = qflip(y1, g01);
          x[S::at(a1, H2, b1, 0)] = qflip(y2, g10);
          x[S::at(a1, H2, b1, K2)] = qflip(y3, g11);
        }
    } else if (R2) {
      const int g = sSa ^ crb;
#pragma unroll
      for (int a1 = 0; a1 < S::NR1; ++a1)
#pragma unroll
        for (int b1 = 0; b1 < S::NC1; ++b1) rot2(x[S::at(a1, 0, b1, 0)], x[S::at(a1, H2, b1, 0)], qm.ca2, qm.sa2, g);
    } else if (C2) {
      const int g = sSb ^ cra;
#pragma unroll
      for (int a1 = 0; a1 < S::NR1; ++a1)
#pragma unroll
        for (int b1 = 0; b1 < S::NC1; ++b1) rot2(x[S::at(a1, 0, b1, 0)], x[S::at(a1, 0, b1, K2)], qm.cb2, qm.sb2, g);
    }
  }
}

template <bool R1, bool R2, bool C1, bool C2>
__device__ __forceinline__ void quad_rows(double* __restrict__ C, int64_t NB, const int4* __restrict__ rowIdx,
                                          const int* __restrict__ rowFlags, int64_t r0, const int4 cl, const int cf,
                                          const QuadMats& qm) {
  using S = Shape<R1, R2, C1, C2>;
  constexpr int TILE = S::TILE, G = S::G, NBATCH = QUAD_ROWS / G;
#if QUAD_ASYNC
  extern __shared__ double quad_smem[];
  double* mine = quad_smem + threadIdx.x;   // slot s of stage t lives at mine[(t * 16 + s) * QUAD_THREADS]
  auto issue_batch = [&](int b) {
    double* st = mine + (b % QUAD_STAGES) * 16 * QUAD_THREADS;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int4 rw = __ldg(rowIdx + r0 + b * G + g);
      if (rw.x >= 0) quad_issue<R1, R2, C1, C2>(C, NB, rw, cl, st + g * TILE * QUAD_THREADS);
    }
  };
#pragma unroll
  for (int b = 0; b < QUAD_STAGES - 1; ++b) {
    if (b < NBATCH) issue_batch(b);
    __pipeline_commit();
  }
#pragma unroll 1
  for (int b = 0; b < NBATCH; ++b) {
    const int bn = b + QUAD_STAGES - 1;
    if (bn < NBATCH) issue_batch(bn);
    __pipeline_commit();
    __pipeline_wait_prior(QUAD_STAGES - 1);
    const double* st = mine + (b % QUAD_STAGES) * 16 * QUAD_THREADS;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int4 rw = __ldg(rowIdx + r0 + b * G + g);
      if (rw.x < 0) continue;
      double x[TILE];
#pragma unroll
      for (int k = 0; k < TILE; ++k) x[k] = st[(g * TILE + k) * QUAD_THREADS];
      quad_apply<R1, R2, C1, C2>(x, __ldg(rowFlags + r0 + b * G + g), cf, qm);
      quad_store<R1, R2, C1, C2>(C, NB, rw, cl, x);
    }
  }
#else
#pragma unroll 1
  for (int b = 0; b < NBATCH; ++b) {
    int4 rw[G];
    double x[G][TILE];
#pragma unroll
    for (int g = 0; g < G; ++g) rw[g] = __ldg(rowIdx + r0 + b * G + g);
#pragma unroll
    for (int g = 0; g < G; ++g)
      if (rw[g].x >= 0) quad_load<R1, R2, C1, C2>(C, NB, rw[g], cl, x[g]);
#pragma unroll
    for (int g = 0; g < G; ++g)
      if (rw[g].x >= 0) {
        quad_apply<R1, R2, C1, C2>(x[g], __ldg(rowFlags + r0 + b * G + g), cf, qm);
        quad_store<R1, R2, C1, C2>(C, NB, rw[g], cl, x[g]);
      }
  }
#endif
}

struct QuadBounds { int c3, c2, c1, c0; int r3, r2, r1, r0; };   // cumulative CTA / row-chunk counts per type

__global__ void __launch_bounds__(QUAD_THREADS, QUAD_MINBLOCKS)
quad_kernel(double* __restrict__ C, const int4* __restrict__ colIdx, const int* __restrict__ colFlags,
            const int4* __restrict__ rowIdx, const int* __restrict__ rowFlags, int64_t NB, const QuadBounds qb,
            const QuadMats qm) {
  const int bx = blockIdx.x, by = blockIdx.y;
  const int ct = bx < qb.c3 ? 3 : (bx < qb.c2 ? 2 : (bx < qb.c1 ? 1 : 0));
  const int rt = by < qb.r3 ? 3 : (by < qb.r2 ? 2 : (by < qb.r1 ? 1 : 0));
  if (ct == 0 && rt == 0) return;
  const int64_t ci = (int64_t)bx * QUAD_THREADS + threadIdx.x;
  const int4 cl = __ldg(colIdx + ci);
  if (cl.x < 0) return;
  const int cf = __ldg(colFlags + ci);
  const int64_t r0 = (int64_t)by * QUAD_ROWS;
  switch (rt * 4 + ct) {
    case 15: quad_rows<true, true, true, true>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    case 14: quad_rows<true, true, true, false>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    case 13: quad_rows<true, true, false, true>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    case 12: quad_rows<true, true, false, false>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    case 11: quad_rows<true, false, true, true>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    case 10: quad_rows<true, false, true, false>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    case 9: quad_rows<true, false, false, true>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    case 8: quad_rows<true, false, false, false>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    case 7: quad_rows<false, true, true, true>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    case 6: quad_rows<false, true, true, false>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    case 5: quad_rows<false, true, false, true>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    case 4: quad_rows<false, true, false, false>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    case 3: quad_rows<false, false, true, true>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    case 2: quad_rows<false, false, true, false>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    case 1: quad_rows<false, false, false, true>(C, NB, rowIdx, rowFlags, r0, cl, cf, qm); break;
    default: break;
  }
}

// ---------------------------------------------------------------------------------------------
// Gradient sweep on quad tiles: the SAME tiles for two vectors (bra, ket).  For every rotation step of brick 1 and then of
// brick 2 (reference ups_wavefunction.py:1114-1138): accumulate <bra|T_step|ket> on the tile, then rotate both vectors.  The two
// bricks commute and map every quad tile onto itself, so "brick 2 after brick 1" holds tile by tile.  One read + one write of
// bra and ket for SIX ansatz operators (tile_grad_kernel_v2: one per three).  Gauge flips as in quad_apply: inside a brick's
// 4-amplitude group every alpha / beta generator element is +1 (sigma on the pair double, folded into GradSteps::sig / s).
// ---------------------------------------------------------------------------------------------
#ifndef QGRAD_FULL_BATCH
#define QGRAD_FULL_BATCH 0   // 1: full batches for every tile shape -- measured: no change (173.5 vs 173.2 ms)
#endif
struct GradSteps {
  int n;
  int kind[SQ_MAX_PROGRAM];
  double c[SQ_MAX_PROGRAM], s[SQ_MAX_PROGRAM], sig[SQ_MAX_PROGRAM];
};
struct QuadGradProg { GradSteps g1, g2; };

__device__ __forceinline__ void qgrad_pair(double& bs, double& bt, double& ks, double& kt, double c, double s, double sig, double& acc) {
  acc += sig * (bt * ks - bs * kt);
  const double a = bs, b = bt, p = ks, q = kt;
  bs = c * a - s * b;
  bt = c * b + s * a;
  ks = c * p - s * q;
  kt = c * q + s * p;
}

// One brick of a quad tile, differentiated step by step.  PAIR1 = true: the brick acts on the (a1, b1) index for every (a2, b2);
// false: on (a2, b2) for every (a1, b1).  RA / CA: the row / column group is active in this brick's pair.  The step loop is NOT
// unrolled (the fully unrolled version had 30 000 instructions for the 16 tile shapes and lived in instruction-cache misses);
// the value of step s is added to this thread's accumulator in shared memory (acc[s * QUAD_THREADS]): 32 registers less than
// sixteen register accumulators.
template <bool R1, bool R2, bool C1, bool C2, bool PAIR1>
__device__ __forceinline__ void quad_grad_brick(double* __restrict__ xb, double* __restrict__ xk, const int rfl, const int cfl,
                                                const GradSteps& gs, double* __restrict__ acc) {
  using S = Shape<R1, R2, C1, C2>;
  constexpr bool RA = PAIR1 ? R1 : R2, CA = PAIR1 ? C1 : C2;
  if (!RA && !CA) return;
  constexpr int NO_R = PAIR1 ? S::NR2 : S::NR1, NO_C = PAIR1 ? S::NC2 : S::NC1;   // sizes of the OTHER pair's indices
  constexpr int HR = RA ? 1 : 0, KC = CA ? 1 : 0;
  auto at = [](int r, int ro, int c, int co) { return PAIR1 ? S::at(r, ro, c, co) : S::at(ro, r, co, c); };
  const int sSa = rfl & 1, cra = (rfl >> 1) & 1, crap = (rfl >> 2) & 1;
  const int sSb = cfl & 1, crb = (cfl >> 1) & 1;
  const int g10 = sSa ^ crb, g01 = sSb ^ cra, g11 = g10 ^ sSb ^ crap;
  auto flips = [&]() {
#pragma unroll
    for (int ro = 0; ro < NO_R; ++ro)
#pragma unroll
      for (int co = 0; co < NO_C; ++co) {
        if (RA && CA) {
          xb[at(0, ro, 1, co)] = qflip(xb[at(0, ro, 1, co)], g01); xk[at(0, ro, 1, co)] = qflip(xk[at(0, ro, 1, co)], g01);
          xb[at(1, ro, 0, co)] = qflip(xb[at(1, ro, 0, co)], g10); xk[at(1, ro, 0, co)] = qflip(xk[at(1, ro, 0, co)], g10);
          xb[at(1, ro, 1, co)] = qflip(xb[at(1, ro, 1, co)], g11); xk[at(1, ro, 1, co)] = qflip(xk[at(1, ro, 1, co)], g11);
        } else if (RA) {
          xb[at(1, ro, 0, co)] = qflip(xb[at(1, ro, 0, co)], g10); xk[at(1, ro, 0, co)] = qflip(xk[at(1, ro, 0, co)], g10);
        } else {
          xb[at(0, ro, 1, co)] = qflip(xb[at(0, ro, 1, co)], g01); xk[at(0, ro, 1, co)] = qflip(xk[at(0, ro, 1, co)], g01);
        }
      }
  };
  flips();
#pragma unroll 1
  for (int s = 0; s < gs.n; ++s) {
    const int kind = gs.kind[s];
    const double c = gs.c[s], sn = gs.s[s], sg = gs.sig[s];
    double a = 0.0;
    if (kind == 0) {
      if (RA) {
#pragma unroll
        for (int ro = 0; ro < NO_R; ++ro)
#pragma unroll
          for (int co = 0; co < NO_C; ++co)
#pragma unroll
            for (int cc = 0; cc <= KC; ++cc)
              qgrad_pair(xb[at(0, ro, cc, co)], xb[at(HR, ro, cc, co)], xk[at(0, ro, cc, co)], xk[at(HR, ro, cc, co)], c, sn, 1.0, a);
      }
    } else if (kind == 1) {
      if (CA) {
#pragma unroll
        for (int ro = 0; ro < NO_R; ++ro)
#pragma unroll
          for (int co = 0; co < NO_C; ++co)
#pragma unroll
            for (int rr = 0; rr <= HR; ++rr)
              qgrad_pair(xb[at(rr, ro, 0, co)], xb[at(rr, ro, KC, co)], xk[at(rr, ro, 0, co)], xk[at(rr, ro, KC, co)], c, sn, 1.0, a);
      }
    } else {
      if (RA && CA) {
#pragma unroll
        for (int ro = 0; ro < NO_R; ++ro)
#pragma unroll
          for (int co = 0; co < NO_C; ++co)
            qgrad_pair(xb[at(0, ro, 0, co)], xb[at(HR, ro, KC, co)], xk[at(0, ro, 0, co)], xk[at(HR, ro, KC, co)], c, sn, sg, a);
      }
    }
    acc[s * QUAD_THREADS] += a;   // this thread's accumulator of step s (shared memory, one column per thread: no bank conflicts)
  }
  flips();
}

template <bool R1, bool R2, bool C1, bool C2>
__device__ __forceinline__ void quad_grad_apply(double* __restrict__ xb, double* __restrict__ xk, const int rf, const int cf,
                                                const QuadGradProg& qp, double* __restrict__ acc) {
  quad_grad_brick<R1, R2, C1, C2, true>(xb, xk, rf & 7, cf & 7, qp.g1, acc);
  quad_grad_brick<R1, R2, C1, C2, false>(xb, xk, (rf >> 4) & 7, (cf >> 4) & 7, qp.g2, acc + SQ_MAX_PROGRAM * QUAD_THREADS);
}

template <bool R1, bool R2, bool C1, bool C2>
__device__ __forceinline__ void quad_rows_grad(double* __restrict__ BRA, double* __restrict__ KET, int64_t NB,
                                               const int4* __restrict__ rowIdx, const int* __restrict__ rowFlags, int64_t r0,
                                               const int4 cl, const int cf, const QuadGradProg& qp, double* __restrict__ acc) {
  using S = Shape<R1, R2, C1, C2>;
  constexpr int TILE = S::TILE;
  constexpr int GH = QGRAD_FULL_BATCH ? S::G : (S::G > 1 ? S::G / 2 : 1);   // amplitudes of each vector in flight per thread: 16, or 8 (16 for the 4 x 4 tile)
  constexpr int NBATCH = QUAD_ROWS / GH;
#pragma unroll 1
  for (int b = 0; b < NBATCH; ++b) {
    int4 rw[GH];
    double xb[GH][TILE], xk[GH][TILE];
#pragma unroll
    for (int g = 0; g < GH; ++g) rw[g] = __ldg(rowIdx + r0 + b * GH + g);
#pragma unroll
    for (int g = 0; g < GH; ++g)
      if (rw[g].x >= 0) {
        quad_load<R1, R2, C1, C2>(BRA, NB, rw[g], cl, xb[g]);
        quad_load<R1, R2, C1, C2>(KET, NB, rw[g], cl, xk[g]);
      }
#pragma unroll
    for (int g = 0; g < GH; ++g)
      if (rw[g].x >= 0) {
        quad_grad_apply<R1, R2, C1, C2>(xb[g], xk[g], __ldg(rowFlags + r0 + b * GH + g), cf, qp, acc);
        quad_store<R1, R2, C1, C2>(BRA, NB, rw[g], cl, xb[g]);
        quad_store<R1, R2, C1, C2>(KET, NB, rw[g], cl, xk[g]);
      }
  }
}

#define QGRAD_NS (2 * SQ_MAX_PROGRAM)
template <int MINB>
__global__ void __launch_bounds__(QUAD_THREADS, MINB)
quad_grad_kernel(double* __restrict__ BRA, double* __restrict__ KET, const int4* __restrict__ colIdx, const int* __restrict__ colFlags,
                 const int4* __restrict__ rowIdx, const int* __restrict__ rowFlags, int64_t NB, const QuadBounds qb,
                 const QuadGradProg qp, double* __restrict__ partial) {
  __shared__ double sacc[QGRAD_NS * QUAD_THREADS];   // [step value][thread]
  double* const acc = sacc + threadIdx.x;
#pragma unroll
  for (int k = 0; k < QGRAD_NS; ++k) acc[k * QUAD_THREADS] = 0.0;
  const int bx = blockIdx.x, by = blockIdx.y;
  const int ct = bx < qb.c3 ? 3 : (bx < qb.c2 ? 2 : (bx < qb.c1 ? 1 : 0));
  const int rt = by < qb.r3 ? 3 : (by < qb.r2 ? 2 : (by < qb.r1 ? 1 : 0));
  const int64_t ci = (int64_t)bx * QUAD_THREADS + threadIdx.x;
  const int4 cl = __ldg(colIdx + ci);
  if ((ct != 0 || rt != 0) && cl.x >= 0) {
    const int cf = __ldg(colFlags + ci);
    const int64_t r0 = (int64_t)by * QUAD_ROWS;
    switch (rt * 4 + ct) {
      case 15: quad_rows_grad<true, true, true, true>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      case 14: quad_rows_grad<true, true, true, false>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      case 13: quad_rows_grad<true, true, false, true>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      case 12: quad_rows_grad<true, true, false, false>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      case 11: quad_rows_grad<true, false, true, true>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      case 10: quad_rows_grad<true, false, true, false>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      case 9: quad_rows_grad<true, false, false, true>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      case 8: quad_rows_grad<true, false, false, false>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      case 7: quad_rows_grad<false, true, true, true>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      case 6: quad_rows_grad<false, true, true, false>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      case 5: quad_rows_grad<false, true, false, true>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      case 4: quad_rows_grad<false, true, false, false>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      case 3: quad_rows_grad<false, false, true, true>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      case 2: quad_rows_grad<false, false, true, false>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      case 1: quad_rows_grad<false, false, false, true>(BRA, KET, NB, rowIdx, rowFlags, r0, cl, cf, qp, acc); break;
      default: break;
    }
  }
  // deterministic block sum of the 16 step values: warp shuffles, then one thread per value adds the warps in order
  __shared__ double sm[QGRAD_NS][QUAD_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < QGRAD_NS; ++k) {
    double v = acc[k * QUAD_THREADS];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) sm[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < QGRAD_NS) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < QUAD_THREADS / 32; ++w) v += sm[threadIdx.x][w];
    // value-major layout: the reduction kernel reads one contiguous run per step value
    partial[(int64_t)threadIdx.x * ((int64_t)gridDim.x * gridDim.y) + ((int64_t)by * gridDim.x + bx)] = v;
  }
}

// out1[k] = sum over CTAs of value k (k < n1), out2[k] = sum of value SQ_MAX_PROGRAM + k (k < n2); one block per value, fixed
// order of additions (deterministic)
__global__ void __launch_bounds__(256) quad_grad_reduce_kernel(const double* __restrict__ partial, int64_t nblocks, int n1,
                                                              double* __restrict__ out1, double* __restrict__ out2) {
  const int j = blockIdx.x, k = j < n1 ? j : SQ_MAX_PROGRAM + (j - n1);
  const double* p = partial + (int64_t)k * nblocks;
  double v = 0.0;
  for (int64_t b = threadIdx.x; b < nblocks; b += 256) v += p[b];
  __shared__ double sm[8];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sm[w];
    if (j < n1) out1[j] = t;
    else out2[j - n1] = t;
  }
}

// ---------------------------------------------------------------------------------------------
// host: work lists
// ---------------------------------------------------------------------------------------------
namespace {
struct Group { int idx[4]; int flags; int type; };

// groups of one spin: `codes1/2` are the PairTables codes of the two pairs, `rank` the mask -> index table
bool build_groups(const std::vector<uint32_t>& strs, const std::vector<int32_t>& rank, const std::vector<uint32_t>& code1,
                  const std::vector<uint32_t>& code2, int i1, int a1, int i2, int a2, int64_t lo, int64_t hi,
                  std::vector<Group>* out) {
  auto bit = [](uint32_t c, int b) -> int { return (int)((c >> b) & 1u); };
  const uint32_t mv1 = (1u << i1) | (1u << a1), mv2 = (1u << i2) | (1u << a2);
  for (int64_t I = lo; I < hi; ++I) {
    const uint32_t c1 = code1[I], c2 = code2[I];
    const uint32_t k1 = c1 & 3u, k2 = c2 & 3u;
    if (k1 == SQ_CLS_TGT || k2 == SQ_CLS_TGT) continue;   // covered by the group of its base string
    const bool act1 = k1 == SQ_CLS_SRC, act2 = k2 == SQ_CLS_SRC;
    const uint32_t m = strs[I];
    Group g;
    g.type = (act1 ? 2 : 0) + (act2 ? 1 : 0);
    for (int k = 0; k < 4; ++k) g.idx[k] = -1;
    const int64_t j00 = I;
    const int64_t j01 = act2 ? rank[m ^ mv2] : -1;
    const int64_t j10 = act1 ? rank[m ^ mv1] : -1;
    const int64_t j11 = (act1 && act2) ? rank[m ^ mv1 ^ mv2] : -1;
    g.idx[0] = (int)j00; g.idx[1] = (int)j01; g.idx[2] = (int)j10; g.idx[3] = (int)j11;
    for (int64_t j : {j01, j10, j11})
      if (j >= 0 && (j < lo || j >= hi)) return false;   // group leaves the shard
    // pair-1 flags: same-spin sign (src), cross flag of the a1=0 string, cross flag of the a1=1 string
    int f = 0;
    if (act1) {
      f |= bit(c1, 2) | (bit(c1, 3) << 1) | (bit(code1[j10], 3) << 2);
      if (act2) {   // the flags must not depend on the pair-2 index
        const uint32_t d = code1[j01];
        if (bit(d, 2) != bit(c1, 2) || bit(d, 3) != bit(c1, 3) || bit(code1[j11], 3) != bit(code1[j10], 3) ||
            bit(d, 4) != bit(c1, 4) || (d & 3u) != SQ_CLS_SRC)
          return false;
      }
    } else {
      f |= bit(c1, 3) << 1;
      if (act2 && bit(code1[j01], 3) != bit(c1, 3)) return false;
    }
    if (act2) {
      f |= (bit(c2, 2) | (bit(c2, 3) << 1) | (bit(code2[j01], 3) << 2)) << 4;
      if (act1) {
        const uint32_t d = code2[j10];
        if (bit(d, 2) != bit(c2, 2) || bit(d, 3) != bit(c2, 3) || bit(code2[j11], 3) != bit(code2[j01], 3) ||
            bit(d, 4) != bit(c2, 4) || (d & 3u) != SQ_CLS_SRC)
          return false;
      }
    } else {
      f |= (bit(c2, 3) << 1) << 4;
      if (act1 && bit(code2[j10], 3) != bit(c2, 3)) return false;
    }
    g.flags = f;
    out->push_back(g);
  }
  return true;
}

template <typename T>
int upload_vec(T** d, const std::vector<T>& v) {
  *d = nullptr;
  if (v.empty()) return SQ_OK;
  SQ_CUDA(cudaMalloc(d, sizeof(T) * v.size()));
  SQ_CUDA(cudaMemcpy(*d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return SQ_OK;
}
}  // namespace

int sq_build_quad_tables(sq_space* sp, const PairTables& p1, const PairTables& p2, int pair1, int pair2, QuadTables* qt) {
  qt->ok = false;
  qt->pair1 = pair1;
  qt->pair2 = pair2;
  if (p1.i == p2.i || p1.i == p2.a || p1.a == p2.i || p1.a == p2.a) return SQ_OK;   // overlapping pairs do not commute
  if (p1.sigma == 0 || p2.sigma == 0 || p1.cross_global || p2.cross_global || p1.n_cross_items || p2.n_cross_items) return SQ_OK;
  std::vector<Group> rows, cols;
  if (!build_groups(sp->strA, sp->rankA, p1.h_codeA, p2.h_codeA, p1.i, p1.a, p2.i, p2.a, sp->row_begin, sp->row_end, &rows))
    return SQ_OK;
  if (!build_groups(sp->strB, sp->rankB, p1.h_codeB, p2.h_codeB, p1.i, p1.a, p2.i, p2.a, 0, sp->NB, &cols)) return SQ_OK;
  std::vector<int4> colIdx, rowIdx;
  std::vector<int> colFlags, rowFlags;
  int64_t ncol[4] = {0, 0, 0, 0}, nrow[4] = {0, 0, 0, 0};
  int order[4] = {3, 2, 1, 0};
  for (int t = 0; t < 4; ++t) {
    for (const Group& g : cols)
      if (g.type == order[t]) {
        colIdx.push_back(make_int4(g.idx[0], g.idx[1], g.idx[2], g.idx[3]));
        colFlags.push_back(g.flags);
        ++ncol[order[t]];
      }
    while (colIdx.size() % QUAD_THREADS) {
      colIdx.push_back(make_int4(-1, -1, -1, -1));
      colFlags.push_back(0);
    }
    qt->colblk_end[t] = (int)(colIdx.size() / QUAD_THREADS);
    for (const Group& g : rows)
      if (g.type == order[t]) {
        auto loc = [&](int j) { return j < 0 ? -1 : (int)(j - sp->row_begin); };
        rowIdx.push_back(make_int4(loc(g.idx[0]), loc(g.idx[1]), loc(g.idx[2]), loc(g.idx[3])));
        rowFlags.push_back(g.flags);
        ++nrow[order[t]];
      }
    while (rowIdx.size() % QUAD_ROWS) {
      rowIdx.push_back(make_int4(-1, -1, -1, -1));
      rowFlags.push_back(0);
    }
    qt->rowchunk_end[t] = (int)(rowIdx.size() / QUAD_ROWS);
  }
  static const int gsize[4] = {1, 2, 2, 4};
  qt->touched = 0;
  for (int rt = 0; rt < 4; ++rt)
    for (int ct = 0; ct < 4; ++ct)
      if (rt || ct) qt->touched += nrow[rt] * gsize[rt] * ncol[ct] * gsize[ct];
  if (sp->device >= 0) {
    SQ_CUDA(cudaSetDevice(sp->device));
    SQ_CHECK(upload_vec(&qt->d_colIdx, colIdx));
    SQ_CHECK(upload_vec(&qt->d_colFlags, colFlags));
    SQ_CHECK(upload_vec(&qt->d_rowIdx, rowIdx));
    SQ_CHECK(upload_vec(&qt->d_rowFlags, rowFlags));
  }
  qt->ok = true;
  return SQ_OK;
}

void sq_free_quad_tables(QuadTables* qt) {
  cudaFree(qt->d_colIdx);
  cudaFree(qt->d_colFlags);
  cudaFree(qt->d_rowIdx);
  cudaFree(qt->d_rowFlags);
  qt->d_colIdx = nullptr; qt->d_rowIdx = nullptr; qt->d_colFlags = nullptr; qt->d_rowFlags = nullptr;
}

// gauge-fixed 4x4 brick matrix and total single rotations (same construction as build_tile_matrices)
static void brick_matrices(const TileStep* steps, int n_steps, int sigma, double* m16, double* ca, double* sa, double* cb,
                           double* sb) {
  double M[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  auto apply = [&](int u, int v, double c, double s) {
    for (int k = 0; k < 4; ++k) {
      const double a = M[u][k], b = M[v][k];
      M[u][k] = c * a - s * b;
      M[v][k] = c * b + s * a;
    }
  };
  double ac = 1, as = 0, bc = 1, bs = 0;
  for (int k = 0; k < n_steps; ++k) {
    const double c = steps[k].c, s = steps[k].s;
    if (steps[k].kind == 0) {
      apply(0, 2, c, s);
      apply(1, 3, c, s);
      const double nc = ac * c - as * s, ns = as * c + ac * s;
      ac = nc; as = ns;
    } else if (steps[k].kind == 1) {
      apply(0, 1, c, s);
      apply(2, 3, c, s);
      const double nc = bc * c - bs * s, ns = bs * c + bc * s;
      bc = nc; bs = ns;
    } else {
      apply(0, 3, c, sigma * s);
    }
  }
  for (int r = 0; r < 4; ++r)
    for (int k = 0; k < 4; ++k) m16[4 * r + k] = M[r][k];
  *ca = ac; *sa = as; *cb = bc; *sb = bs;
}

int sq_launch_quad(sq_space* sp, const QuadTables& qt, const TileStep* steps1, int n1, int sigma1, const TileStep* steps2,
                   int n2, int sigma2, double* state, cudaStream_t st) {
  if (!qt.ok) {
    sq_set_error("quad launch on a pair combination without quad tables");
    return SQ_ERR_INVALID;
  }
  QuadMats qm;
  brick_matrices(steps1, n1, sigma1, qm.m1, &qm.ca1, &qm.sa1, &qm.cb1, &qm.sb1);
  brick_matrices(steps2, n2, sigma2, qm.m2, &qm.ca2, &qm.sa2, &qm.cb2, &qm.sb2);
  QuadBounds qb = {qt.colblk_end[0], qt.colblk_end[1], qt.colblk_end[2], qt.colblk_end[3],
                   qt.rowchunk_end[0], qt.rowchunk_end[1], qt.rowchunk_end[2], qt.rowchunk_end[3]};
  if (qb.c0 == 0 || qb.r0 == 0) return SQ_OK;
  dim3 grid((unsigned)qb.c0, (unsigned)qb.r0);
#if QUAD_ASYNC
  const size_t smem = sizeof(double) * QUAD_STAGES * 16 * QUAD_THREADS;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
#else
  const size_t smem = 0;
#endif
  quad_kernel<<<grid, QUAD_THREADS, smem, st>>>(state, qt.d_colIdx, qt.d_colFlags, qt.d_rowIdx, qt.d_rowFlags, sp->NB, qb,
                                               qm);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sq_set_error("quad_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}

static void grad_steps(const TileStep* steps, int n, int sigma, GradSteps* gs) {
  gs->n = n;
  for (int k = 0; k < SQ_MAX_PROGRAM; ++k) {
    const bool on = k < n;
    const double sig = (on && steps[k].kind == 2) ? (double)sigma : 1.0;
    gs->kind[k] = on ? steps[k].kind : -1;
    gs->c[k] = on ? steps[k].c : 1.0;
    gs->s[k] = on ? sig * steps[k].s : 0.0;
    gs->sig[k] = sig;
  }
}

// d_out1[k] / d_out2[k] = <bra|T_k|ket> of step k of brick 1 / 2 (taken before that step), then both vectors are rotated
int sq_launch_quad_grad(sq_space* sp, const QuadTables& qt, const TileStep* steps1, int n1, int sigma1, const TileStep* steps2,
                        int n2, int sigma2, double* bra, double* ket, double* d_out1, double* d_out2, cudaStream_t st) {
  if (!qt.ok || n1 < 1 || n2 < 1 || n1 > SQ_MAX_PROGRAM || n2 > SQ_MAX_PROGRAM) {
    sq_set_error("quad gradient launch on a pair combination without quad tables or with a bad program");
    return SQ_ERR_INVALID;
  }
  QuadGradProg qp;
  grad_steps(steps1, n1, sigma1, &qp.g1);
  grad_steps(steps2, n2, sigma2, &qp.g2);
  QuadBounds qb = {qt.colblk_end[0], qt.colblk_end[1], qt.colblk_end[2], qt.colblk_end[3],
                   qt.rowchunk_end[0], qt.rowchunk_end[1], qt.rowchunk_end[2], qt.rowchunk_end[3]};
  if (qb.c0 == 0 || qb.r0 == 0) {
    SQ_CUDA(cudaMemsetAsync(d_out1, 0, sizeof(double) * n1, st));
    SQ_CUDA(cudaMemsetAsync(d_out2, 0, sizeof(double) * n2, st));
    return SQ_OK;
  }
  dim3 grid((unsigned)qb.c0, (unsigned)qb.r0);
  const int64_t nblocks = (int64_t)grid.x * grid.y;
  SQ_CHECK(sq_ensure_partial(sp, nblocks * QGRAD_NS + QGRAD_NS));
  static int minb = 0;
  if (!minb) {
    const char* e = getenv("SQ_QGRAD_MINB");   // resident CTAs per SM the kernel is compiled for: 2 (default, 128 registers), 1 or 3
    minb = (e && e[0] == '1') ? 1 : ((e && e[0] == '3') ? 3 : 2);
  }
  if (minb == 1)
    quad_grad_kernel<1><<<grid, QUAD_THREADS, 0, st>>>(bra, ket, qt.d_colIdx, qt.d_colFlags, qt.d_rowIdx, qt.d_rowFlags, sp->NB, qb, qp,
                                                      sp->d_partial);
  else if (minb == 3)
    quad_grad_kernel<3><<<grid, QUAD_THREADS, 0, st>>>(bra, ket, qt.d_colIdx, qt.d_colFlags, qt.d_rowIdx, qt.d_rowFlags, sp->NB, qb, qp,
                                                      sp->d_partial);
  else
    quad_grad_kernel<2><<<grid, QUAD_THREADS, 0, st>>>(bra, ket, qt.d_colIdx, qt.d_colFlags, qt.d_rowIdx, qt.d_rowFlags, sp->NB, qb, qp,
                                                      sp->d_partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sq_set_error("quad_grad_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  quad_grad_reduce_kernel<<<n1 + n2, 256, 0, st>>>(sp->d_partial, nblocks, n1, d_out1, d_out2);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    sq_set_error("quad_grad_reduce_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}
