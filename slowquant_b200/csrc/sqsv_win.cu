// "window" kernel: a whole group of bricks in ONE read + ONE write of the CI vector, through shared memory.
//
// Pick an orbital window [w0, w0+H) (the same for both spins).  A string is (prefix on orbitals < w0, window part,
// suffix on orbitals >= w0+H).  An operator whose orbitals all lie inside the window never changes prefix or
// suffix, conserves the electron count e_w of the window part, and its fermionic sign depends on window bits only.
// Hence the coefficient matrix C[Ia][Ib] decomposes into independent tiles
//     (alpha prefix, alpha suffix; beta prefix, beta suffix)  x  (C(H,e_wa) rows) x (C(H,e_wb) columns)
// and EVERY brick inside the window maps each tile onto itself.  In itertools.combinations order
//     I(prefix, w, suffix) = start(prefix) + off_{e_rem}(w) + rank(suffix),
// so the rows/columns of a tile sit at base + delta[class][j] (class = (electrons left after the prefix, e_w),
// j = rank of the window part), and consecutive suffixes of the same (prefix, e_w) are consecutive indices.
//
// A CTA works on a BATCH of G = 16 tiles with the same alpha rows and the same beta class:
//   * "run" windows (a beta suffix exists): the 16 tiles are 16 consecutive beta suffixes, i.e. every tile column is
//     a contiguous 128-byte run in memory;
//   * the "block" window (the window reaches the last orbital): the 16 tiles are 16 beta prefixes of one class, each
//     with contiguous columns.
// The batch index g is the fastest index in shared memory (tile[row][col][g], padded to 17), so the 16 lanes of a
// half-warp work on the SAME (row item, column item) of 16 tiles: every data access is bank-conflict free and every
// table look-up is uniform.  Brick work comes as ready-made lists per (orbital pair, e_wa, e_wb) built on the host:
// quad entries {r, r', c, c'} (4x4 update) and single entries (2x2 rotation), fetched one brick ahead.
//
// Sign-free gauge.  The reference orders spin orbitals a0 b0 a1 b1 ...; re-ordering them as (all alpha)(all beta)
// multiplies determinant |A,B> by D(A,B) = (-1)^{#{(p,q): p in A, q in B, q < p}}.  In that gauge every
// nearest-neighbour hop p <-> p+1 of either spin and the pair double carry a constant sign (checked on the host
// per pair, folded into the brick matrices), so a brick is ONE constant 4x4 matrix / 2x2 rotation for all tiles.
// The window-local part of D is applied when a tile is loaded and again when it is stored.
//
// The commutation-aware planner of sqsv_api.cu chooses windows and brick groups.
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <unordered_map>

#include "sqsv_internal.h"

#define WIN_THREADS 256
#define WIN_WARPS (WIN_THREADS / 32)
#define WIN_G 16                      // tiles per CTA
#define WIN_SLOTS (WIN_THREADS / (WIN_G / 2))   // a thread works on two tiles (128-bit shared-memory accesses)

struct WinDev {
  // alpha side
  const int2* groupsA;     // {first row (shard-relative), class}
  const int2* clsA;        // [ncls] {number of window strings, e_w}
  const int* deltaA;       // [ncls][LTA] offset of window string j from the group base
  // beta side
  const int2* chunksB;     // {chunk id, class | tiles in the batch << 16}
  const int* gbaseB;       // [chunk][WIN_G] first column of every tile of the batch
  const int2* rangesB;     // {first entry of rchunks, batches}: the batches of one CTA (all of one class)
  const int* rchunks;      // chunk ids ordered by class
  const int2* clsB;
  const int* deltaB;
  // brick work lists
  const uint32_t* lists;   // quad entries jr | jr' << 8 | jc << 16 | jc' << 24, then single entries jr0 | jc0 << 8 | jr1 << 16 | jc1 << 24
  const int4* listidx;     // [pair][e_wa][e_wb] {offset, n_quad, n_alpha_single, n_beta_single}
  int LTA, LTB, H1;        // H1 = H + 1
  int lanes_j;             // top window: the tiles of a batch are far apart, global accesses run along the columns
  int gp;                  // batch stride in shared memory: 16, or 18 for the top window (column-wise copies)
  int tile_doubles, maxQ, maxS;
};

struct WinProgram {
  int n;
  int pair[SQ_WIN_MAX_BRICKS];
  WinBrick br[SQ_WIN_MAX_BRICKS];
};

#ifndef WIN_CP_HINT
#define WIN_CP_HINT ""   // L2 prefetch hint of the tile copies (".L2::128B" / ".L2::256B": compile-time A/B, tools/ab_win_variants.sh)
#endif
__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const double* src) {
  asm volatile("cp.async.ca.shared.global" WIN_CP_HINT " [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void stg_stream(double* p, double v) {
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double lds64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ double2 lds128(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, double2 v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}

#define WIN_RMAX 16   // batches per CTA (one range of the beta chunks of one class)

// The vector is in the sign-free gauge (gauge_kernel) while window sweeps run, so the kernel is sign-free:
// load a batch, apply the bricks, store the batch.  A CTA owns one alpha group and a RANGE of up to WIN_RMAX beta
// batches of one class: class tables and work lists are staged once, and with NBUF = 2 the copies of the next batch
// are in flight while the bricks of the current one run (the top window, whose padded tile leaves room for one buffer
// only at three CTAs per SM, runs with NBUF = 1).
// Shared memory: [NBUF tiles: Rn x Wn x gp doubles][beta delta: LTB][alpha delta: LTA][batch bases: WIN_RMAX x 16]
// [tiles per batch: WIN_RMAX][list headers: int4 x SQ_WIN_MAX_BRICKS][per brick: maxQ uint2 quad entries + maxS uint32
// single entries, BYTE offsets in 16-bit fields].
template <int NBUF, bool BATCH>
__global__ void __launch_bounds__(WIN_THREADS, NBUF == 2 ? 2 : 3)
win_kernel(double* __restrict__ C0, int64_t NB, const WinDev W, const __grid_constant__ WinProgram P, int n_states, int64_t state_stride) {
  // n_states vectors C0 + s * state_stride share this CTA's tables and work lists (state-averaged batches, osa.py:1415-1864:
  // the class tables, batch bases and brick lists are staged ONCE, then every state streams through the same program)
  extern __shared__ double tile[];
  int* const sdB = reinterpret_cast<int*>(tile + NBUF * W.tile_doubles);
  int* const sdA = sdB + W.LTB;
  int* const sbase = sdA + W.LTA;                                        // [batch][16]
  int* const skcnt = sbase + WIN_RMAX * WIN_G;                           // [batch]
  int4* const shdr = reinterpret_cast<int4*>(skcnt + WIN_RMAX);          // 16-byte aligned (host pads the table sizes)
  uint2* const qall = reinterpret_cast<uint2*>(shdr + SQ_WIN_MAX_BRICKS);
  const int per_brick = 2 * W.maxQ + W.maxS;                             // words per brick: quads first, then singles

  const int2 ga = __ldg(W.groupsA + blockIdx.y);
  const int2 rg = __ldg(W.rangesB + blockIdx.x);                         // {first entry of rchunks, batches}
  const int n_items = rg.y;
  const int clsA = ga.y, clsB = __ldg(W.chunksB + __ldg(W.rchunks + rg.x)).y & 0xffff;
  const int2 ca2 = __ldg(W.clsA + clsA), cb2 = __ldg(W.clsB + clsB);
  const int Rn = ca2.x, Wn = cb2.x;
  const int GP = W.gp;
  const int RS = Wn * GP;                   // row stride (doubles)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = threadIdx.x & (WIN_G - 1);

  // ---- round trip 1: class tables, batch bases, list headers ----
  for (int t = threadIdx.x; t < Wn; t += WIN_THREADS) sdB[t] = __ldg(W.deltaB + clsB * W.LTB + t);
  for (int t = threadIdx.x; t < Rn; t += WIN_THREADS) sdA[t] = __ldg(W.deltaA + clsA * W.LTA + t);
  if ((int)threadIdx.x < n_items * WIN_G) {
    const int2 ch = __ldg(W.chunksB + __ldg(W.rchunks + rg.x + (threadIdx.x >> 4)));
    sbase[threadIdx.x] = __ldg(W.gbaseB + ch.x * WIN_G + g);
    if (g == 0) skcnt[threadIdx.x >> 4] = ch.y >> 16;
  }
  if (threadIdx.x >= 128 && (int)threadIdx.x < 128 + P.n)
    shdr[threadIdx.x - 128] = __ldg(W.listidx + (P.pair[threadIdx.x - 128] * W.H1 + ca2.y) * W.H1 + cb2.y);
  __syncthreads();

  const uint32_t tb0 = (uint32_t)__cvta_generic_to_shared(tile);
  // the whole batch is requested at once with 8-byte async copies
  auto issue_loads = [&](int q) {
    const int it = BATCH ? q % n_items : q;
    const double* C = BATCH ? C0 + (int64_t)(q / n_items) * state_stride : C0;
    const uint32_t tb = tb0 + (uint32_t)((q % NBUF) * W.tile_doubles) * 8u;
    const int kcnt = skcnt[it];
    const int* sb = sbase + it * WIN_G;
    if (W.lanes_j) {
      const int NX = kcnt * Wn;
      const float invW = 1.0f / (float)Wn;
      for (int r = warp; r < Rn; r += WIN_WARPS) {
        const double* src = C + (int64_t)(ga.x + sdA[r]) * NB;
        const uint32_t dst = tb + (uint32_t)(r * RS) * 8u;
        for (int x = lane; x < NX; x += 32) {
          const int gg = (int)(((float)x + 0.5f) * invW), j = x - gg * Wn;
          cp_async8(dst + (uint32_t)(j * GP + gg) * 8u, src + sb[gg] + sdB[j]);
        }
      }
    } else if (g < kcnt) {
      const int myb = sb[g];
      for (int r = warp; r < Rn; r += WIN_WARPS) {
        const double* src = C + (int64_t)(ga.x + sdA[r]) * NB + myb;
        const uint32_t dst = tb + (uint32_t)(r * RS + g) * 8u;
#pragma unroll 4
        for (int j = lane >> 4; j < Wn; j += 2) cp_async8(dst + (uint32_t)(j * GP) * 8u, src + sdB[j]);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // ---- round trip 2: the first batch, and (meanwhile) the list entries ----
  issue_loads(0);
  {
    // one entry per thread and brick (maxQ + maxS <= WIN_THREADS): all loads first, then scale to byte offsets
    uint32_t raw[SQ_WIN_MAX_BRICKS];
#pragma unroll
    for (int b = 0; b < SQ_WIN_MAX_BRICKS; ++b) {
      raw[b] = 0u;
      if (b < P.n) {
        const int4 li = shdr[b];
        if ((int)threadIdx.x < li.y + li.z + li.w) raw[b] = __ldg(W.lists + li.x + threadIdx.x);
      }
    }
#pragma unroll
    for (int b = 0; b < SQ_WIN_MAX_BRICKS; ++b) {
      if (b < P.n) {
        const int4 li = shdr[b];
        const int t = threadIdx.x;
        const uint32_t w = raw[b];
        uint32_t* dst = reinterpret_cast<uint32_t*>(qall) + b * per_brick;
        if (t < li.y) {
          const uint32_t r0 = (w & 255u) * RS, r1 = ((w >> 8) & 255u) * RS, c0 = ((w >> 16) & 255u) * GP, c1 = (w >> 24) * GP;
          reinterpret_cast<uint2*>(dst)[t] = make_uint2((r0 | (r1 << 16)) << 3, (c0 | (c1 << 16)) << 3);
        } else if (t < li.y + li.z + li.w) {
          const uint32_t o0 = (w & 255u) * RS + ((w >> 8) & 255u) * GP, o1 = ((w >> 16) & 255u) * RS + (w >> 24) * GP;
          dst[2 * W.maxQ + (t - li.y)] = (o0 | (o1 << 16)) << 3;
        }
      }
    }
  }

  // 8 lanes x 2 tiles per work-list entry (32 entries in flight per CTA); every shared-memory access moves 16 bytes
  const int g2 = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const uint32_t lb = (uint32_t)__cvta_generic_to_shared(qall);
  const int n_total = BATCH ? n_items * n_states : n_items;   // BATCH = false: one vector, the round-1 code path without index arithmetic
  for (int q = 0; q < n_total; ++q) {
    const int it = BATCH ? q % n_items : q;
    double* const C = BATCH ? C0 + (int64_t)(q / n_items) * state_stride : C0;
    const int kcnt = skcnt[it];
    const uint32_t tb = tb0 + (uint32_t)((q % NBUF) * W.tile_doubles) * 8u;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();   // batch `q` has landed; every thread is done with the stores of batch q - 1
    if (NBUF == 2 && q + 1 < n_total) issue_loads(q + 1);   // into the other buffer, in flight during the bricks

    // ---- bricks ----
    const bool on = 2 * g2 < kcnt;   // an odd batch computes one unused tile along (its lanes are never stored to memory)
    const uint32_t tgb = tb + (uint32_t)g2 * 16u;
    for (int b = 0; b < P.n; ++b) {
      if (b) __syncthreads();   // convergent: every thread of the CTA, also the lanes of unused tiles
      if (!on) continue;
      const int4 hd = shdr[b];
      const int nQ = hd.y, nSa = hd.z, nS = hd.z + hd.w;
      const uint32_t ql = lb + (uint32_t)(b * per_brick) * 4u, sl = ql + (uint32_t)W.maxQ * 8u;
      const WinBrick& br = P.br[b];
      // ONE loop over the brick's work: 4x4 entries first (padded to a multiple of 4 = the slots of a warp, so that a warp never
      // mixes the two entry kinds), then the 2x2 entries two at a time.  The slots left over by the last round of 4x4 entries
      // take 2x2 entries instead of idling (36 + 96 entries of a 20 x 20 tile: 3 rounds instead of 4).
      const int nQ4 = (nQ + 3) & ~3, nU = nQ4 + ((nS + 1) >> 1);
      for (int e = slot; e < nU; e += WIN_SLOTS) {
        if (e < nQ4) {
          if (e >= nQ) continue;
          uint32_t ux, uy;
          asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ux), "=r"(uy) : "r"(ql + (uint32_t)e * 8u));
          const uint32_t a0 = tgb + (ux & 0xffffu), a1 = tgb + (ux >> 16), c0 = uy & 0xffffu, c1 = uy >> 16;
          const double2 y0 = lds128(a0 + c0), y1 = lds128(a0 + c1), y2 = lds128(a1 + c0), y3 = lds128(a1 + c1);
          double2 z;
          z.x = br.m[0] * y0.x + br.m[1] * y1.x + br.m[2] * y2.x + br.m[3] * y3.x;
          z.y = br.m[0] * y0.y + br.m[1] * y1.y + br.m[2] * y2.y + br.m[3] * y3.y;
          sts128(a0 + c0, z);
          z.x = br.m[4] * y0.x + br.m[5] * y1.x + br.m[6] * y2.x + br.m[7] * y3.x;
          z.y = br.m[4] * y0.y + br.m[5] * y1.y + br.m[6] * y2.y + br.m[7] * y3.y;
          sts128(a0 + c1, z);
          z.x = br.m[8] * y0.x + br.m[9] * y1.x + br.m[10] * y2.x + br.m[11] * y3.x;
          z.y = br.m[8] * y0.y + br.m[9] * y1.y + br.m[10] * y2.y + br.m[11] * y3.y;
          sts128(a1 + c0, z);
          z.x = br.m[12] * y0.x + br.m[13] * y1.x + br.m[14] * y2.x + br.m[15] * y3.x;
          z.y = br.m[12] * y0.y + br.m[13] * y1.y + br.m[14] * y2.y + br.m[15] * y3.y;
          sts128(a1 + c1, z);
        } else {
          const int e0 = 2 * (e - nQ4), e1 = e0 + 1;   // alpha singles first, then beta singles
          uint32_t u, v;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u) : "r"(sl + (uint32_t)e0 * 4u));
          const uint32_t o0 = tgb + (u & 0xffffu), o1 = tgb + (u >> 16);
          const double2 y0 = lds128(o0), y1 = lds128(o1);
          const bool al = e0 < nSa;
          const double c = al ? br.ca : br.cb, sn = al ? br.sa : br.sb;
          if (e1 < nS) {
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sl + (uint32_t)e1 * 4u));
            const uint32_t p0 = tgb + (v & 0xffffu), p1 = tgb + (v >> 16);
            const double2 w0 = lds128(p0), w1 = lds128(p1);
            const bool bl = e1 < nSa;
            const double c2 = bl ? br.ca : br.cb, s2 = bl ? br.sa : br.sb;
            sts128(p0, make_double2(c2 * w0.x - s2 * w1.x, c2 * w0.y - s2 * w1.y));
            sts128(p1, make_double2(c2 * w1.x + s2 * w0.x, c2 * w1.y + s2 * w0.y));
          }
          sts128(o0, make_double2(c * y0.x - sn * y1.x, c * y0.y - sn * y1.y));
          sts128(o1, make_double2(c * y1.x + sn * y0.x, c * y1.y + sn * y0.y));
        }
      }
    }
    __syncthreads();

    // ---- store ----
    const int* sb = sbase + it * WIN_G;
    const double* tbuf = tile + (q % NBUF) * W.tile_doubles;
    if (W.lanes_j) {
      const int NX = kcnt * Wn;
      const float invW = 1.0f / (float)Wn;
      for (int r = warp; r < Rn; r += WIN_WARPS) {
        double* dst = C + (int64_t)(ga.x + sdA[r]) * NB;
        const double* srct = tbuf + r * RS;
        for (int x = lane; x < NX; x += 32) {
          const int gg = (int)(((float)x + 0.5f) * invW), j = x - gg * Wn;
          stg_stream(dst + sb[gg] + sdB[j], srct[j * GP + gg]);
        }
      }
    } else if (g < kcnt) {
      const int myb = sb[g];
      for (int r = warp; r < Rn; r += WIN_WARPS) {
        double* dst = C + (int64_t)(ga.x + sdA[r]) * NB + myb;
        const double* srct = tbuf + r * RS + g;
#pragma unroll 4
        for (int j = lane >> 4; j < Wn; j += 2) stg_stream(dst + sdB[j], srct[j * GP]);
      }
    }
    if (NBUF == 1 && q + 1 < n_total) {
      __syncthreads();   // the single buffer is free again
      issue_loads(q + 1);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Gradient sweep on windows: the same batches for TWO vectors (bra, ket).  For every rotation step k of every brick
// (reference ups_wavefunction.py:1114-1138) accumulate <bra|T_k|ket> on the current tiles, then rotate both vectors.
// A work-list entry is closed under all steps of its brick (alpha rotation: rows r,r'; beta rotation: columns c,c';
// pair double: (r,c) <-> (r',c')), so an entry stays in registers for the whole brick.  Reordering commuting bricks
// leaves every <bra|T_k|ket> unchanged (the moved operators commute with T_k and act on both vectors).
// ---------------------------------------------------------------------------------------------
struct WinGradBrick {
  int n;                        // rotation steps
  int kind[SQ_MAX_PROGRAM];     // 0 alpha rotation, 1 beta rotation, 2 pair double; bit 4: T has sign -1 in the window gauge
  double c[SQ_MAX_PROGRAM], s[SQ_MAX_PROGRAM];   // s carries the sign
  int slot0;                    // first output slot of this brick's steps
};
struct WinGradProgram {
  int n;
  int pair[SQ_WIN_MAX_BRICKS];
  WinGradBrick br[SQ_WIN_MAX_BRICKS];
};

__device__ __forceinline__ void grot(double& xs, double& xt, double c, double s) {
  const double a = xs, b = xt;
  xs = c * a - s * b;
  xt = c * b + s * a;
}

#define WING_REPL 8   // replicas of the output slots (CTAs add into replica blockIdx.x % 8: spreads the global atomics)

__device__ __forceinline__ void grot2(double2& xs, double2& xt, double c, double s) {
  const double2 a = xs, b = xt;
  xs = make_double2(c * a.x - s * b.x, c * a.y - s * b.y);
  xt = make_double2(c * b.x + s * a.x, c * b.y + s * a.y);
}
// <bra|T|ket> contribution of one (source, target) amplitude pair: bra_tgt * ket_src - bra_src * ket_tgt, both tiles of the thread
__device__ __forceinline__ double gdot2(const double2& bs, const double2& bt, const double2& ks, const double2& kt) {
  return (bt.x * ks.x - bs.x * kt.x) + (bt.y * ks.y - bs.y * kt.y);
}

// Shared memory: [bra tile][ket tile][per-warp accumulators: WIN_WARPS x SQ_WIN_MAX_BRICKS x SQ_MAX_PROGRAM doubles][tables as
// win_kernel].  Same thread layout as win_kernel: 8 lanes x 2 tiles per work-list entry, 128-bit shared-memory accesses.  The
// per-step values of a brick are reduced inside the warp with a packed butterfly (9 double shuffles for 8 values) and added to
// warp-private accumulators; one global atomic per (brick, step) and CTA at the end.
__global__ void __launch_bounds__(WIN_THREADS, 2)
win_grad_kernel(double* __restrict__ BRA, double* __restrict__ KET, int64_t NB, const WinDev W,
                const __grid_constant__ WinGradProgram P, double* __restrict__ grad_out, int n_out) {
  extern __shared__ double tile[];
  const int TD = W.tile_doubles;
  double* const sacc = tile + 2 * TD;
  int* const sdB = reinterpret_cast<int*>(sacc + WIN_WARPS * SQ_WIN_MAX_BRICKS * SQ_MAX_PROGRAM);
  int* const sdA = sdB + W.LTB;
  int* const sbase = sdA + W.LTA;
  int* const skcnt = sbase + WIN_RMAX * WIN_G;
  int4* const shdr = reinterpret_cast<int4*>(skcnt + WIN_RMAX);
  uint2* const qall = reinterpret_cast<uint2*>(shdr + SQ_WIN_MAX_BRICKS);
  const int per_brick = 2 * W.maxQ + W.maxS;

  const int2 ga = __ldg(W.groupsA + blockIdx.y);
  const int2 rg = __ldg(W.rangesB + blockIdx.x);
  const int n_items = rg.y;
  const int clsA = ga.y, clsB = __ldg(W.chunksB + __ldg(W.rchunks + rg.x)).y & 0xffff;
  const int2 ca2 = __ldg(W.clsA + clsA), cb2 = __ldg(W.clsB + clsB);
  const int Rn = ca2.x, Wn = cb2.x;
  const int GP = W.gp;
  const int RS = Wn * GP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = threadIdx.x & (WIN_G - 1);

  for (int t = threadIdx.x; t < Wn; t += WIN_THREADS) sdB[t] = __ldg(W.deltaB + clsB * W.LTB + t);
  for (int t = threadIdx.x; t < Rn; t += WIN_THREADS) sdA[t] = __ldg(W.deltaA + clsA * W.LTA + t);
  if ((int)threadIdx.x < n_items * WIN_G) {
    const int2 ch = __ldg(W.chunksB + __ldg(W.rchunks + rg.x + (threadIdx.x >> 4)));
    sbase[threadIdx.x] = __ldg(W.gbaseB + ch.x * WIN_G + g);
    if (g == 0) skcnt[threadIdx.x >> 4] = ch.y >> 16;
  }
  if (threadIdx.x >= 128 && (int)threadIdx.x < 128 + P.n)
    shdr[threadIdx.x - 128] = __ldg(W.listidx + (P.pair[threadIdx.x - 128] * W.H1 + ca2.y) * W.H1 + cb2.y);
  for (int t = threadIdx.x; t < WIN_WARPS * SQ_WIN_MAX_BRICKS * SQ_MAX_PROGRAM; t += WIN_THREADS) sacc[t] = 0.0;
  __syncthreads();

  const uint32_t tb0 = (uint32_t)__cvta_generic_to_shared(tile);
  auto issue_loads = [&](int it) {
    const int kcnt = skcnt[it];
    const int* sb = sbase + it * WIN_G;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const double* C = v ? KET : BRA;
      const uint32_t tb = tb0 + (uint32_t)(v * TD) * 8u;
      if (W.lanes_j) {
        const int NX = kcnt * Wn;
        const float invW = 1.0f / (float)Wn;
        for (int r = warp; r < Rn; r += WIN_WARPS) {
          const double* src = C + (int64_t)(ga.x + sdA[r]) * NB;
          const uint32_t dst = tb + (uint32_t)(r * RS) * 8u;
          for (int x = lane; x < NX; x += 32) {
            const int gg = (int)(((float)x + 0.5f) * invW), j = x - gg * Wn;
            cp_async8(dst + (uint32_t)(j * GP + gg) * 8u, src + sb[gg] + sdB[j]);
          }
        }
      } else if (g < kcnt) {
        const int myb = sb[g];
        for (int r = warp; r < Rn; r += WIN_WARPS) {
          const double* src = C + (int64_t)(ga.x + sdA[r]) * NB + myb;
          const uint32_t dst = tb + (uint32_t)(r * RS + g) * 8u;
#pragma unroll 4
          for (int j = lane >> 4; j < Wn; j += 2) cp_async8(dst + (uint32_t)(j * GP) * 8u, src + sdB[j]);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue_loads(0);
  {
    uint32_t raw[SQ_WIN_MAX_BRICKS];
#pragma unroll
    for (int b = 0; b < SQ_WIN_MAX_BRICKS; ++b) {
      raw[b] = 0u;
      if (b < P.n) {
        const int4 li = shdr[b];
        if ((int)threadIdx.x < li.y + li.z + li.w) raw[b] = __ldg(W.lists + li.x + threadIdx.x);
      }
    }
#pragma unroll
    for (int b = 0; b < SQ_WIN_MAX_BRICKS; ++b) {
      if (b < P.n) {
        const int4 li = shdr[b];
        const int t = threadIdx.x;
        const uint32_t w = raw[b];
        uint32_t* dst = reinterpret_cast<uint32_t*>(qall) + b * per_brick;
        if (t < li.y) {
          const uint32_t r0 = (w & 255u) * RS, r1 = ((w >> 8) & 255u) * RS, c0 = ((w >> 16) & 255u) * GP, c1 = (w >> 24) * GP;
          reinterpret_cast<uint2*>(dst)[t] = make_uint2((r0 | (r1 << 16)) << 3, (c0 | (c1 << 16)) << 3);
        } else if (t < li.y + li.z + li.w) {
          const uint32_t o0 = (w & 255u) * RS + ((w >> 8) & 255u) * GP, o1 = ((w >> 16) & 255u) * RS + (w >> 24) * GP;
          dst[2 * W.maxQ + (t - li.y)] = (o0 | (o1 << 16)) << 3;
        }
      }
    }
  }

  const int g2 = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const uint32_t lb = (uint32_t)__cvta_generic_to_shared(qall);
  const uint32_t kofs = (uint32_t)TD * 8u;   // ket tile behind the bra tile
  double* const wacc = sacc + warp * (SQ_WIN_MAX_BRICKS * SQ_MAX_PROGRAM);
  for (int it = 0; it < n_items; ++it) {
    const int kcnt = skcnt[it];
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const bool on = 2 * g2 < kcnt;
    const bool second = 2 * g2 + 1 < kcnt;        // the second tile of an odd batch holds stale data: it must not reach the sums
    const uint32_t tg = tb0 + (uint32_t)g2 * 16u;
    for (int b = 0; b < P.n; ++b) {
      if (b) __syncthreads();
      const WinGradBrick& br = P.br[b];
      double acc[SQ_MAX_PROGRAM];
#pragma unroll
      for (int k = 0; k < SQ_MAX_PROGRAM; ++k) acc[k] = 0.0;
      if (on) {
        const int4 hd = shdr[b];
        const int nQ = hd.y, nSa = hd.z, nS = hd.z + hd.w;
        const uint32_t ql = lb + (uint32_t)(b * per_brick) * 4u, sl = ql + (uint32_t)W.maxQ * 8u;
        for (int e = slot; e < nQ; e += WIN_SLOTS) {
          uint32_t ux, uy;
          asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ux), "=r"(uy) : "r"(ql + (uint32_t)e * 8u));
          const uint32_t a0 = tg + (ux & 0xffffu), a1 = tg + (ux >> 16), c0 = uy & 0xffffu, c1 = uy >> 16;
          double2 b00 = lds128(a0 + c0), b01 = lds128(a0 + c1), b10 = lds128(a1 + c0), b11 = lds128(a1 + c1);
          double2 k00 = lds128(a0 + c0 + kofs), k01 = lds128(a0 + c1 + kofs), k10 = lds128(a1 + c0 + kofs), k11 = lds128(a1 + c1 + kofs);
          if (!second) { b00.y = b01.y = b10.y = b11.y = 0.0; k00.y = k01.y = k10.y = k11.y = 0.0; }
#pragma unroll
          for (int k = 0; k < SQ_MAX_PROGRAM; ++k) {
            if (k >= br.n) break;
            const int kind = br.kind[k] & 3;
            const double c = br.c[k], sn = br.s[k];
            if (kind == 0) {
              acc[k] += gdot2(b00, b10, k00, k10) + gdot2(b01, b11, k01, k11);
              grot2(b00, b10, c, sn); grot2(b01, b11, c, sn); grot2(k00, k10, c, sn); grot2(k01, k11, c, sn);
            } else if (kind == 1) {
              acc[k] += gdot2(b00, b01, k00, k01) + gdot2(b10, b11, k10, k11);
              grot2(b00, b01, c, sn); grot2(b10, b11, c, sn); grot2(k00, k01, c, sn); grot2(k10, k11, c, sn);
            } else {
              acc[k] += gdot2(b00, b11, k00, k11);
              grot2(b00, b11, c, sn); grot2(k00, k11, c, sn);
            }
          }
          sts128(a0 + c0, b00); sts128(a0 + c1, b01); sts128(a1 + c0, b10); sts128(a1 + c1, b11);
          sts128(a0 + c0 + kofs, k00); sts128(a0 + c1 + kofs, k01); sts128(a1 + c0 + kofs, k10); sts128(a1 + c1 + kofs, k11);
        }
        for (int e = slot; e < nS; e += WIN_SLOTS) {
          uint32_t u;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u) : "r"(sl + (uint32_t)e * 4u));
          const uint32_t o0 = tg + (u & 0xffffu), o1 = tg + (u >> 16);
          const int want = e < nSa ? 0 : 1;   // alpha singles see the alpha rotations only, beta singles the beta ones
          double2 b0 = lds128(o0), b1 = lds128(o1), k0 = lds128(o0 + kofs), k1 = lds128(o1 + kofs);
          if (!second) { b0.y = b1.y = 0.0; k0.y = k1.y = 0.0; }
#pragma unroll
          for (int k = 0; k < SQ_MAX_PROGRAM; ++k) {
            if (k >= br.n) break;
            if ((br.kind[k] & 3) != want) continue;
            acc[k] += gdot2(b0, b1, k0, k1);
            grot2(b0, b1, br.c[k], br.s[k]);
            grot2(k0, k1, br.c[k], br.s[k]);
          }
          sts128(o0, b0); sts128(o1, b1); sts128(o0 + kofs, k0); sts128(o1 + kofs, k1);
        }
      }
      // packed butterfly over the warp: after the three exchange levels lane l holds value (l & 7) summed over 8 lanes, two more
      // levels finish the sum; lanes 0..7 add the 8 step values to the warp's accumulators (all lanes take part in the shuffles)
      {
        double v4[4], v2[2], v1;
        const bool h4 = lane & 4, h2 = lane & 2, h1 = lane & 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double keep = h4 ? acc[i + 4] : acc[i], send = h4 ? acc[i] : acc[i + 4];
          v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double keep = h2 ? v4[i + 2] : v4[i], send = h2 ? v4[i] : v4[i + 2];
          v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        {
          const double keep = h1 ? v2[1] : v2[0], send = h1 ? v2[0] : v2[1];
          v1 = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
        v1 += __shfl_xor_sync(0xffffffffu, v1, 8);
        v1 += __shfl_xor_sync(0xffffffffu, v1, 16);
        // lane l (< 8) now holds step index  4 * bit2(l) + 2 * bit1(l) + bit0(l) = l
        if (lane < 8 && lane < br.n) wacc[b * SQ_MAX_PROGRAM + lane] += v1;
      }
    }
    __syncthreads();
    // ---- store both vectors ----
    const int* sb = sbase + it * WIN_G;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      double* C = v ? KET : BRA;
      const double* tbuf = tile + v * TD;
      if (W.lanes_j) {
        const int NX = kcnt * Wn;
        const float invW = 1.0f / (float)Wn;
        for (int r = warp; r < Rn; r += WIN_WARPS) {
          double* dst = C + (int64_t)(ga.x + sdA[r]) * NB;
          const double* srct = tbuf + r * RS;
          for (int x = lane; x < NX; x += 32) {
            const int gg = (int)(((float)x + 0.5f) * invW), j = x - gg * Wn;
            stg_stream(dst + sb[gg] + sdB[j], srct[j * GP + gg]);
          }
        }
      } else if (g < kcnt) {
        const int myb = sb[g];
        for (int r = warp; r < Rn; r += WIN_WARPS) {
          double* dst = C + (int64_t)(ga.x + sdA[r]) * NB + myb;
          const double* srct = tbuf + r * RS + g;
#pragma unroll 4
          for (int j = lane >> 4; j < Wn; j += 2) stg_stream(dst + sdB[j], srct[j * GP]);
        }
      }
    }
    if (it + 1 < n_items) {
      __syncthreads();
      issue_loads(it + 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < SQ_WIN_MAX_BRICKS * SQ_MAX_PROGRAM) {
    const int b = threadIdx.x / SQ_MAX_PROGRAM, k = threadIdx.x - b * SQ_MAX_PROGRAM;
    if (b < P.n && k < P.br[b].n) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < WIN_WARPS; ++w) v += sacc[w * (SQ_WIN_MAX_BRICKS * SQ_MAX_PROGRAM) + threadIdx.x];
      if (v != 0.0) atomicAdd(grad_out + (size_t)(blockIdx.x % WING_REPL) * n_out + P.br[b].slot0 + k, (P.br[b].kind[k] & 16) ? -v : v);
    }
  }
}

// x[Ia][Ib] *= D(A,B) = (-1)^{popc(A & gword(B))}: into / out of the sign-free gauge (its own inverse)
__global__ void __launch_bounds__(256)
gauge_kernel(double* __restrict__ C, int64_t NB, int64_t n_rows, int64_t row_begin, const uint32_t* __restrict__ strA,
             const uint32_t* __restrict__ gwordB, int64_t state_stride) {
  const int64_t ib = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (ib >= NB) return;
  C += (int64_t)blockIdx.z * state_stride;   // one vector of a batch per grid plane
  const uint32_t gw = __ldg(gwordB + ib);
  const int64_t r0 = (int64_t)blockIdx.y * 32, r1 = min(r0 + 32, n_rows);
  for (int64_t r = r0; r < r1; ++r) {
    const uint32_t a = __ldg(strA + row_begin + r);
    if (__popc(a & gw) & 1) {
      double* p = C + r * NB + ib;
      *p = -*p;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host tables
// ---------------------------------------------------------------------------------------------
static inline int wneg(uint32_t m, uint32_t par) { return __builtin_popcount(m & par) & 1; }

template <typename T>
static int win_upload(T** d, const std::vector<T>& v) {
  *d = nullptr;
  if (v.empty()) return SQ_OK;
  SQ_CUDA(cudaMalloc(d, sizeof(T) * v.size()));
  SQ_CUDA(cudaMemcpy(*d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return SQ_OK;
}

// gauge words (global orbital positions): alpha = occupation mask of the window part; beta = mask whose bit p is
// the parity of the beta window electrons on orbitals < p.  D(A,B) = parity(popc(alpha word & beta word)).
static inline uint32_t gauge_beta_word(uint32_t mB) {
  uint32_t w = 0, par = 0;
  for (int p = 0; p < 32; ++p) {
    if (par) w |= 1u << p;
    if (mB & (1u << p)) par ^= 1u;
  }
  return w;
}
static inline int gauge_bit(uint32_t mA, uint32_t mB) { return __builtin_popcount(mA & gauge_beta_word(mB)) & 1; }


// largest tile dimension of a window of H orbitals starting at w0 (n orbitals, ne electrons of this spin)
int sq_win_max_class(int n, int ne, int w0, int H) {
  int best = 0;
  for (int e = 0; e <= H && e <= ne; ++e) {
    if (ne - e > n - H) continue;   // the other orbitals cannot hold the remaining electrons
    double v = 1;
    for (int i = 1; i <= e; ++i) v = v * (H - e + i) / i;
    best = std::max(best, (int)(v + 0.5));
  }
  (void)w0;
  return best;
}


// Tables of one spin.  Returns false (no error) when the window cannot be used with this space / partition.
static bool build_side(const sq_space* sp, int spin, int w0, int H, SideHost* out) {
  const std::vector<uint32_t>& strs = spin ? sp->strB : sp->strA;
  const int ne = spin ? sp->n_beta : sp->n_alpha;
  const int64_t lo = spin ? 0 : sp->row_begin, hi = spin ? sp->NB : sp->row_end;
  if (H < 1 || H > 8 || w0 < 0 || w0 + H > sp->n_orb) return false;
  const bool block = (w0 + H == sp->n_orb);   // no suffix: batches are made of prefixes
  const uint32_t wmask = (1u << H) - 1u, wbits = wmask << w0, premask = (1u << w0) - 1u;
  std::vector<std::vector<uint32_t>>& wl = out->wl;
  wl.assign(H + 1, {});
  for (uint32_t w = 0; w <= wmask; ++w) wl[__builtin_popcount(w)].push_back(w);
  std::vector<int> pos(wmask + 1, 0);
  int LT = 1;
  for (int e = 0; e <= H; ++e) {
    std::sort(wl[e].begin(), wl[e].end(), [](uint32_t a, uint32_t b) {
      const uint32_t d = a ^ b;
      if (!d) return false;
      return (a & (d & (0u - d))) != 0;   // a lower orbital occupied sorts first
    });
    for (size_t j = 0; j < wl[e].size(); ++j) pos[wl[e][j]] = (int)j;
    LT = std::max(LT, (int)wl[e].size());
  }
  if (LT > 255) return false;   // list entries hold 8-bit string ranks
  // pass 1: groups = strings sharing prefix and suffix
  struct G {
    int64_t base;
    int cls, e_w, seen;
    uint32_t pre;
  };
  std::vector<G> gs;
  std::unordered_map<uint32_t, int> gid;
  std::map<std::pair<int, int>, int> cls_of;
  std::vector<int> cls_ew;
  for (int64_t I = lo; I < hi; ++I) {
    const uint32_t m = strs[I], key = m & ~wbits;
    auto it = gid.find(key);
    if (it == gid.end()) {
      const int e_w = __builtin_popcount(m & wbits), e_rem = ne - __builtin_popcount(m & premask);
      auto ck = std::make_pair(e_rem, e_w);
      auto ci = cls_of.find(ck);
      if (ci == cls_of.end()) {
        ci = cls_of.emplace(ck, (int)cls_ew.size()).first;
        cls_ew.push_back(e_w);
      }
      it = gid.emplace(key, (int)gs.size()).first;
      gs.push_back({-1, ci->second, e_w, 0, m & premask});
    }
    G& g = gs[it->second];
    if (pos[(m & wbits) >> w0] == 0) g.base = I;
    ++g.seen;
  }
  const int ncls = (int)cls_ew.size();
  if (ncls > 0xffff) return false;
  for (const G& g : gs)
    if (g.base < 0 || g.seen != (int)wl[g.e_w].size()) return false;   // group straddles the shard boundary
  // pass 2: offsets of the window strings from the group base are a function of the class only
  std::vector<int> delta((size_t)ncls * LT, INT_MIN);
  for (int64_t I = lo; I < hi; ++I) {
    const uint32_t m = strs[I];
    const G& g = gs[gid[m & ~wbits]];
    const int j = pos[(m & wbits) >> w0];
    const int64_t d = I - g.base;
    if (d < 0 || d > INT_MAX) return false;
    int& slot = delta[(size_t)g.cls * LT + j];
    if (slot == INT_MIN) slot = (int)d;
    else if (slot != (int)d) return false;
  }
  for (int& d : delta)
    if (d == INT_MIN) d = 0;
  out->cls.resize(ncls);
  out->max_cnt = 0;
  for (int c = 0; c < ncls; ++c) {
    out->cls[c] = make_int2((int)wl[cls_ew[c]].size(), cls_ew[c]);
    out->max_cnt = std::max(out->max_cnt, out->cls[c].x);
  }
  std::vector<int> order(gs.size());
  for (size_t i = 0; i < gs.size(); ++i) order[i] = (int)i;
  std::vector<int2> groups;
  if (spin == 0) {
    std::sort(order.begin(), order.end(), [&](int a, int b) { return gs[a].base < gs[b].base; });
    for (int i : order) groups.push_back(make_int2((int)(gs[i].base - sp->row_begin), gs[i].cls));
    std::stable_sort(groups.begin(), groups.end(), [&](const int2& a, const int2& b) { return out->cls[a.y].x > out->cls[b.y].x; });
  } else {
    // batches of up to WIN_G tiles of one class, in memory order (consecutive suffixes of one prefix are
    // consecutive columns, so most lanes of a batch read neighbouring addresses)
    (void)block;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
      return gs[a].cls != gs[b].cls ? gs[a].cls < gs[b].cls : gs[a].base < gs[b].base;
    });
    size_t i = 0;
    while (i < order.size()) {
      const G& g0 = gs[order[i]];
      int c = 1;
      while (c < WIN_G && i + c < order.size()) {
        const G& g1 = gs[order[i + c]];
        if (g1.cls != g0.cls) break;
        ++c;
      }
      const int chunk = (int)(out->gbase.size() / WIN_G);
      for (int t = 0; t < WIN_G; ++t) out->gbase.push_back(t < c ? (int)gs[order[i + t]].base : 0);
      groups.push_back(make_int2(chunk, g0.cls | (c << 16)));
      i += c;
    }
    std::stable_sort(groups.begin(), groups.end(), [&](const int2& a, const int2& b) {
      return out->cls[a.y & 0xffff].x * (a.y >> 16) > out->cls[b.y & 0xffff].x * (b.y >> 16);
    });
  }
  out->groups.swap(groups);
  out->delta.swap(delta);
  out->ncls = ncls;
  out->LT = LT;
  return true;
}

void sq_free_win_tables(WinTables* wt) {
  if (!wt) return;
  cudaFree(wt->d_groupsA);
  cudaFree(wt->d_clsA);
  cudaFree(wt->d_deltaA);
  cudaFree(wt->d_chunksB);
  cudaFree(wt->d_gbaseB);
  cudaFree(wt->d_rangesB);
  cudaFree(wt->d_rchunks);
  cudaFree(wt->d_clsB);
  cudaFree(wt->d_deltaB);
  cudaFree(wt->d_lists);
  cudaFree(wt->d_listidx);
  sq_free_win3(wt->w3);
  delete wt;
}

// pairs of the layout that the window kernel can take: nearest-neighbour orbitals inside the window (their signs
// vanish in the window gauge), no row pair spanning two devices
bool sq_win_pair_ok(const sq_layout* lay, int pair, int w0, int H) {
  const PairTables& pt = lay->pairs[pair];
  const int lo = std::min(pt.i, pt.a), hi = std::max(pt.i, pt.a);
  return hi == lo + 1 && lo >= w0 && hi < w0 + H && !pt.blocked && !pt.cross_global && pt.n_cross_items == 0;
}

size_t sq_win_smem_bytes(int max_a, int max_b, int gp, int nbuf, int lta, int ltb, int maxQ, int maxS, int n_bricks) {
  const size_t tile = sizeof(double) * (size_t)max_a * (size_t)max_b * (size_t)gp;   // even number of doubles
  const size_t tabs = 4 * (size_t)(ltb + lta + WIN_RMAX * WIN_G + WIN_RMAX + 4 * SQ_WIN_MAX_BRICKS);
  return (size_t)nbuf * tile + tabs + (size_t)n_bricks * ((size_t)maxQ * 8 + (size_t)maxS * 4);
}

// Signs of Ta, Tb and the pair double of orbital pair (i,a) in the window gauge, checked over every pair of
// window parts; false if one of them is not constant (the pair then stays with the tile kernels).
static bool gauge_signs(const sq_space* sp, const PairTables& pt, int a0, int Ha, int b0, int Hb, int8_t* eps3) {
  const int i = pt.i, a = pt.a;
  StringAction actA, actB, actD;
  int32_t la[2] = {2 * (2 * a) + 1, 2 * (2 * i)};
  int32_t lb[2] = {2 * (2 * a + 1) + 1, 2 * (2 * i + 1)};
  int32_t ld[4] = {2 * (2 * a + 1) + 1, 2 * (2 * a) + 1, 2 * (2 * i + 1), 2 * (2 * i)};
  if (sq_make_string_action(sp, la, 2, &actA) != SQ_OK || sq_make_string_action(sp, lb, 2, &actB) != SQ_OK ||
      sq_make_string_action(sp, ld, 4, &actD) != SQ_OK)
    return false;
  const uint32_t maskA = ((1u << Ha) - 1u) << a0, maskB = ((1u << Hb) - 1u) << b0;
  // every sign factor must be a function of window bits only
  if ((actA.parA & ~maskA) || (actA.parB & ~maskB) || (actB.parA & ~maskA) || (actB.parB & ~maskB) || (actD.parA & ~maskA) ||
      (actD.parB & ~maskB))
    return false;
  const uint32_t bi = 1u << i, ba = 1u << a;
  int seen[3] = {-1, -1, -1};
  auto note = [&](int kind, int bit) {
    if (seen[kind] < 0) seen[kind] = bit;
    return seen[kind] == bit;
  };
  std::vector<uint32_t> gw(1u << Hb);   // gauge words of all beta window parts
  for (uint32_t wb = 0; wb < (1u << Hb); ++wb) gw[wb] = gauge_beta_word(wb << b0);
  auto gbit = [&](uint32_t mA, uint32_t wb) { return __builtin_popcount(mA & gw[wb]) & 1; };
  const uint32_t fb = (bi | ba) >> b0;   // beta partner = wb ^ fb
  for (uint32_t wa = 0; wa < (1u << Ha); ++wa) {
    const uint32_t mA = wa << a0;
    const bool srcA = (mA & bi) && !(mA & ba);
    const uint32_t mAp = mA ^ bi ^ ba;
    for (uint32_t wb = 0; wb < (1u << Hb); ++wb) {
      const uint32_t mB = wb << b0;
      const bool srcB = (mB & bi) && !(mB & ba);
      if (srcA) {   // alpha single: sign = s0 * (-1)^{popc(A & parA) + popc(B & parB)}
        const int bit = ((actA.s0 < 0) ? 1 : 0) ^ wneg(mA, actA.parA) ^ wneg(mB, actA.parB) ^ gbit(mA, wb) ^ gbit(mAp, wb);
        if (!note(0, bit)) return false;
      }
      if (srcB) {
        const int bit = ((actB.s0 < 0) ? 1 : 0) ^ wneg(mA, actB.parA) ^ wneg(mB, actB.parB) ^ gbit(mA, wb) ^ gbit(mA, wb ^ fb);
        if (!note(1, bit)) return false;
      }
      if (srcA && srcB) {   // pair double: the normal-ordered string carries -1 (build_pair_tables)
        const int bit = ((-actD.s0 < 0) ? 1 : 0) ^ wneg(mA, actD.parA) ^ wneg(mB, actD.parB) ^ gbit(mA, wb) ^ gbit(mAp, wb ^ fb);
        if (!note(2, bit)) return false;
      }
    }
  }
  for (int k = 0; k < 3; ++k) eps3[k] = (seen[k] == 1) ? -1 : 1;
  return true;
}


// Tables for the window [w0, w0+H); cached on the layout.  (*out)->ok is false when the window is unusable.
int sq_get_win(sq_space* sp, sq_layout* lay, int w0, int H, const WinTables** out) {
  const std::array<int, 5> key = {w0, H, 0, 0, 0};
  auto it = lay->wins.find(key);
  if (it != lay->wins.end()) {
    *out = it->second;
    return SQ_OK;
  }
  WinTables* wt = new WinTables();
  lay->wins[key] = wt;
  *out = wt;
  wt->w0 = w0;
  wt->H = H;
  if (sp->world > 1) {
    int k = 0;
    while ((1 << k) < sp->world) ++k;
    if (w0 < k) return SQ_OK;   // alpha prefixes would straddle devices
  }
  wt->pair_local.assign(lay->pairs.size(), -1);
  std::vector<int> pairs;
  for (size_t p = 0; p < lay->pairs.size(); ++p)
    if (sq_win_pair_ok(lay, (int)p, w0, H)) {
      wt->pair_local[p] = (int)pairs.size();
      pairs.push_back((int)p);
    }
  if (pairs.empty()) return SQ_OK;
  // constant signs of the three generators of every pair in the window gauge
  wt->eps.assign(3 * pairs.size(), 1);
  for (int p : pairs) {
    wt->pair_lo.push_back(std::min(lay->pairs[p].i, lay->pairs[p].a) - w0);
    wt->pair_flip.push_back(lay->pairs[p].i > lay->pairs[p].a ? 1 : 0);
  }
  for (size_t lp = 0; lp < pairs.size(); ++lp)
    if (!gauge_signs(sp, lay->pairs[pairs[lp]], w0, H, w0, H, &wt->eps[3 * lp])) return SQ_OK;
  SideHost hA, hB;
  if (!build_side(sp, 0, w0, H, &hA)) return SQ_OK;
  if (!build_side(sp, 1, w0, H, &hB)) return SQ_OK;
  if (hA.groups.size() > 65535 || hB.groups.empty() || hA.groups.empty()) return SQ_OK;
  // ranges of up to WIN_RMAX batches of one class per CTA; heavy ranges first
  std::vector<int> rchunks;
  std::vector<int2> ranges;
  {
    std::vector<int> order(hB.groups.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return (hB.groups[a].y & 0xffff) < (hB.groups[b].y & 0xffff); });
    size_t i = 0;
    while (i < order.size()) {
      const int cls = hB.groups[order[i]].y & 0xffff;
      size_t j = i;
      while (j < order.size() && (hB.groups[order[j]].y & 0xffff) == cls) ++j;
      static int rmax = 0;
      if (!rmax) {
        const char* e = getenv("SQ_WIN_RANGE");   // batches per CTA, 1..WIN_RMAX
        rmax = e ? std::max(1, std::min(atoi(e), WIN_RMAX)) : 8;   // 8 measured as good as 16, 4 is 2 % slower, 1 is 16 % slower
      }
      const int n = (int)(j - i), parts = (n + rmax - 1) / rmax;
      for (int q = 0; q < parts; ++q) {
        const int b0 = (int)((int64_t)n * q / parts), b1 = (int)((int64_t)n * (q + 1) / parts);
        ranges.push_back(make_int2((int)rchunks.size(), b1 - b0));
        for (int t = b0; t < b1; ++t) rchunks.push_back(order[i + t]);
      }
      i = j;
    }
    std::stable_sort(ranges.begin(), ranges.end(), [&](const int2& a, const int2& b) {
      auto weight = [&](const int2& r) {
        int w = 0;
        for (int t = 0; t < r.y; ++t) w += hB.groups[rchunks[r.x + t]].y >> 16;
        return w * hB.cls[hB.groups[rchunks[r.x]].y & 0xffff].x;
      };
      return weight(a) > weight(b);
    });
  }
  // work lists per (pair, e_wa, e_wb)
  const int H1 = H + 1;
  std::vector<uint32_t> lists;
  std::vector<int4> listidx(pairs.size() * (size_t)H1 * H1, make_int4(0, 0, 0, 0));
  int maxQ = 1, maxS = 1;
  for (size_t lp = 0; lp < pairs.size(); ++lp) {
    const PairTables& pt = lay->pairs[pairs[lp]];
    const uint32_t bi = 1u << (pt.i - w0), ba = 1u << (pt.a - w0);
    // per electron count: src strings (j, j') and inert strings of this pair
    std::vector<std::vector<std::pair<int, int>>> src(H1);
    std::vector<std::vector<int>> inert(H1);
    for (int e = 0; e <= H; ++e) {
      const std::vector<uint32_t>& w = hA.wl[e];
      std::unordered_map<uint32_t, int> pos;
      for (size_t j = 0; j < w.size(); ++j) pos[w[j]] = (int)j;
      for (size_t j = 0; j < w.size(); ++j) {
        const uint32_t m = w[j];
        if ((m & bi) && !(m & ba)) src[e].push_back({(int)j, pos[m ^ bi ^ ba]});
        else if (((m & bi) != 0) == ((m & ba) != 0)) inert[e].push_back((int)j);
      }
    }
    for (int ea = 0; ea <= H; ++ea)
      for (int eb = 0; eb <= H; ++eb) {
        int4 idx = make_int4((int)lists.size(), 0, 0, 0);
        for (auto& r : src[ea])
          for (auto& c : src[eb]) {
            lists.push_back((uint32_t)r.first | ((uint32_t)r.second << 8) | ((uint32_t)c.first << 16) | ((uint32_t)c.second << 24));
            ++idx.y;
          }
        for (auto& r : src[ea])       // alpha single: (r, c) <-> (r', c), c inert
          for (int c : inert[eb]) {
            lists.push_back((uint32_t)r.first | ((uint32_t)c << 8) | ((uint32_t)r.second << 16) | ((uint32_t)c << 24));
            ++idx.z;
          }
        for (int r : inert[ea])       // beta single: (r, c) <-> (r, c'), r inert
          for (auto& c : src[eb]) {
            lists.push_back((uint32_t)r | ((uint32_t)c.first << 8) | ((uint32_t)r << 16) | ((uint32_t)c.second << 24));
            ++idx.w;
          }
        listidx[(lp * H1 + ea) * H1 + eb] = idx;
        maxQ = std::max(maxQ, idx.y);
        maxS = std::max(maxS, idx.z + idx.w);
      }
  }
  wt->lanes_j = (w0 + H == sp->n_orb) ? 1 : 0;
  wt->gp = wt->lanes_j ? 18 : 16;   // even: 16-byte aligned tile pairs; 18 spreads the column-wise copies over the banks
  if ((size_t)hA.max_cnt * hB.max_cnt * wt->gp * 8 > 65535) return SQ_OK;   // 16-bit byte offsets inside a batch
  if (maxQ + maxS > WIN_THREADS) return SQ_OK;   // the kernel fetches one list entry per thread and brick
  maxS = (maxS + 1) & ~1;   // every brick's list area stays 8-byte aligned
  wt->maxQ = maxQ;
  wt->maxS = maxS;
  wt->tile_doubles = hA.max_cnt * hB.max_cnt * wt->gp;   // even
  wt->LTA = (hA.LT + 3) & ~3;   // multiples of 4 words: the headers behind the tables stay 16-byte aligned
  wt->LTB = (hB.LT + 3) & ~3;
  {
    // one tile buffer and three CTAs per SM measured faster than two buffers (next batch in flight during the bricks)
    // and two CTAs per SM; SQ_WIN_NBUF=2 selects the latter for experiments
    const char* nb = getenv("SQ_WIN_NBUF");
    wt->nbuf = (nb && nb[0] == '2' && !wt->lanes_j) ? 2 : 1;
  }
  if (sq_win_smem_bytes(hA.max_cnt, hB.max_cnt, wt->gp, wt->nbuf, wt->LTA, wt->LTB, maxQ, maxS, SQ_WIN_MAX_BRICKS) > 220 * 1024) return SQ_OK;
  // the device tables use the padded leading dimensions
  auto repad = [](std::vector<int>& v, int ncls, int lt_old, int lt_new) {
    std::vector<int> o((size_t)ncls * lt_new, 0);
    for (int c = 0; c < ncls; ++c)
      for (int j = 0; j < lt_old; ++j) o[(size_t)c * lt_new + j] = v[(size_t)c * lt_old + j];
    v.swap(o);
  };
  repad(hA.delta, hA.ncls, hA.LT, wt->LTA);
  repad(hB.delta, hB.ncls, hB.LT, wt->LTB);
  wt->smem = sq_win_smem_bytes(hA.max_cnt, hB.max_cnt, wt->gp, wt->nbuf, wt->LTA, wt->LTB, maxQ, maxS, SQ_WIN_MAX_BRICKS);
  wt->max_a = hA.max_cnt;
  wt->max_b = hB.max_cnt;
  wt->n_groups_a = (int)hA.groups.size();
  wt->n_chunks_b = (int)hB.groups.size();
  wt->n_ranges_b = (int)ranges.size();
  wt->touched = sp->local_len();
  SQ_CHECK(sq_build_win3(sp, wt, hA, hB));   // tables of win3_kernel (register blocks over orbital triples); wt->w3->ok says whether usable
  if (sp->device >= 0) {
    SQ_CUDA(cudaSetDevice(sp->device));
    SQ_CHECK(win_upload(&wt->d_groupsA, hA.groups));
    SQ_CHECK(win_upload(&wt->d_clsA, hA.cls));
    SQ_CHECK(win_upload(&wt->d_deltaA, hA.delta));
    SQ_CHECK(win_upload(&wt->d_chunksB, hB.groups));
    SQ_CHECK(win_upload(&wt->d_gbaseB, hB.gbase));
    SQ_CHECK(win_upload(&wt->d_rangesB, ranges));
    SQ_CHECK(win_upload(&wt->d_rchunks, rchunks));
    SQ_CHECK(win_upload(&wt->d_clsB, hB.cls));
    SQ_CHECK(win_upload(&wt->d_deltaB, hB.delta));
    SQ_CHECK(win_upload(&wt->d_lists, lists));
    SQ_CHECK(win_upload(&wt->d_listidx, listidx));
  }
  wt->ok = true;
  return SQ_OK;
}

// bricks[k] = (layout pair index, rotation steps of the fused program); every pair must be usable in `wt`
int sq_launch_win(sq_space* sp, const WinTables& wt, const int* pair_idx, const TileStep* const* steps, const int* n_steps,
                  int n_bricks, double* state, cudaStream_t st, int n_states, int64_t state_stride) {
  if (!wt.ok || n_bricks < 1 || n_bricks > SQ_WIN_MAX_BRICKS) {
    sq_set_error("window launch with %d bricks (max %d) or without tables", n_bricks, SQ_WIN_MAX_BRICKS);
    return SQ_ERR_INVALID;
  }
  if (wt.w3 && wt.w3->ok && sq_win3_enabled()) {
    Win3Program P3;
    SQ_CHECK(sq_win3_program(wt, pair_idx, steps, n_steps, n_bricks, &P3));
    return sq_launch_win3(sp, wt, P3, state, st, n_states, state_stride);
  }
  WinProgram P;
  P.n = n_bricks;
  for (int k = 0; k < n_bricks; ++k) {
    const int lp = wt.pair_local[pair_idx[k]];
    if (lp < 0) {
      sq_set_error("window launch: orbital pair %d is outside the window", pair_idx[k]);
      return SQ_ERR_INVALID;
    }
    P.pair[k] = lp;
    TileMatrices tm;
    sq_build_tile_matrices3(steps[k], n_steps[k], wt.eps[3 * lp], wt.eps[3 * lp + 1], wt.eps[3 * lp + 2], &tm);
    for (int e = 0; e < 16; ++e) P.br[k].m[e] = tm.m[e];
    P.br[k].ca = tm.ca; P.br[k].sa = tm.sa; P.br[k].cb = tm.cb; P.br[k].sb = tm.sb;
  }
  for (int k = n_bricks; k < SQ_WIN_MAX_BRICKS; ++k) P.pair[k] = 0;
  WinDev W;
  W.groupsA = wt.d_groupsA; W.clsA = wt.d_clsA; W.deltaA = wt.d_deltaA;
  W.chunksB = wt.d_chunksB; W.gbaseB = wt.d_gbaseB; W.rangesB = wt.d_rangesB; W.rchunks = wt.d_rchunks; W.clsB = wt.d_clsB; W.deltaB = wt.d_deltaB;
  W.lists = wt.d_lists; W.listidx = wt.d_listidx;
  W.LTA = wt.LTA; W.LTB = wt.LTB; W.H1 = wt.H + 1;
  W.lanes_j = wt.lanes_j;
  W.gp = wt.gp;
  W.tile_doubles = wt.tile_doubles; W.maxQ = wt.maxQ; W.maxS = wt.maxS;
  const size_t smem = sq_win_smem_bytes(wt.max_a, wt.max_b, wt.gp, wt.nbuf, wt.LTA, wt.LTB, wt.maxQ, wt.maxS, n_bricks);
  const dim3 grid((unsigned)wt.n_ranges_b, (unsigned)wt.n_groups_a);
  cudaError_t e = cudaSuccess;
  static size_t attr[3] = {0, 0, 0};
  if (smem > 48 * 1024 && smem > attr[wt.nbuf]) {
    e = cudaFuncSetAttribute(win_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(win_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(win_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(win_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) attr[wt.nbuf] = smem;
  }
  if (e == cudaSuccess) {
    const bool batch = n_states > 1;
    if (wt.nbuf == 2) {
      if (batch) win_kernel<2, true><<<grid, WIN_THREADS, smem, st>>>(state, sp->NB, W, P, n_states, state_stride);
      else win_kernel<2, false><<<grid, WIN_THREADS, smem, st>>>(state, sp->NB, W, P, 1, 0);
    } else {
      if (batch) win_kernel<1, true><<<grid, WIN_THREADS, smem, st>>>(state, sp->NB, W, P, n_states, state_stride);
      else win_kernel<1, false><<<grid, WIN_THREADS, smem, st>>>(state, sp->NB, W, P, 1, 0);
    }
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    sq_set_error("win_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}

// Gradient sweep of a window launch: out[slot0[k] + step] += <bra|T_step|ket> for every rotation step of brick k
// (evaluated before that step), then both vectors are rotated.  Both vectors must be in the sign-free gauge.
int sq_win_grad_replicas() { return WING_REPL; }

// d_out: WING_REPL replicas of n_out slots (the caller adds the replicas)
int sq_launch_win_grad(sq_space* sp, const WinTables& wt, const int* pair_idx, const TileStep* const* steps, const int* n_steps,
                       const int* slot0, int n_bricks, double* bra, double* ket, double* d_out, int n_out, cudaStream_t st) {
  if (!wt.ok || n_bricks < 1 || n_bricks > SQ_WIN_MAX_BRICKS) {
    sq_set_error("window gradient launch with %d bricks (max %d) or without tables", n_bricks, SQ_WIN_MAX_BRICKS);
    return SQ_ERR_INVALID;
  }
  WinGradProgram P;
  memset(&P, 0, sizeof(P));
  P.n = n_bricks;
  for (int k = 0; k < n_bricks; ++k) {
    const int lp = wt.pair_local[pair_idx[k]];
    if (lp < 0 || n_steps[k] < 1 || n_steps[k] > SQ_MAX_PROGRAM) {
      sq_set_error("window gradient launch: orbital pair %d is outside the window or has a bad program", pair_idx[k]);
      return SQ_ERR_INVALID;
    }
    P.pair[k] = lp;
    WinGradBrick& b = P.br[k];
    b.n = n_steps[k];
    b.slot0 = slot0[k];
    for (int q = 0; q < n_steps[k]; ++q) {
      const int kind = steps[k][q].kind;
      const int eps = wt.eps[3 * lp + kind];   // sign of T_alpha / T_beta / T_double in the window gauge
      b.kind[q] = kind | (eps < 0 ? 16 : 0);
      b.c[q] = steps[k][q].c;
      b.s[q] = eps * steps[k][q].s;
    }
  }
  WinDev W;
  W.groupsA = wt.d_groupsA; W.clsA = wt.d_clsA; W.deltaA = wt.d_deltaA;
  W.chunksB = wt.d_chunksB; W.gbaseB = wt.d_gbaseB; W.rangesB = wt.d_rangesB; W.rchunks = wt.d_rchunks; W.clsB = wt.d_clsB; W.deltaB = wt.d_deltaB;
  W.lists = wt.d_lists; W.listidx = wt.d_listidx;
  W.LTA = wt.LTA; W.LTB = wt.LTB; W.H1 = wt.H + 1;
  W.lanes_j = wt.lanes_j;
  W.gp = wt.gp;
  W.tile_doubles = wt.tile_doubles; W.maxQ = wt.maxQ; W.maxS = wt.maxS;
  const size_t smem = sq_win_smem_bytes(wt.max_a, wt.max_b, wt.gp, 2, wt.LTA, wt.LTB, wt.maxQ, wt.maxS, n_bricks) +
                      sizeof(double) * WIN_WARPS * SQ_WIN_MAX_BRICKS * SQ_MAX_PROGRAM;
  if (smem > 220 * 1024) {
    sq_set_error("window gradient launch: %zu bytes of shared memory for two %d x %d tiles", smem, wt.max_a, wt.max_b);
    return SQ_ERR_UNSUPPORTED;
  }
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(win_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      sq_set_error("win_grad_kernel: cannot get %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return SQ_ERR_CUDA;
    }
    attr = smem;
  }
  const dim3 grid((unsigned)wt.n_ranges_b, (unsigned)wt.n_groups_a);
  win_grad_kernel<<<grid, WIN_THREADS, smem, st>>>(bra, ket, sp->NB, W, P, d_out, n_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sq_set_error("win_grad_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}

// Multiply the local vector by the sign-free gauge D (an involution).  The beta gauge words live on the space.
int sq_launch_gauge(sq_space* sp, double* state, cudaStream_t st, int n_states, int64_t state_stride) {
  if (!sp->d_gwordB) {
    std::vector<uint32_t> gw((size_t)sp->NB);
    for (int64_t I = 0; I < sp->NB; ++I) gw[I] = gauge_beta_word(sp->strB[I]);
    SQ_CUDA(cudaMalloc(&sp->d_gwordB, sizeof(uint32_t) * gw.size()));
    SQ_CUDA(cudaMemcpy(sp->d_gwordB, gw.data(), sizeof(uint32_t) * gw.size(), cudaMemcpyHostToDevice));
  }
  const int64_t n_rows = sp->row_end - sp->row_begin;
  if (n_rows == 0) return SQ_OK;
  const dim3 grid((unsigned)((sp->NB + 255) / 256), (unsigned)((n_rows + 31) / 32), (unsigned)n_states);
  gauge_kernel<<<grid, 256, 0, st>>>(state, sp->NB, n_rows, sp->row_begin, sp->d_strA, sp->d_gwordB, state_stride);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sq_set_error("gauge_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}
