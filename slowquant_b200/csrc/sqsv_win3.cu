// Window kernel, version 3: register blocks over THREE orbitals, merged small tiles.
//
// win_kernel (sqsv_win.cu) sends every amplitude a brick touches through shared memory once per brick (16 bytes of
// LDS + STS traffic per touched amplitude and brick, one __syncthreads per brick) and gives every (alpha group, beta batch) its
// own CTA iteration however small its tiles are.  The ncu source view of round 2 shows what that costs: the brick phases run at
// half of the shared-memory peak for the 20 x 20 tiles, and 55 % of the batches (the classes with 1 or 6 rows / columns) hold
// 23 % of the amplitudes but pay the same per-brick barrier and loop overhead as the large ones.
//
// This kernel keeps the window decomposition, the sign-free gauge, the batches of 16 tiles (batch index fastest in shared
// memory) and the launch planner, and changes two things:
//
// 1. STEPS over orbital triples.  Take three neighbouring window orbitals {t, t+1, t+2}.  The strings of a tile fall into groups
//    that differ only on the triple: groups of 3 strings (one electron on the triple: "particle" groups, or two: "hole" groups)
//    and single strings (no or three electrons on it).  Every brick on pair (t, t+1) or (t+1, t+2) maps a 3 x 3 block
//    (row group x column group) onto itself, so a thread that holds the block in registers applies ALL bricks of the step --
//    typically a brick and its successor on the neighbouring pair of the next sublayer -- between one load and one store:
//    133 steps instead of 240 brick passes for the 16-layer tUPS circuit at 16 orbitals, 0.65 of the shared-memory traffic, half
//    of the barriers.  With the strings of a group ordered by particle position (particle groups) or by descending hole position
//    (hole groups), a brick on the lower pair acts on block positions (0,1) of a particle group and (1,2) of a hole group, the
//    upper pair the other way round, always with the source string first: position = pair ^ hole.  Every work item is a 3 x 3
//    block of 16 tiles: full items (group x group), row items (row group x three inert columns) and column items.
// 2. MERGED tiles.  A CTA stacks Ka alpha groups and Kb beta batches of one class pair so that Ka Rn x Kb Wn stays below ~450
//    amplitudes per tile slot: the small classes get as much work per barrier as the 20 x 20 class.  The per-class item lists are
//    expanded over (ka, kb) into byte offsets when the CTA starts.
//
// Reference: the bricks are the tUPS / QNP operator triples [sa_single, double, sa_single] of util.py:694-745 applied by
// construct_ups_state (operator_state_algebra.py:1002-1085); one launch replaces ~290 numba passes over the vector.
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "sqsv_internal.h"

#define W3_G 16   // tiles per batch (a thread works on two of them: 128-bit shared-memory accesses)

static int g_win3_enabled = -1;
void sq_win3_set_enabled(int on) { g_win3_enabled = on ? 1 : 0; }
bool sq_win3_enabled() {
  if (g_win3_enabled < 0) {
    const char* e = getenv("SQ_WIN3");   // off by default: measured slower than win_kernel on the B200 (profiles/r2_visit14_ab_win3.txt)
    g_win3_enabled = (e && e[0] == '1') ? 1 : 0;
  }
  return g_win3_enabled == 1;
}

// ---------------------------------------------------------------------------------------------
// The brick on a 3 x 3 block (shared by the kernel and the host emulation).  X[3 * i + j], i = row position, j = column
// position; double2 = the two tiles of a thread.  The strings of a group are ordered by the position of the particle (particle
// groups) or of the hole (hole groups), so a brick on the lower pair of the triple acts on positions (0,1) and a brick on the upper
// pair on (1,2) of EVERY group: the code is the same for all items.  What differs is the orientation -- in a hole group the source
// string of a hop (lower orbital occupied) comes second -- and that is folded into the matrices: every brick carries 8 variants,
// one per item type (0..3 full items, bit 0: hole rows, bit 1: hole columns; 4..5 row items, bit 0: hole rows, the three columns
// are inert, matrix = R_alpha x 1; 6..7 column items, bit 0: hole columns, matrix = 1 x R_beta).
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void w3_rot(double2& x0, double2& x1, double c, double s) {
  const double2 a = x0, b = x1;
  x0.x = c * a.x - s * b.x;
  x0.y = c * a.y - s * b.y;
  x1.x = c * b.x + s * a.x;
  x1.y = c * b.y + s * a.y;
}
template <int P>   // P = 0: positions (0,1) of rows and columns, third position 2; P = 1: positions (1,2), third position 0
__host__ __device__ __forceinline__ void w3_full(double2* X, const WinBrick& br) {
  constexpr int P2 = P == 0 ? 2 : 0;
  const double2 y0 = X[3 * P + P], y1 = X[3 * P + P + 1], y2 = X[3 * (P + 1) + P], y3 = X[3 * (P + 1) + P + 1];
  double2 z;
  z.x = br.m[0] * y0.x + br.m[1] * y1.x + br.m[2] * y2.x + br.m[3] * y3.x;
  z.y = br.m[0] * y0.y + br.m[1] * y1.y + br.m[2] * y2.y + br.m[3] * y3.y;
  X[3 * P + P] = z;
  z.x = br.m[4] * y0.x + br.m[5] * y1.x + br.m[6] * y2.x + br.m[7] * y3.x;
  z.y = br.m[4] * y0.y + br.m[5] * y1.y + br.m[6] * y2.y + br.m[7] * y3.y;
  X[3 * P + P + 1] = z;
  z.x = br.m[8] * y0.x + br.m[9] * y1.x + br.m[10] * y2.x + br.m[11] * y3.x;
  z.y = br.m[8] * y0.y + br.m[9] * y1.y + br.m[10] * y2.y + br.m[11] * y3.y;
  X[3 * (P + 1) + P] = z;
  z.x = br.m[12] * y0.x + br.m[13] * y1.x + br.m[14] * y2.x + br.m[15] * y3.x;
  z.y = br.m[12] * y0.y + br.m[13] * y1.y + br.m[14] * y2.y + br.m[15] * y3.y;
  X[3 * (P + 1) + P + 1] = z;
  w3_rot(X[3 * P + P2], X[3 * (P + 1) + P2], br.ca, br.sa);   // alpha single on the column the brick leaves alone
  w3_rot(X[3 * P2 + P], X[3 * P2 + P + 1], br.cb, br.sb);     // beta single on the row it leaves alone
}
// lp: 0 the brick sits on the lower two orbitals of the triple, 1 on the upper two (uniform over the CTA)
__host__ __device__ __forceinline__ void w3_apply(double2* X, int lp, const WinBrick& br) {
  if (lp == 0) w3_full<0>(X, br);
  else w3_full<1>(X, br);
}
// the 8 variants of a brick; `tm` in the natural orientation (source string = lower orbital occupied)
static void w3_variants(const WinBrick& tm, WinBrick* out8) {
  for (int t = 0; t < 8; ++t) {
    WinBrick& o = out8[t];
    if (t < 4) {
      const int fr = t & 1, fc = t >> 1, x = (fr ? 2 : 0) ^ (fc ? 1 : 0);
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) o.m[4 * a + b] = tm.m[4 * (a ^ x) + (b ^ x)];
      o.ca = tm.ca; o.sa = fr ? -tm.sa : tm.sa; o.cb = tm.cb; o.sb = fc ? -tm.sb : tm.sb;
    } else {
      const bool rows = t < 6;
      const double c = rows ? tm.ca : tm.cb, s0 = rows ? tm.sa : tm.sb, s = (t & 1) ? -s0 : s0;
      const double R[2][2] = {{c, -s}, {s, c}};
      for (int ar = 0; ar < 2; ++ar)
        for (int ac = 0; ac < 2; ++ac)
          for (int br_ = 0; br_ < 2; ++br_)
            for (int bc = 0; bc < 2; ++bc)
              o.m[4 * (2 * ar + ac) + (2 * br_ + bc)] = rows ? (ac == bc ? R[ar][br_] : 0.0) : (ar == br_ ? R[ac][bc] : 0.0);
      o.ca = rows ? c : 1.0; o.sa = rows ? s : 0.0; o.cb = rows ? 1.0 : c; o.sb = rows ? 0.0 : s;
    }
  }
}

// expanded item: row / column offsets inside the merged tile in units of 16 bytes
//   x = ro0 | ro1 << 16, y = ro2 | type << 16, z = co0 | co1 << 16, w = co2
__host__ __device__ __forceinline__ uint4 w3_expand(uint2 raw, int ka, int kb, int Rn, int Wn, int RS2, int GP2) {
  const uint32_t r0 = (raw.x & 255u) + ka * Rn, r1 = ((raw.x >> 8) & 255u) + ka * Rn, r2 = ((raw.x >> 16) & 255u) + ka * Rn;
  const uint32_t c0 = (raw.y & 255u) + kb * Wn, c1 = ((raw.y >> 8) & 255u) + kb * Wn, c2 = ((raw.y >> 16) & 255u) + kb * Wn;
  uint4 e;
  e.x = (r0 * RS2) | ((r1 * RS2) << 16);
  e.y = (r2 * RS2) | ((raw.x >> 24) << 16);
  e.z = (c0 * GP2) | ((c1 * GP2) << 16);
  e.w = c2 * GP2;
  return e;
}
// position of item i (type tp, counts cnt[8] of the un-merged list) for merge index k of K: type-major, then k, then item
__host__ __device__ __forceinline__ int w3_item_pos(int i, int tp, int k, int K, const int* cnt) {
  int toff = 0;
  for (int t = 0; t < 8; ++t)
    if (t < tp) toff += cnt[t];
  return toff * K + k * cnt[tp] + (i - toff);
}

struct Win3Dev {
  const int2* agroups;
  const int2* acls;
  const int* adelta;
  const int2* bchunks;
  const int* bgbase;
  const int2* bcls;
  const int* bdelta;
  const int4* work;
  const uint2* items;
  const int4* itemidx;
  int LTA, LTB, H1, lanes_j, gp, tile_doubles, lmax, max_rows, max_cols, max_chunks;
};

__device__ __forceinline__ void w3_cp_async8(uint32_t dst_smem, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void w3_stg_stream(double* p, double v) {
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double2 w3_lds128(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void w3_sts128(uint32_t a, double2 v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}

// Copy table of a merged batch: element x of a tile row -> {column in the vector (-1: no such tile), byte offset in the tile row}.
// Run windows: the tile index is the fastest index (16 consecutive suffixes = 128 contiguous bytes); the top window: the string
// rank is the fastest index (the columns of one tile are contiguous).  Shared by the kernel and the host emulation.
__host__ __device__ __forceinline__ int2 w3_copy_entry(int x, int lanes_j, int Wn, int GP, int mb, int Kb, int nch, const int* sbase,
                                                       const int* skcnt, const int* sdB) {
  int kb, gg, j;
  if (lanes_j) {
    kb = x / (W3_G * Wn);
    const int y = x - kb * (W3_G * Wn);
    gg = y / Wn;
    j = y - gg * Wn;
  } else {
    const int cj = x >> 4;
    gg = x & 15;
    kb = cj / Wn;
    j = cj - kb * Wn;
  }
  const int c = mb * Kb + kb;
  int2 e;
  e.y = ((kb * Wn + j) * GP + gg) * 8;
  e.x = (c < nch && gg < skcnt[c]) ? sbase[c * W3_G + gg] + sdB[j] : -1;
  return e;
}

// Shared memory: [tile: NROW x NCOL x gp doubles][expanded items: n_lists x lmax uint4][copy table: max_cols x 16 int2]
// [row table: max_rows][beta delta: LTB][tile bases: max_chunks x 16][tiles per chunk: max_chunks][items per list: W3_MAXLISTS]
template <int THREADS, bool BATCH>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 3 : 2)
win3_kernel(double* __restrict__ C0, int64_t NB, const Win3Dev W, const __grid_constant__ Win3Program P, int n_states, int64_t state_stride) {
  constexpr int WARPS = THREADS / 32, SLOTS = THREADS / 8;
  extern __shared__ double tile[];
  uint4* const sent = reinterpret_cast<uint4*>(tile + W.tile_doubles);
  int2* const selem = reinterpret_cast<int2*>(sent + P.n_lists * W.lmax);
  int* const srow = reinterpret_cast<int*>(selem + W.max_cols * W3_G);
  int* const sdB = srow + W.max_rows;
  int* const sbase = sdB + W.LTB;
  int* const skcnt = sbase + W.max_chunks * W3_G;
  int* const slcnt = skcnt + W.max_chunks;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int4 wk = __ldg(W.work + blockIdx.x);
  const int a_first = wk.x, a_cnt = wk.y & 0xffff, Ka = wk.y >> 16, b_first = wk.z, nch = wk.w & 0xffff, Kb = wk.w >> 16;
  const int clsA = __ldg(W.agroups + a_first).y, clsB = __ldg(W.bchunks + b_first).y & 0xffff;
  const int2 ca2 = __ldg(W.acls + clsA), cb2 = __ldg(W.bcls + clsB);
  const int Rn = ca2.x, Wn = cb2.x;
  const int NROW = Ka * Rn, NCOL = Kb * Wn, GP = W.gp, RS = NCOL * GP, NE = NCOL * W3_G;
  const int n_mb = (nch + Kb - 1) / Kb;   // merged batches of this CTA

  // ---- round trip 1: row table, column offsets, tile bases ----
  for (int t = tid; t < NROW; t += THREADS) {
    const int ka = t / Rn, i = t - ka * Rn;
    srow[t] = ka < a_cnt ? __ldg(W.agroups + a_first + ka).x + __ldg(W.adelta + clsA * W.LTA + i) : -1;
  }
  for (int t = tid; t < Wn; t += THREADS) sdB[t] = __ldg(W.bdelta + clsB * W.LTB + t);
  for (int t = tid; t < nch * W3_G; t += THREADS) {
    const int2 ch = __ldg(W.bchunks + b_first + (t >> 4));
    sbase[t] = __ldg(W.bgbase + ch.x * W3_G + (t & 15));
    if ((t & 15) == 0) skcnt[t >> 4] = ch.y >> 16;
  }
  __syncthreads();

  const uint32_t tb = (uint32_t)__cvta_generic_to_shared(tile);
  const uint32_t selem_s = (uint32_t)__cvta_generic_to_shared(selem);
  auto build_copy_table = [&](int mb) {
    for (int x = tid; x < NE; x += THREADS) selem[x] = w3_copy_entry(x, W.lanes_j, Wn, GP, mb, Kb, nch, sbase, skcnt, sdB);
  };
  auto issue_loads = [&](int q) {
    const double* C = BATCH ? C0 + (int64_t)(q / n_mb) * state_stride : C0;
    for (int r = warp; r < NROW; r += WARPS) {
      const int rr = srow[r];
      if (rr < 0) continue;
      const double* src = C + (int64_t)rr * NB;
      const uint32_t dst = tb + (uint32_t)(r * RS) * 8u;
#pragma unroll 4
      for (int x = lane; x < NE; x += 32) {
        int so, dof;
        asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(so), "=r"(dof) : "r"(selem_s + (uint32_t)x * 8u));
        if (so >= 0) w3_cp_async8(dst + (uint32_t)dof, src + so);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // ---- round trip 2: the first merged batch, and (meanwhile) the item lists, expanded over (ka, kb) ----
  build_copy_table(0);
  {
    const int K = Ka * Kb, RS2 = RS >> 1, GP2 = GP >> 1;
    for (int l = 0; l < P.n_lists; ++l) {
      const int4 hd = __ldg(W.itemidx + (P.list_t0[l] * W.H1 + ca2.y) * W.H1 + cb2.y);
      int cnt[8];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        cnt[t] = (hd.z >> (8 * t)) & 255;
        cnt[4 + t] = (hd.w >> (8 * t)) & 255;
      }
      const int total = hd.y * K;
      for (int u = tid; u < total; u += THREADS) {
        const int i = u / K, k = u - i * K, ka = k / Kb, kb = k - ka * Kb;
        const uint2 raw = __ldg(W.items + hd.x + i);
        sent[l * W.lmax + w3_item_pos(i, (int)(raw.x >> 24), k, K, cnt)] = w3_expand(raw, ka, kb, Rn, Wn, RS2, GP2);
      }
      if (tid == 0) slcnt[l] = total;
    }
  }
  __syncthreads();
  issue_loads(0);

  // 8 lanes x 2 tiles per item: every shared-memory access moves 16 bytes per lane, conflict-free
  const int g2 = tid & 7, slot = tid >> 3;
  const uint32_t eb = (uint32_t)__cvta_generic_to_shared(sent);
  const uint32_t tgb = tb + (uint32_t)g2 * 16u;
  const int n_total = BATCH ? n_mb * n_states : n_mb;
  for (int q = 0; q < n_total; ++q) {
    double* const C = BATCH ? C0 + (int64_t)(q / n_mb) * state_stride : C0;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();   // batch q has landed

    // ---- steps: all bricks of a step between one load and one store of a 3 x 3 block ----
    for (int s = 0; s < P.n_steps; ++s) {
      if (s) __syncthreads();
      const int l = P.step_list[s], n_items = slcnt[l];
      const int b0 = P.step_first[s], b1 = P.step_first[s + 1];
      const uint32_t el = eb + (uint32_t)(l * W.lmax) * 16u;
      for (int u = slot; u < n_items; u += SLOTS) {
        uint32_t ex, ey, ez, ew;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ex), "=r"(ey), "=r"(ez), "=r"(ew) : "r"(el + (uint32_t)u * 16u));
        const uint32_t r0 = tgb + ((ex & 0xffffu) << 4), r1 = tgb + ((ex >> 16) << 4), r2 = tgb + ((ey & 0xffffu) << 4);
        const uint32_t c0 = (ez & 0xffffu) << 4, c1 = (ez >> 16) << 4, c2 = ew << 4;
        const int type = (int)(ey >> 16);
        double2 X[9];
        X[0] = w3_lds128(r0 + c0); X[1] = w3_lds128(r0 + c1); X[2] = w3_lds128(r0 + c2);
        X[3] = w3_lds128(r1 + c0); X[4] = w3_lds128(r1 + c1); X[5] = w3_lds128(r1 + c2);
        X[6] = w3_lds128(r2 + c0); X[7] = w3_lds128(r2 + c1); X[8] = w3_lds128(r2 + c2);
        for (int b = b0; b < b1; ++b) w3_apply(X, P.brick_lp[b], P.brv[b][type]);
        w3_sts128(r0 + c0, X[0]); w3_sts128(r0 + c1, X[1]); w3_sts128(r0 + c2, X[2]);
        w3_sts128(r1 + c0, X[3]); w3_sts128(r1 + c1, X[4]); w3_sts128(r1 + c2, X[5]);
        w3_sts128(r2 + c0, X[6]); w3_sts128(r2 + c1, X[7]); w3_sts128(r2 + c2, X[8]);
      }
    }
    __syncthreads();

    // ---- store ----
    for (int r = warp; r < NROW; r += WARPS) {
      const int rr = srow[r];
      if (rr < 0) continue;
      double* dst = C + (int64_t)rr * NB;
      const uint32_t srct = tb + (uint32_t)(r * RS) * 8u;
#pragma unroll 4
      for (int x = lane; x < NE; x += 32) {
        int so, dof;
        asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(so), "=r"(dof) : "r"(selem_s + (uint32_t)x * 8u));
        if (so >= 0) {
          double v;
          asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(srct + (uint32_t)dof));
          w3_stg_stream(dst + so, v);
        }
      }
    }
    if (q + 1 < n_total) {
      __syncthreads();   // the tile buffer and the copy table are free again
      const int mb1 = BATCH ? (q + 1) % n_mb : q + 1;
      if (!BATCH || n_mb > 1) {
        build_copy_table(mb1);
        __syncthreads();
      }
      issue_loads(q + 1);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host tables
// ---------------------------------------------------------------------------------------------
template <typename T>
static int w3_upload(T** d, const std::vector<T>& v) {
  *d = nullptr;
  if (v.empty()) return SQ_OK;
  SQ_CUDA(cudaMalloc(d, sizeof(T) * v.size()));
  SQ_CUDA(cudaMemcpy(*d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return SQ_OK;
}

void sq_free_win3(Win3Tables* w3) {
  if (!w3) return;
  cudaFree(w3->d_agroups); cudaFree(w3->d_acls); cudaFree(w3->d_adelta);
  cudaFree(w3->d_bchunks); cudaFree(w3->d_bgbase); cudaFree(w3->d_bcls); cudaFree(w3->d_bdelta);
  cudaFree(w3->d_work); cudaFree(w3->d_items); cudaFree(w3->d_itemidx);
  delete w3;
}

// groups of the window strings `wl` (ranks = positions in wl) with respect to the triple {t0, t0+1, t0+2}
struct TripleSide {
  std::vector<std::array<int, 3>> groups;   // string ranks in block order (see the header of this file)
  std::vector<char> hole;                   // 1: two electrons on the triple
  std::vector<int> inert;                   // strings with no or three electrons on the triple
};
static void triple_side(const std::vector<uint32_t>& wl, int t0, TripleSide* out) {
  out->groups.clear(); out->hole.clear(); out->inert.clear();
  const uint32_t T = 7u << t0;
  std::map<uint32_t, int> pos;
  for (size_t j = 0; j < wl.size(); ++j) pos[wl[j]] = (int)j;
  for (size_t j = 0; j < wl.size(); ++j) {
    const uint32_t w = wl[j], rest = w & ~T;
    const int n = __builtin_popcount(w & T);
    if (n == 0 || n == 3) { out->inert.push_back((int)j); continue; }
    // the group is emitted once, when its first member in block order is met
    std::array<int, 3> grp;
    bool complete = true;
    for (int k = 0; k < 3; ++k) {
      // particle groups: block position k = particle on orbital t0 + k; hole groups: position k = hole on orbital t0 + k
      const uint32_t part = (n == 1) ? (1u << (t0 + k)) : (T & ~(1u << (t0 + k)));
      auto it = pos.find(rest | part);
      if (it == pos.end()) { complete = false; break; }
      grp[k] = it->second;
    }
    if (!complete) { out->inert.push_back((int)j); continue; }   // cannot happen for complete string lists; stay safe
    if (grp[0] != (int)j) continue;
    out->groups.push_back(grp);
    out->hole.push_back(n == 2 ? 1 : 0);
  }
}

int sq_build_win3(sq_space* sp, WinTables* wt, const SideHost& hA, const SideHost& hB) {
  Win3Tables* w3 = new Win3Tables();
  wt->w3 = w3;
  const int H = wt->H, H1 = H + 1;
  if (H < 3 || H - 2 > W3_MAXLISTS) return SQ_OK;
  w3->H = H;
  w3->LTA = wt->LTA;
  w3->LTB = wt->LTB;
  w3->gp = wt->gp;
  w3->lanes_j = wt->lanes_j;
  w3->acls = hA.cls;
  w3->bcls = hB.cls;
  w3->adelta = hA.delta;   // padded to LTA / LTB by the caller
  w3->bdelta = hB.delta;
  w3->bgbase = hB.gbase;
  // class-major group / chunk lists (memory order inside a class)
  w3->agroups = hA.groups;
  std::stable_sort(w3->agroups.begin(), w3->agroups.end(), [](const int2& a, const int2& b) { return a.y != b.y ? a.y < b.y : a.x < b.x; });
  w3->bchunks = hB.groups;
  std::stable_sort(w3->bchunks.begin(), w3->bchunks.end(), [](const int2& a, const int2& b) {
    return (a.y & 0xffff) != (b.y & 0xffff) ? (a.y & 0xffff) < (b.y & 0xffff) : a.x < b.x;
  });
  std::vector<int> afirst(hA.ncls + 1, 0), bfirst(hB.ncls + 1, 0);
  for (const int2& g : w3->agroups) ++afirst[g.y + 1];
  for (const int2& c : w3->bchunks) ++bfirst[(c.y & 0xffff) + 1];
  for (int c = 0; c < hA.ncls; ++c) afirst[c + 1] += afirst[c];
  for (int c = 0; c < hB.ncls; ++c) bfirst[c + 1] += bfirst[c];

  // ---- item lists per (triple, e_wa, e_wb) ----
  w3->itemidx.assign((size_t)(H - 2) * H1 * H1, make_int4(0, 0, 0, 0));
  for (int t0 = 0; t0 + 3 <= H; ++t0) {
    std::vector<TripleSide> sa(H1), sb(H1);
    for (int e = 0; e <= H; ++e) {
      triple_side(hA.wl[e], t0, &sa[e]);
      triple_side(hB.wl[e], t0, &sb[e]);
    }
    for (int ea = 0; ea <= H; ++ea)
      for (int eb = 0; eb <= H; ++eb) {
        std::vector<uint2> byType[8];
        auto pack = [](const int* r, const int* c, int type) {
          return make_uint2((uint32_t)r[0] | ((uint32_t)r[1] << 8) | ((uint32_t)r[2] << 16) | ((uint32_t)type << 24),
                            (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16));
        };
        const TripleSide &A = sa[ea], &B = sb[eb];
        for (size_t ia = 0; ia < A.groups.size(); ++ia)
          for (size_t ib = 0; ib < B.groups.size(); ++ib) {
            const int type = A.hole[ia] | (B.hole[ib] << 1);
            byType[type].push_back(pack(A.groups[ia].data(), B.groups[ib].data(), type));
          }
        // row items: a row group x three inert columns (the last chunk repeats its last column: the thread computes and
        // stores the same value twice)
        for (size_t ia = 0; ia < A.groups.size(); ++ia)
          for (size_t k = 0; k < B.inert.size(); k += 3) {
            int c[3];
            for (int q = 0; q < 3; ++q) c[q] = B.inert[std::min(k + q, B.inert.size() - 1)];
            byType[4 + A.hole[ia]].push_back(pack(A.groups[ia].data(), c, 4 + A.hole[ia]));
          }
        for (size_t ib = 0; ib < B.groups.size(); ++ib)
          for (size_t k = 0; k < A.inert.size(); k += 3) {
            int r[3];
            for (int q = 0; q < 3; ++q) r[q] = A.inert[std::min(k + q, A.inert.size() - 1)];
            byType[6 + B.hole[ib]].push_back(pack(r, B.groups[ib].data(), 6 + B.hole[ib]));
          }
        int4 idx = make_int4((int)w3->items.size(), 0, 0, 0);
        for (int t = 0; t < 8; ++t) {
          if (byType[t].size() > 255) return SQ_OK;   // byte counters; windows this large are not planned anyway
          idx.y += (int)byType[t].size();
          if (t < 4) idx.z |= (int)byType[t].size() << (8 * t);
          else idx.w |= (int)byType[t].size() << (8 * (t - 4));
          w3->items.insert(w3->items.end(), byType[t].begin(), byType[t].end());
        }
        w3->itemidx[((size_t)t0 * H1 + ea) * H1 + eb] = idx;
      }
  }
  auto max_items = [&](int ea, int eb) {
    int m = 0;
    for (int t0 = 0; t0 + 3 <= H; ++t0) m = std::max(m, w3->itemidx[((size_t)t0 * H1 + ea) * H1 + eb].y);
    return m;
  };

  // ---- merge factors and CTA work items per class pair ----
  static int tmax_env = -1, rb_env = -1;
  if (tmax_env < 0) {
    const char* e = getenv("SQ_WIN3_TMAX");   // amplitudes per merged tile slot
    tmax_env = e ? std::max(1, atoi(e)) : 0;
    const char* r = getenv("SQ_WIN3_RANGE");  // merged batches per CTA
    rb_env = r ? std::max(1, atoi(r)) : 6;
  }
  int TMAX = tmax_env ? tmax_env : (wt->gp == 16 ? 448 : 400);
  TMAX = std::max(TMAX, hA.max_cnt * hB.max_cnt);
  const int RCAP = 64, CCAP = 40;   // rows / columns of a merged tile (table sizes; 8-bit string ranks stay tile-local)
  struct WorkW { int4 w; int64_t weight; };
  std::vector<WorkW> work;
  int max_rows = 1, max_cols = 1, max_chunks = 1, lmax = 1, tile_amps = 1;
  for (int ca = 0; ca < hA.ncls; ++ca)
    for (int cb = 0; cb < hB.ncls; ++cb) {
      const int nga = afirst[ca + 1] - afirst[ca], nch = bfirst[cb + 1] - bfirst[cb];
      if (!nga || !nch) continue;
      const int Rn = hA.cls[ca].x, Wn = hB.cls[cb].x, ea = hA.cls[ca].y, eb = hB.cls[cb].y;
      const int mi = max_items(ea, eb);
      if (mi == 0) continue;   // no brick of this window touches these tiles
      const int K = std::max(1, TMAX / (Rn * Wn));
      int Ka = std::max(1, std::min(std::min(K, nga), std::max(1, RCAP / Rn)));
      Ka = (nga + ((nga + Ka - 1) / Ka) - 1) / ((nga + Ka - 1) / Ka);   // balanced
      int Kb = std::max(1, std::min(std::min(K / Ka, nch), std::max(1, CCAP / Wn)));
      const int n_mb = (nch + Kb - 1) / Kb;
      Kb = (nch + n_mb - 1) / n_mb;
      if (Ka * Rn > 255 || Kb * Wn > 255) return SQ_OK;
      const int per_cta = std::max(1, std::min(rb_env, 64 / Kb));   // merged batches per CTA
      const int n_mb2 = (nch + Kb - 1) / Kb, parts = (n_mb2 + per_cta - 1) / per_cta;
      for (int a0 = 0; a0 < nga; a0 += Ka) {
        const int a_cnt = std::min(Ka, nga - a0);
        for (int q = 0; q < parts; ++q) {
          const int m0 = (int)((int64_t)n_mb2 * q / parts), m1 = (int)((int64_t)n_mb2 * (q + 1) / parts);
          const int c0 = m0 * Kb, c1 = std::min(nch, m1 * Kb);
          if (c1 <= c0) continue;
          int64_t tiles = 0;
          for (int c = c0; c < c1; ++c) tiles += w3->bchunks[bfirst[cb] + c].y >> 16;
          work.push_back({make_int4(afirst[ca] + a0, a_cnt | (Ka << 16), bfirst[cb] + c0, (c1 - c0) | (Kb << 16)),
                          tiles * a_cnt * Rn * Wn});
          max_chunks = std::max(max_chunks, c1 - c0);
        }
      }
      max_rows = std::max(max_rows, Ka * Rn);
      max_cols = std::max(max_cols, Kb * Wn);
      lmax = std::max(lmax, mi * Ka * Kb);
      tile_amps = std::max(tile_amps, Ka * Rn * Kb * Wn);
    }
  if (work.empty()) return SQ_OK;
  std::stable_sort(work.begin(), work.end(), [](const WorkW& a, const WorkW& b) { return a.weight > b.weight; });
  for (const WorkW& x : work) w3->work.push_back(x.w);
  w3->max_rows = (max_rows + 3) & ~3;
  w3->max_cols = (max_cols + 1) & ~1;
  w3->max_chunks = (max_chunks + 3) & ~3;
  w3->lmax = lmax;
  w3->tile_doubles = tile_amps * wt->gp;   // even
  if ((size_t)w3->tile_doubles * 8 / 16 > 65535) return SQ_OK;   // 16-bit offsets in units of 16 bytes
  if (sp->device >= 0) {
    SQ_CUDA(cudaSetDevice(sp->device));
    SQ_CHECK(w3_upload(&w3->d_agroups, w3->agroups));
    SQ_CHECK(w3_upload(&w3->d_acls, w3->acls));
    SQ_CHECK(w3_upload(&w3->d_adelta, w3->adelta));
    SQ_CHECK(w3_upload(&w3->d_bchunks, w3->bchunks));
    SQ_CHECK(w3_upload(&w3->d_bgbase, w3->bgbase));
    SQ_CHECK(w3_upload(&w3->d_bcls, w3->bcls));
    SQ_CHECK(w3_upload(&w3->d_bdelta, w3->bdelta));
    SQ_CHECK(w3_upload(&w3->d_work, w3->work));
    SQ_CHECK(w3_upload(&w3->d_items, w3->items));
    SQ_CHECK(w3_upload(&w3->d_itemidx, w3->itemidx));
  }
  w3->ok = true;
  if (getenv("SQ_PLAN_DEBUG")) {
    int64_t mb = 0, kamax = 0, kbmax = 0;
    for (const int4& x : w3->work) {
      mb += ((x.w & 0xffff) + (x.w >> 16) - 1) / (x.w >> 16);
      kamax = std::max<int64_t>(kamax, x.y >> 16);
      kbmax = std::max<int64_t>(kbmax, x.w >> 16);
    }
    fprintf(stderr, "win3 [%d,%d): %zu CTAs, %lld merged batches, Ka <= %lld, Kb <= %lld, lmax %d, tile %d doubles, rows <= %d, chunks <= %d, "
            "%zu raw items\n", wt->w0, wt->w0 + H, w3->work.size(), (long long)mb, (long long)kamax, (long long)kbmax, w3->lmax,
            w3->tile_doubles, w3->max_rows, w3->max_chunks, w3->items.size());
  }
  return SQ_OK;
}

static size_t w3_smem_bytes(const Win3Tables& w3, int n_lists) {
  return sizeof(double) * (size_t)w3.tile_doubles + (size_t)n_lists * w3.lmax * 16 + (size_t)w3.max_cols * W3_G * 8 +
         4 * (size_t)(w3.max_rows + w3.LTB + w3.max_chunks * W3_G + w3.max_chunks + W3_MAXLISTS + 2);
}

// ---------------------------------------------------------------------------------------------
// Steps of a launch.  Bricks in program order (a valid execution order); a brick may move in front of earlier bricks it shares
// no orbital with.  Greedy: per step take the triple whose closure -- bricks inside the triple that can run now, in program
// order -- is largest.
// ---------------------------------------------------------------------------------------------
int sq_win3_program(const WinTables& wt, const int* pair_idx, const TileStep* const* steps, const int* n_steps, int n_bricks,
                    Win3Program* P) {
  memset(P, 0, sizeof(*P));
  const int H = wt.H;
  std::vector<int> lo(n_bricks), lpair(n_bricks);
  for (int k = 0; k < n_bricks; ++k) {
    lpair[k] = wt.pair_local[pair_idx[k]];
    if (lpair[k] < 0) {
      sq_set_error("window launch: orbital pair %d is outside the window", pair_idx[k]);
      return SQ_ERR_INVALID;
    }
    lo[k] = wt.pair_lo[lpair[k]];
  }
  std::vector<char> done(n_bricks, 0);
  int n_done = 0, nb_out = 0;
  std::map<int, int> list_of;   // t0 -> list
  while (n_done < n_bricks) {
    int best_t0 = -1;
    std::vector<int> best;
    for (int t0 = 0; t0 + 3 <= H; ++t0) {
      std::vector<int> sel;
      std::vector<char> taken(n_bricks, 0);
      for (int j = 0; j < n_bricks; ++j) {
        if (done[j] || (lo[j] != t0 && lo[j] != t0 + 1)) continue;
        bool free_ = true;
        for (int i = 0; i < j && free_; ++i)
          if (!done[i] && !taken[i] && std::abs(lo[i] - lo[j]) < 2) free_ = false;
        if (free_) {
          taken[j] = 1;
          sel.push_back(j);
        }
      }
      if (sel.size() > best.size()) {
        best = sel;
        best_t0 = t0;
      }
    }
    if (best.empty() || P->n_steps >= W3_MAXSTEPS) {
      sq_set_error("window launch: step grouping failed (%d of %d bricks placed)", n_done, n_bricks);
      return SQ_ERR_INVALID;
    }
    auto it = list_of.find(best_t0);
    if (it == list_of.end()) {
      if (P->n_lists >= W3_MAXLISTS) {
        sq_set_error("window launch: more than %d orbital triples", W3_MAXLISTS);
        return SQ_ERR_INVALID;
      }
      P->list_t0[P->n_lists] = best_t0;
      it = list_of.emplace(best_t0, P->n_lists++).first;
    }
    const int s = P->n_steps++;
    P->step_list[s] = (unsigned char)it->second;
    P->step_first[s] = (unsigned char)nb_out;
    for (int j : best) {
      done[j] = 1;
      ++n_done;
      TileMatrices tm;
      const int lp = lpair[j];
      sq_build_tile_matrices3(steps[j], n_steps[j], wt.eps[3 * lp], wt.eps[3 * lp + 1], wt.eps[3 * lp + 2], &tm);
      WinBrick br;
      if (wt.pair_flip[lp]) {
        // the pair's source orbital is the upper one: source and target strings change places in both spins
        for (int a = 0; a < 4; ++a)
          for (int b = 0; b < 4; ++b) br.m[4 * a + b] = tm.m[4 * (3 - a) + (3 - b)];
        br.ca = tm.ca; br.sa = -tm.sa; br.cb = tm.cb; br.sb = -tm.sb;
      } else {
        for (int e = 0; e < 16; ++e) br.m[e] = tm.m[e];
        br.ca = tm.ca; br.sa = tm.sa; br.cb = tm.cb; br.sb = tm.sb;
      }
      w3_variants(br, P->brv[nb_out]);
      P->brick_lp[nb_out] = (unsigned char)(lo[j] - best_t0);
      ++nb_out;
    }
    P->step_first[s + 1] = (unsigned char)nb_out;
  }
  return SQ_OK;
}

int sq_launch_win3(sq_space* sp, const WinTables& wt, const Win3Program& P, double* state, cudaStream_t st, int n_states,
                   int64_t state_stride) {
  const Win3Tables& w3 = *wt.w3;
  Win3Dev W;
  W.agroups = w3.d_agroups; W.acls = w3.d_acls; W.adelta = w3.d_adelta;
  W.bchunks = w3.d_bchunks; W.bgbase = w3.d_bgbase; W.bcls = w3.d_bcls; W.bdelta = w3.d_bdelta;
  W.work = w3.d_work; W.items = w3.d_items; W.itemidx = w3.d_itemidx;
  W.LTA = w3.LTA; W.LTB = w3.LTB; W.H1 = w3.H + 1; W.lanes_j = w3.lanes_j; W.gp = w3.gp;
  W.tile_doubles = w3.tile_doubles; W.lmax = w3.lmax; W.max_rows = w3.max_rows; W.max_cols = w3.max_cols; W.max_chunks = w3.max_chunks;
  const size_t smem = w3_smem_bytes(w3, P.n_lists);
  if (smem > 220 * 1024) {
    sq_set_error("win3_kernel: %zu bytes of shared memory", smem);
    return SQ_ERR_UNSUPPORTED;
  }
  static int threads = 0;
  if (!threads) {
    const char* e = getenv("SQ_WIN3_THREADS");
    threads = (e && atoi(e) == 384) ? 384 : 256;
  }
  static size_t attr = 0;
  cudaError_t e = cudaSuccess;
  if (smem > 48 * 1024 && smem > attr) {
    e = cudaFuncSetAttribute(win3_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(win3_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(win3_kernel<384, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(win3_kernel<384, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) attr = smem;
  }
  if (e == cudaSuccess) {
    const unsigned grid = (unsigned)w3.work.size();
    const bool batch = n_states > 1;
    if (threads == 384) {
      if (batch) win3_kernel<384, true><<<grid, 384, smem, st>>>(state, sp->NB, W, P, n_states, state_stride);
      else win3_kernel<384, false><<<grid, 384, smem, st>>>(state, sp->NB, W, P, 1, 0);
    } else {
      if (batch) win3_kernel<256, true><<<grid, 256, smem, st>>>(state, sp->NB, W, P, n_states, state_stride);
      else win3_kernel<256, false><<<grid, 256, smem, st>>>(state, sp->NB, W, P, 1, 0);
    }
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    sq_set_error("win3_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Host emulation of win3_kernel, CTA by CTA, on the same tables, expanded item lists and block arithmetic (w3_expand,
// w3_item_pos, w3_apply).  TEST INFRASTRUCTURE: lets the CPU tests check the table builders, the step grouping and the block
// algebra against the oracle on host-only spaces; no product call reaches it (sq_debug_win3_emulate is its only caller).
// ---------------------------------------------------------------------------------------------
static inline uint32_t w3_gauge_beta_word(uint32_t mB) {
  uint32_t w = 0, par = 0;
  for (int p = 0; p < 32; ++p) {
    if (par) w |= 1u << p;
    if (mB & (1u << p)) par ^= 1u;
  }
  return w;
}
void sq_gauge_host(const sq_space* sp, double* x) {
  std::vector<uint32_t> gw((size_t)sp->NB);
  for (int64_t b = 0; b < sp->NB; ++b) gw[b] = w3_gauge_beta_word(sp->strB[b]);
  for (int64_t r = sp->row_begin; r < sp->row_end; ++r)
    for (int64_t b = 0; b < sp->NB; ++b)
      if (__builtin_popcount(sp->strA[r] & gw[b]) & 1) x[(r - sp->row_begin) * sp->NB + b] = -x[(r - sp->row_begin) * sp->NB + b];
}

int sq_win3_emulate_host(const sq_space* sp, const WinTables& wt, const Win3Program& P, double* C) {
  if (!wt.w3 || !wt.w3->ok) {
    sq_set_error("win3 emulation: the window has no version-3 tables");
    return SQ_ERR_UNSUPPORTED;
  }
  const Win3Tables& W = *wt.w3;
  const int64_t NB = sp->NB;
  const int H1 = W.H + 1, GP = W.gp;
  std::vector<double> tile((size_t)W.tile_doubles);
  std::vector<uint4> sent((size_t)P.n_lists * W.lmax);
  std::vector<int> slcnt(W3_MAXLISTS, 0);
  for (const int4& wk : W.work) {
    const int a_first = wk.x, a_cnt = wk.y & 0xffff, Ka = wk.y >> 16, b_first = wk.z, nch = wk.w & 0xffff, Kb = wk.w >> 16;
    const int clsA = W.agroups[a_first].y, clsB = W.bchunks[b_first].y & 0xffff;
    const int Rn = W.acls[clsA].x, Wn = W.bcls[clsB].x, ea = W.acls[clsA].y, eb = W.bcls[clsB].y;
    const int NROW = Ka * Rn, NCOL = Kb * Wn, RS = NCOL * GP;
    const int n_mb = (nch + Kb - 1) / Kb;
    if ((size_t)NROW * RS > tile.size() || NROW > W.max_rows || nch > W.max_chunks) {
      sq_set_error("win3 emulation: work item exceeds the table sizes");
      return SQ_ERR_INVALID;
    }
    std::vector<int> srow(NROW);
    for (int t = 0; t < NROW; ++t) {
      const int ka = t / Rn, i = t - ka * Rn;
      srow[t] = ka < a_cnt ? W.agroups[a_first + ka].x + W.adelta[clsA * W.LTA + i] : -1;
    }
    const int K = Ka * Kb;
    for (int l = 0; l < P.n_lists; ++l) {
      const int4 hd = W.itemidx[((size_t)P.list_t0[l] * H1 + ea) * H1 + eb];
      int cnt[8];
      for (int t = 0; t < 4; ++t) {
        cnt[t] = (hd.z >> (8 * t)) & 255;
        cnt[4 + t] = (hd.w >> (8 * t)) & 255;
      }
      const int total = hd.y * K;
      if (total > W.lmax) {
        sq_set_error("win3 emulation: expanded list longer than lmax");
        return SQ_ERR_INVALID;
      }
      std::vector<char> filled(total, 0);
      for (int u = 0; u < total; ++u) {
        const int i = u / K, k = u - i * K, ka = k / Kb, kb = k - ka * Kb;
        const uint2 raw = W.items[hd.x + i];
        const int pos = w3_item_pos(i, (int)(raw.x >> 24), k, K, cnt);
        if (pos < 0 || pos >= total || filled[pos]) {
          sq_set_error("win3 emulation: item positions are not a permutation");
          return SQ_ERR_INVALID;
        }
        filled[pos] = 1;
        sent[(size_t)l * W.lmax + pos] = w3_expand(raw, ka, kb, Rn, Wn, RS >> 1, GP >> 1);
      }
      slcnt[l] = total;
    }
    for (int mb = 0; mb < n_mb; ++mb) {
      std::fill(tile.begin(), tile.end(), 0.0);
      // the copy table of the kernel (w3_copy_entry), for both of its lane orders; every tile element exactly once
      std::vector<int> sbase((size_t)nch * W3_G), skcnt(nch), sdB(Wn);
      for (int c = 0; c < nch; ++c) {
        const int2 ch = W.bchunks[b_first + c];
        skcnt[c] = ch.y >> 16;
        for (int g = 0; g < W3_G; ++g) sbase[(size_t)c * W3_G + g] = W.bgbase[ch.x * W3_G + g];
      }
      for (int j = 0; j < Wn; ++j) sdB[j] = W.bdelta[clsB * W.LTB + j];
      const int NE = NCOL * W3_G;
      if (NCOL > W.max_cols) {
        sq_set_error("win3 emulation: merged tile wider than max_cols");
        return SQ_ERR_INVALID;
      }
      std::vector<int2> selem(NE);
      std::vector<char> hit((size_t)RS, 0);
      for (int x = 0; x < NE; ++x) {
        selem[x] = w3_copy_entry(x, W.lanes_j, Wn, GP, mb, Kb, nch, sbase.data(), skcnt.data(), sdB.data());
        const int d = selem[x].y / 8;
        if (selem[x].y % 8 || d < 0 || d >= RS || hit[d]) {
          sq_set_error("win3 emulation: copy table is not a one-to-one map into the tile row");
          return SQ_ERR_INVALID;
        }
        hit[d] = 1;
      }
      auto copy = [&](bool load) {
        for (int r = 0; r < NROW; ++r) {
          if (srow[r] < 0) continue;
          for (int x = 0; x < NE; ++x) {
            if (selem[x].x < 0) continue;
            double& t = tile[(size_t)r * RS + selem[x].y / 8];
            double& xx = C[(int64_t)srow[r] * NB + selem[x].x];
            if (load) t = xx;
            else xx = t;
          }
        }
      };
      copy(true);
      for (int s = 0; s < P.n_steps; ++s) {
        const int l = P.step_list[s];
        for (int u = 0; u < slcnt[l]; ++u) {
          const uint4 e = sent[(size_t)l * W.lmax + u];
          const uint32_t ro[3] = {(e.x & 0xffffu) << 4, (e.x >> 16) << 4, (e.y & 0xffffu) << 4};
          const uint32_t co[3] = {(e.z & 0xffffu) << 4, (e.z >> 16) << 4, e.w << 4};
          const int type = (int)(e.y >> 16);
          for (int g2 = 0; g2 < 8; ++g2) {
            double2 X[9];
            double* p[9];
            for (int i = 0; i < 3; ++i)
              for (int j = 0; j < 3; ++j) {
                const size_t byte = (size_t)ro[i] + co[j] + (size_t)g2 * 16;
                if (byte + 16 > tile.size() * 8) {
                  sq_set_error("win3 emulation: item offset outside the tile");
                  return SQ_ERR_INVALID;
                }
                p[3 * i + j] = tile.data() + byte / 8;
                X[3 * i + j] = make_double2(p[3 * i + j][0], p[3 * i + j][1]);
              }
            for (int b = P.step_first[s]; b < P.step_first[s + 1]; ++b) w3_apply(X, P.brick_lp[b], P.brv[b][type]);
            for (int q = 0; q < 9; ++q) {
              p[q][0] = X[q].x;
              p[q][1] = X[q].y;
            }
          }
        }
      }
      copy(false);
    }
  }
  return SQ_OK;
}

// debug (SQ_PLAN_DEBUG): work of one launch counted from the tables -- merged batches, item rounds per CTA slot, brick applications
void sq_win3_print_stats(const WinTables& wt, const int* pair_idx, int n_bricks) {
  if (!wt.w3 || !wt.w3->ok) return;
  TileStep one[1] = {{0, 1.0, 0.0}};
  const TileStep* sp[SQ_WIN_MAX_BRICKS];
  int ns[SQ_WIN_MAX_BRICKS];
  for (int k = 0; k < n_bricks; ++k) { sp[k] = one; ns[k] = 1; }
  Win3Program P;
  if (sq_win3_program(wt, pair_idx, sp, ns, n_bricks, &P) != SQ_OK) return;
  const Win3Tables& W = *wt.w3;
  const int H1 = W.H + 1;
  double mbs = 0, rounds = 0, items = 0, full_apps = 0, rc_apps = 0, amps = 0, ideal_items = 0;
  for (const int4& wk : W.work) {
    const int a_cnt = wk.y & 0xffff, Ka = wk.y >> 16, nch = wk.w & 0xffff, Kb = wk.w >> 16;
    const int clsA = W.agroups[wk.x].y, clsB = W.bchunks[wk.z].y & 0xffff;
    const int Rn = W.acls[clsA].x, Wn = W.bcls[clsB].x, ea = W.acls[clsA].y, eb = W.bcls[clsB].y;
    const int n_mb = (nch + Kb - 1) / Kb, K = Ka * Kb;
    double tiles = 0;
    for (int c = 0; c < nch; ++c) tiles += W.bchunks[wk.z + c].y >> 16;
    amps += tiles * a_cnt * Rn * Wn;
    mbs += n_mb;
    for (int s = 0; s < P.n_steps; ++s) {
      const int4 hd = W.itemidx[((size_t)P.list_t0[P.step_list[s]] * H1 + ea) * H1 + eb];
      const int nb = P.step_first[s + 1] - P.step_first[s];
      const int nfull = (hd.z & 255) + ((hd.z >> 8) & 255) + ((hd.z >> 16) & 255) + ((hd.z >> 24) & 255);
      const int n = hd.y * K;
      rounds += (double)n_mb * ((n + 31) / 32);
      items += (double)n_mb * n;
      ideal_items += tiles / 16.0 * a_cnt * hd.y;
      full_apps += (double)n_mb * nfull * K * nb;
      rc_apps += (double)n_mb * (hd.y - nfull) * K * nb;
    }
  }
  fprintf(stderr, "win3 stats [%d,%d): %d bricks in %d steps, %zu CTAs, %.0f merged batches, %.3g amplitudes; items %.4g (%.4g without merge "
          "padding), rounds of 32 slots %.4g (slot use %.2f); DP warp instructions %.4g M\n", wt.w0, wt.w0 + wt.H, n_bricks, P.n_steps,
          W.work.size(), mbs, amps, items, ideal_items, rounds, items / (32.0 * rounds), (full_apps * 48 + rc_apps * 24) * 8 / 32 / 1e6);
}
