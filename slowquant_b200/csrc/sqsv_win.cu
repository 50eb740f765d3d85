// "window" kernel: a whole group of bricks in ONE read + ONE write of the CI vector, through shared memory.
//
// Pick an orbital window [w0, w0+H) per spin.  A string is (prefix on orbitals < w0, window part, suffix on
// orbitals >= w0+H).  An operator whose orbitals all lie inside the window never changes prefix or suffix,
// conserves the electron count e_w of the window part, and its fermionic sign depends on window bits only.
// Hence the coefficient matrix C[Ia][Ib] decomposes into independent tiles
//     (alpha prefix, alpha suffix; beta prefix, beta suffix)  x  (C(Ha,e_wa) rows) x (C(Hb,e_wb) columns)
// and EVERY brick inside the window maps each tile onto itself.  In itertools.combinations order
//     I(prefix, w, suffix) = start(prefix) + off_{e_rem}(w) + rank(suffix),
// so the rows/columns of a tile sit at base + delta[class][j] (class = (electrons left after the prefix, e_w),
// j = rank of the window part), and consecutive suffixes of the same (prefix, e_w) are consecutive indices.
//   * beta "run" windows (large suffix): a CTA takes K consecutive suffixes, so every tile column is a
//     contiguous K-double run in memory (K = 16 -> 128 B);
//   * beta "block" windows (the window reaches the last orbital): delta[j] = j, the tile columns are one
//     contiguous segment.
// A CTA loads its tile (rows need no contiguity: one row is one coalesced stream), applies up to
// SQ_WIN_MAX_BRICKS bricks on it in shared memory with the gauge-fixed 4x4 / 2x2 matrices of tile_kernel_v2,
// and stores it back.  The commutation-aware planner of sqsv_api.cu chooses windows and brick groups; a tUPS
// layer at n = 16 needs ~3 sweeps of the vector (often fewer, bricks of the next layer ride along) instead of
// 15 (tile_kernel_v2) or 8 (quad_kernel).
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <unordered_map>

#include "sqsv_internal.h"

#define WIN_THREADS 256
#define WIN_WARPS (WIN_THREADS / 32)

struct WinSideDev {
  const int2* groups;      // alpha: {first row (shard-relative), class}; beta: {first column, class | n_suffix << 16}
  const int* delta;        // [ncls][LT] offset of window string j from the group base
  const uint32_t* gbits;   // [ncls][LT] gauge word of window string j: alpha = occupation mask, beta = parity-prefix mask
  const int* cnt;          // [ncls] number of window strings
  const uint32_t* items;   // [n_pairs][ncls][LT] work items of a brick: src strings first, then inert ones
  const int2* itemcnt;     // [n_pairs][ncls] {n_src, n_inert}
  int LT, ncls;
};

struct WinBrick {
  double m[16];            // 4x4 on (x[r][c], x[r][c'], x[r'][c], x[r'][c']), row-major
  double ca, sa, cb, sb;   // alpha single on (x[r][c], x[r'][c]); beta single on (x[r][c], x[r][c'])
};
struct WinProgram {
  int n;
  int pair[SQ_WIN_MAX_BRICKS];
  WinBrick br[SQ_WIN_MAX_BRICKS];
};

__device__ __forceinline__ double wflip(double x, int neg) {
  return __hiloint2double(__double2hiint(x) ^ (neg << 31), __double2loint(x));
}

// item code: bits 9:0 window string j, 19:10 partner j' (src only)
#define IT_J(c) ((int)((c)&1023u))
#define IT_JP(c) ((int)(((c) >> 10) & 1023u))

// tile data is touched once per sweep: keep it out of L1 so that the (small, hot) window tables stay there
__device__ __forceinline__ double ldg_stream(const double* p) {
  double v;
  asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void stg_stream(double* p, double v) {
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// Sign-free gauge.  The reference orders spin orbitals a0 b0 a1 b1 ...; re-ordering them as (all alpha)(all beta)
// multiplies determinant |A,B> by D(A,B) = (-1)^{#{(p,q): p in A, q in B, q < p}}.  In that gauge every
// nearest-neighbour hop p <-> p+1 of either spin and the pair double carry a constant sign (checked on the host
// per pair, folded into the brick matrices), so a brick is ONE constant 4x4 matrix / 2x2 rotation for all tiles.
// The window-local part of D is applied when a tile is loaded and again when it is stored.
//
// Brick work inside a tile, flat over the CTA's threads (one item per thread per step):
//   (src row item) x (src column item) x k : 4x4 on {r,r'} x {c,c'}
//   (src row item) x (inert column)    x k : alpha single on {r,r'} x {c}
//   (inert row)    x (src column item) x k : beta single on {r} x {c,c'}
// Shared memory: [tile Rn x LD doubles][beta delta, beta gauge words: LTB each][alpha delta, alpha gauge words:
// LTA each][brick items, double buffered: 2 x (LTA + LTB)].  All table reads inside the loops are LDS: the tables
// of a CTA's classes are staged once, the item lists of brick b+1 are fetched while brick b is computed.
template <int LOGK>
__global__ void __launch_bounds__(WIN_THREADS, 3)
win_kernel(double* __restrict__ C, int64_t NB, const WinSideDev A, const WinSideDev B, const WinProgram P, int tile_doubles) {
  constexpr int K = 1 << LOGK;
  extern __shared__ double tile[];
  int* const sdB = reinterpret_cast<int*>(tile + tile_doubles);
  uint32_t* const sgB = reinterpret_cast<uint32_t*>(sdB + B.LT);
  int* const sdA = reinterpret_cast<int*>(sgB + B.LT);
  uint32_t* const sgA = reinterpret_cast<uint32_t*>(sdA + A.LT);
  uint32_t* const sitems = sgA + A.LT;
  const int ITS = A.LT + B.LT;   // words per item buffer: [alpha items LTA][beta items LTB]
  const int2 ga = __ldg(A.groups + blockIdx.y), gb = __ldg(B.groups + blockIdx.x);
  const int clsA = ga.y, clsB = gb.y & 0xffff, kcnt = gb.y >> 16;
  const int Rn = __ldg(A.cnt + clsA), Wn = __ldg(B.cnt + clsB);
  const int LD = Wn << LOGK;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* const base = C + gb.x;

  int lp = P.pair[0];
  int2 nA = __ldg(A.itemcnt + lp * A.ncls + clsA), nB = __ldg(B.itemcnt + lp * B.ncls + clsB);
  for (int t = threadIdx.x; t < Wn; t += WIN_THREADS) {
    sdB[t] = __ldg(B.delta + clsB * B.LT + t);
    sgB[t] = __ldg(B.gbits + clsB * B.LT + t);
  }
  for (int t = threadIdx.x; t < Rn; t += WIN_THREADS) {
    sdA[t] = __ldg(A.delta + clsA * A.LT + t);
    sgA[t] = __ldg(A.gbits + clsA * A.LT + t);
  }
  for (int t = threadIdx.x; t < ITS; t += WIN_THREADS)
    sitems[t] = (t < A.LT) ? __ldg(A.items + (size_t)(lp * A.ncls + clsA) * A.LT + t)
                           : __ldg(B.items + (size_t)(lp * B.ncls + clsB) * B.LT + (t - A.LT));
  __syncthreads();

  // the whole tile is requested at once with 8-byte async copies (no registers in between: the CTA has its full
  // tile in flight), then the gauge sign is applied in shared memory
  const uint32_t tb = (uint32_t)__cvta_generic_to_shared(tile);
  for (int r = warp; r < Rn; r += WIN_WARPS) {
    const double* src = base + (int64_t)(ga.x + sdA[r]) * NB;
    const uint32_t dst = tb + (uint32_t)(r * LD) * 8u;
#pragma unroll 4
    for (int x = lane; x < LD; x += 32) {
      const int j = x >> LOGK, k = x & (K - 1);
      if (LOGK == 0 || k < kcnt) cp_async8(dst + (uint32_t)x * 8u, src + sdB[j] + k);
    }
  }
  cp_async_wait_all();
  __syncthreads();
  for (int r = warp; r < Rn; r += WIN_WARPS) {
    const uint32_t wa = sgA[r];
    double* dst = tile + r * LD;
    for (int x = lane; x < LD; x += 32)
      if (__popc(wa & sgB[x >> LOGK]) & 1) dst[x] = -dst[x];
  }

  // Brick loop, column-stationary: a thread owns one column lane (column item x k) and walks down the row items,
  // so the per-item work is one LDS of the packed row offsets, four address adds and the 4x4 / 2x2 update.
  // Threads [0,TS) take the src column lanes (4x4 with src rows, beta single with inert rows), threads [TS,256)
  // the inert column lanes (alpha single with src rows); TS follows the work ratio in whole warps.
  for (int b = 0; b < P.n; ++b) {
    const uint32_t* itA = sitems + (b & 1) * ITS;
    const uint32_t* itB = itA + A.LT;
    const int nRS = nA.x, nRI = nA.y, nCS = nB.x, nCI = nB.y;
    const int NS = nCS << LOGK, NI = nCI << LOGK;
    // fetch the next brick's scalars and item lists now; they are parked in the other buffer after the loops
    uint32_t nxt[3] = {0u, 0u, 0u};
    const bool more = b + 1 < P.n;
    if (more) {
      lp = P.pair[b + 1];
      nA = __ldg(A.itemcnt + lp * A.ncls + clsA);
      nB = __ldg(B.itemcnt + lp * B.ncls + clsB);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int t = threadIdx.x + q * WIN_THREADS;
        if (t < ITS)
          nxt[q] = (t < A.LT) ? __ldg(A.items + (size_t)(lp * A.ncls + clsA) * A.LT + t)
                              : __ldg(B.items + (size_t)(lp * B.ncls + clsB) * B.LT + (t - A.LT));
      }
    }
    const int ws = NS * (2 * nRS + nRI), wi = NI * nRS;
    int TS = WIN_THREADS;
    if (wi > 0) {
      TS = ws > 0 ? ((int)((float)(WIN_THREADS / 32) * (float)ws / (float)(ws + wi) + 0.5f)) * 32 : 0;
      TS = ws > 0 ? min(max(TS, 32), WIN_THREADS - 32) : 0;
    }
    __syncthreads();
    if ((int)threadIdx.x < TS) {
      if (NS > 0) {
        const int CS = min(NS, TS);
        const int RG = TS / CS;
        const int rg = (int)(((float)threadIdx.x + 0.5f) / (float)CS);
        const int cl0 = threadIdx.x - rg * CS;
        if (rg < RG) {
          double m[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) m[e] = P.br[b].m[e];
          const double cbt = P.br[b].cb, sbt = P.br[b].sb;
          for (int cl = cl0; cl < NS; cl += CS) {
            const int k = cl & (K - 1);
            if (LOGK != 0 && k >= kcnt) continue;
            const uint32_t cw = itB[cl >> LOGK];
            const int c = (IT_J(cw) << LOGK) + k, cp = (IT_JP(cw) << LOGK) + k;
            int ri = rg;
            for (; ri + RG < nRS; ri += 2 * RG) {   // two src row items per step, loads before stores
              const uint32_t w0 = itA[ri], w1 = itA[ri + RG];
              const int a0 = IT_J(w0) * LD, a1 = IT_JP(w0) * LD, b0 = IT_J(w1) * LD, b1 = IT_JP(w1) * LD;
              const double y0 = tile[a0 + c], y1 = tile[a0 + cp], y2 = tile[a1 + c], y3 = tile[a1 + cp];
              const double z0 = tile[b0 + c], z1 = tile[b0 + cp], z2 = tile[b1 + c], z3 = tile[b1 + cp];
              tile[a0 + c] = m[0] * y0 + m[1] * y1 + m[2] * y2 + m[3] * y3;
              tile[a0 + cp] = m[4] * y0 + m[5] * y1 + m[6] * y2 + m[7] * y3;
              tile[a1 + c] = m[8] * y0 + m[9] * y1 + m[10] * y2 + m[11] * y3;
              tile[a1 + cp] = m[12] * y0 + m[13] * y1 + m[14] * y2 + m[15] * y3;
              tile[b0 + c] = m[0] * z0 + m[1] * z1 + m[2] * z2 + m[3] * z3;
              tile[b0 + cp] = m[4] * z0 + m[5] * z1 + m[6] * z2 + m[7] * z3;
              tile[b1 + c] = m[8] * z0 + m[9] * z1 + m[10] * z2 + m[11] * z3;
              tile[b1 + cp] = m[12] * z0 + m[13] * z1 + m[14] * z2 + m[15] * z3;
            }
            if (ri < nRS) {
              const uint32_t w0 = itA[ri];
              const int a0 = IT_J(w0) * LD, a1 = IT_JP(w0) * LD;
              const double y0 = tile[a0 + c], y1 = tile[a0 + cp], y2 = tile[a1 + c], y3 = tile[a1 + cp];
              tile[a0 + c] = m[0] * y0 + m[1] * y1 + m[2] * y2 + m[3] * y3;
              tile[a0 + cp] = m[4] * y0 + m[5] * y1 + m[6] * y2 + m[7] * y3;
              tile[a1 + c] = m[8] * y0 + m[9] * y1 + m[10] * y2 + m[11] * y3;
              tile[a1 + cp] = m[12] * y0 + m[13] * y1 + m[14] * y2 + m[15] * y3;
            }
            // inert rows: beta single on (c, c')
            ri = rg;
            for (; ri + RG < nRI; ri += 2 * RG) {
              const int a0 = IT_J(itA[nRS + ri]) * LD, b0 = IT_J(itA[nRS + ri + RG]) * LD;
              const double y0 = tile[a0 + c], y1 = tile[a0 + cp], z0 = tile[b0 + c], z1 = tile[b0 + cp];
              tile[a0 + c] = cbt * y0 - sbt * y1;
              tile[a0 + cp] = cbt * y1 + sbt * y0;
              tile[b0 + c] = cbt * z0 - sbt * z1;
              tile[b0 + cp] = cbt * z1 + sbt * z0;
            }
            if (ri < nRI) {
              const int a0 = IT_J(itA[nRS + ri]) * LD;
              const double y0 = tile[a0 + c], y1 = tile[a0 + cp];
              tile[a0 + c] = cbt * y0 - sbt * y1;
              tile[a0 + cp] = cbt * y1 + sbt * y0;
            }
          }
        }
      }
    } else if (NI > 0 && nRS > 0) {
      const int TI = WIN_THREADS - TS, t = threadIdx.x - TS;
      const int CS = min(NI, TI);
      const int RG = TI / CS;
      const int rg = (int)(((float)t + 0.5f) / (float)CS);
      const int cl0 = t - rg * CS;
      if (rg < RG) {
        const double ca = P.br[b].ca, sa = P.br[b].sa;
        for (int cl = cl0; cl < NI; cl += CS) {
          const int k = cl & (K - 1);
          if (LOGK != 0 && k >= kcnt) continue;
          const int c = (IT_J(itB[nCS + (cl >> LOGK)]) << LOGK) + k;
          int ri = rg;
          for (; ri + RG < nRS; ri += 2 * RG) {
            const uint32_t w0 = itA[ri], w1 = itA[ri + RG];
            const int a0 = IT_J(w0) * LD + c, a1 = IT_JP(w0) * LD + c, b0 = IT_J(w1) * LD + c, b1 = IT_JP(w1) * LD + c;
            const double y0 = tile[a0], y1 = tile[a1], z0 = tile[b0], z1 = tile[b1];
            tile[a0] = ca * y0 - sa * y1;
            tile[a1] = ca * y1 + sa * y0;
            tile[b0] = ca * z0 - sa * z1;
            tile[b1] = ca * z1 + sa * z0;
          }
          if (ri < nRS) {
            const uint32_t w0 = itA[ri];
            const int a0 = IT_J(w0) * LD + c, a1 = IT_JP(w0) * LD + c;
            const double y0 = tile[a0], y1 = tile[a1];
            tile[a0] = ca * y0 - sa * y1;
            tile[a1] = ca * y1 + sa * y0;
          }
        }
      }
    }
    if (more) {
      uint32_t* dst = sitems + ((b + 1) & 1) * ITS;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int t = threadIdx.x + q * WIN_THREADS;
        if (t < ITS) dst[t] = nxt[q];
      }
    }
  }
  __syncthreads();

  for (int r = warp; r < Rn; r += WIN_WARPS) {
    double* dstg = base + (int64_t)(ga.x + sdA[r]) * NB;
    const uint32_t wa = sgA[r];
    const double* srct = tile + r * LD;
#pragma unroll 4
    for (int x = lane; x < LD; x += 32) {
      const int j = x >> LOGK, k = x & (K - 1);
      if (LOGK == 0 || k < kcnt) stg_stream(dstg + sdB[j] + k, wflip(srct[x], __popc(wa & sgB[j]) & 1));
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

struct SideHost {
  std::vector<int2> groups;
  std::vector<int> delta, cnt;
  std::vector<uint32_t> items, gbits;
  std::vector<int2> itemcnt;
  int ncls = 0, LT = 0, max_cnt = 0;
};

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
static bool build_side(const sq_space* sp, const sq_layout* lay, int spin, int w0, int H, int K, const std::vector<int>& pairs,
                       SideHost* out) {
  const std::vector<uint32_t>& strs = spin ? sp->strB : sp->strA;
  const int ne = spin ? sp->n_beta : sp->n_alpha;
  const int64_t lo = spin ? 0 : sp->row_begin, hi = spin ? sp->NB : sp->row_end;
  if (H < 1 || H > 12 || w0 < 0 || w0 + H > sp->n_orb) return false;
  const uint32_t wmask = (1u << H) - 1u, wbits = wmask << w0, premask = (1u << w0) - 1u;
  // window parts per electron count, in combination order (a lower orbital occupied sorts first)
  std::vector<std::vector<uint32_t>> wl(H + 1);
  for (uint32_t w = 0; w <= wmask; ++w) wl[__builtin_popcount(w)].push_back(w);
  std::vector<int> pos(wmask + 1, 0);
  int LT = 1;
  for (int e = 0; e <= H; ++e) {
    std::sort(wl[e].begin(), wl[e].end(), [](uint32_t a, uint32_t b) {
      const uint32_t d = a ^ b;
      if (!d) return false;
      return (a & (d & (0u - d))) != 0;
    });
    for (size_t j = 0; j < wl[e].size(); ++j) pos[wl[e][j]] = (int)j;
    LT = std::max(LT, (int)wl[e].size());
  }
  if (LT > 1023) return false;
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
  out->cnt.resize(ncls);
  out->max_cnt = 0;
  for (int c = 0; c < ncls; ++c) {
    out->cnt[c] = (int)wl[cls_ew[c]].size();
    out->max_cnt = std::max(out->max_cnt, out->cnt[c]);
  }
  // groups / chunks of K consecutive suffixes
  std::vector<int> order(gs.size());
  for (size_t i = 0; i < gs.size(); ++i) order[i] = (int)i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return gs[a].base < gs[b].base; });
  std::vector<int2> groups;
  if (spin == 0) {
    for (int i : order) groups.push_back(make_int2((int)(gs[i].base - sp->row_begin), gs[i].cls));
    std::stable_sort(groups.begin(), groups.end(), [&](const int2& a, const int2& b) { return out->cnt[a.y] > out->cnt[b.y]; });
  } else {
    size_t i = 0;
    while (i < order.size()) {
      const G& g0 = gs[order[i]];
      int c = 1;
      while (c < K && i + c < order.size()) {
        const G& g1 = gs[order[i + c]];
        if (g1.pre != g0.pre || g1.e_w != g0.e_w || g1.base != g0.base + c) break;
        ++c;
      }
      groups.push_back(make_int2((int)g0.base, g0.cls | (c << 16)));
      i += c;
    }
    std::stable_sort(groups.begin(), groups.end(), [&](const int2& a, const int2& b) {
      return out->cnt[a.y & 0xffff] * (a.y >> 16) > out->cnt[b.y & 0xffff] * (b.y >> 16);
    });
  }
  // brick work items (signs live in the gauge, see win_kernel)
  std::vector<uint32_t> items(pairs.size() * (size_t)ncls * LT, 0u);
  std::vector<int2> itemcnt(pairs.size() * (size_t)ncls, make_int2(0, 0));
  for (size_t lp = 0; lp < pairs.size(); ++lp) {
    const PairTables& pt = lay->pairs[pairs[lp]];
    const uint32_t bi = 1u << pt.i, ba = 1u << pt.a;
    if (!(bi & wbits) || !(ba & wbits)) return false;
    for (int c = 0; c < ncls; ++c) {
      const std::vector<uint32_t>& w = wl[cls_ew[c]];
      uint32_t* dst = items.data() + (lp * (size_t)ncls + c) * LT;
      int nS = 0, nI = 0;
      for (size_t j = 0; j < w.size(); ++j) {
        const uint32_t m = w[j] << w0;
        if ((m & bi) && !(m & ba)) dst[nS++] = (uint32_t)j | ((uint32_t)pos[(m ^ bi ^ ba) >> w0] << 10);
      }
      for (size_t j = 0; j < w.size(); ++j) {
        const uint32_t m = w[j] << w0;
        if (((m & bi) != 0) == ((m & ba) != 0)) dst[nS + nI++] = (uint32_t)j;
      }
      itemcnt[lp * (size_t)ncls + c] = make_int2(nS, nI);
    }
  }
  std::vector<uint32_t> gbits((size_t)ncls * LT, 0u);
  for (int c = 0; c < ncls; ++c) {
    const std::vector<uint32_t>& w = wl[cls_ew[c]];
    for (size_t j = 0; j < w.size(); ++j) gbits[(size_t)c * LT + j] = spin ? gauge_beta_word(w[j] << w0) : (w[j] << w0);
  }
  out->gbits.swap(gbits);
  out->groups.swap(groups);
  out->delta.swap(delta);
  out->items.swap(items);
  out->itemcnt.swap(itemcnt);
  out->ncls = ncls;
  out->LT = LT;
  return true;
}

static int upload_side(const SideHost& h, int w0, int H, WinSide* s) {
  s->w0 = w0;
  s->H = H;
  s->ncls = h.ncls;
  s->LT = h.LT;
  s->n_groups = (int)h.groups.size();
  s->max_cnt = h.max_cnt;
  SQ_CHECK(win_upload(&s->d_groups, h.groups));
  SQ_CHECK(win_upload(&s->d_delta, h.delta));
  SQ_CHECK(win_upload(&s->d_gbits, h.gbits));
  SQ_CHECK(win_upload(&s->d_cnt, h.cnt));
  SQ_CHECK(win_upload(&s->d_items, h.items));
  SQ_CHECK(win_upload(&s->d_itemcnt, h.itemcnt));
  return SQ_OK;
}

static void free_side(WinSide* s) {
  cudaFree(s->d_groups);
  cudaFree(s->d_delta);
  cudaFree(s->d_gbits);
  cudaFree(s->d_cnt);
  cudaFree(s->d_items);
  cudaFree(s->d_itemcnt);
}

void sq_free_win_tables(WinTables* wt) {
  if (!wt) return;
  free_side(&wt->A);
  free_side(&wt->B);
  delete wt;
}

// pairs of the layout that the window kernel can take: both orbitals inside both windows, uniform gauge sign,
// no row pair spanning two devices
bool sq_win_pair_ok(const sq_layout* lay, int pair, int a0, int Ha, int b0, int Hb) {
  const PairTables& pt = lay->pairs[pair];
  const int lo = std::min(pt.i, pt.a), hi = std::max(pt.i, pt.a);
  // nearest-neighbour pairs only: their signs vanish in the window gauge
  return hi == lo + 1 && lo >= a0 && hi < a0 + Ha && lo >= b0 && hi < b0 + Hb && !pt.cross_global && pt.n_cross_items == 0;
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

// Tables for alpha window [a0,a0+Ha) and beta window [b0,b0+Hb) with K consecutive beta suffixes per CTA;
// cached on the layout.  (*out)->ok is false when the combination is unusable.
int sq_get_win(sq_space* sp, sq_layout* lay, int a0, int Ha, int b0, int Hb, int K, const WinTables** out) {
  const std::array<int, 5> key = {a0, Ha, b0, Hb, K};
  auto it = lay->wins.find(key);
  if (it != lay->wins.end()) {
    *out = it->second;
    return SQ_OK;
  }
  WinTables* wt = new WinTables();
  lay->wins[key] = wt;
  *out = wt;
  wt->K = K;
  wt->logK = 0;
  while ((1 << wt->logK) < K) ++wt->logK;
  if ((1 << wt->logK) != K || wt->logK > 4) return SQ_OK;
  if (sp->world > 1) {
    int k = 0;
    while ((1 << k) < sp->world) ++k;
    if (a0 < k) return SQ_OK;   // alpha prefixes would straddle devices
  }
  wt->pair_local.assign(lay->pairs.size(), -1);
  std::vector<int> pairs;
  for (size_t p = 0; p < lay->pairs.size(); ++p)
    if (sq_win_pair_ok(lay, (int)p, a0, Ha, b0, Hb)) {
      wt->pair_local[p] = (int)pairs.size();
      pairs.push_back((int)p);
    }
  if (pairs.empty()) return SQ_OK;
  // constant signs of the three generators of every pair in the window gauge
  wt->eps.assign(3 * pairs.size(), 1);
  for (size_t lp = 0; lp < pairs.size(); ++lp)
    if (!gauge_signs(sp, lay->pairs[pairs[lp]], a0, Ha, b0, Hb, &wt->eps[3 * lp])) return SQ_OK;
  SideHost hA, hB;
  if (!build_side(sp, lay, 0, a0, Ha, 1, pairs, &hA)) return SQ_OK;
  if (!build_side(sp, lay, 1, b0, Hb, K, pairs, &hB)) return SQ_OK;
  if (hA.groups.size() > 65535 || hB.groups.empty() || hA.groups.empty()) return SQ_OK;
  if (hA.LT + hB.LT > 3 * WIN_THREADS) return SQ_OK;   // item lists are fetched 3 words per thread
  wt->tile_doubles = hA.max_cnt * hB.max_cnt * K;
  wt->smem = sizeof(double) * (size_t)wt->tile_doubles + 4 * (size_t)(2 * hB.LT + 2 * hA.LT + 2 * (hA.LT + hB.LT));
  if (wt->smem > 220 * 1024) return SQ_OK;
  wt->touched = sp->local_len();
  if (sp->device >= 0) {
    SQ_CUDA(cudaSetDevice(sp->device));
    SQ_CHECK(upload_side(hA, a0, Ha, &wt->A));
    SQ_CHECK(upload_side(hB, b0, Hb, &wt->B));
  } else {   // host-only space: plan / partition logic without device tables
    wt->A.w0 = a0; wt->A.H = Ha; wt->A.n_groups = (int)hA.groups.size();
    wt->B.w0 = b0; wt->B.H = Hb; wt->B.n_groups = (int)hB.groups.size();
  }
  wt->ok = true;
  return SQ_OK;
}

static WinSideDev side_dev(const WinSide& s) {
  WinSideDev d;
  d.groups = s.d_groups;
  d.delta = s.d_delta;
  d.gbits = s.d_gbits;
  d.cnt = s.d_cnt;
  d.items = s.d_items;
  d.itemcnt = s.d_itemcnt;
  d.LT = s.LT;
  d.ncls = s.ncls;
  return d;
}

template <int LOGK>
static cudaError_t launch_win_k(dim3 grid, size_t smem, cudaStream_t st, double* state, int64_t NB, const WinSideDev& A,
                                const WinSideDev& B, const WinProgram& P, int tile_doubles) {
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(win_kernel<LOGK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = smem;
  }
  win_kernel<LOGK><<<grid, WIN_THREADS, smem, st>>>(state, NB, A, B, P, tile_doubles);
  return cudaGetLastError();
}

// bricks[k] = (layout pair index, rotation steps of the fused program); every pair must be usable in `wt`
int sq_launch_win(sq_space* sp, const WinTables& wt, const int* pair_idx, const TileStep* const* steps, const int* n_steps,
                  int n_bricks, double* state, cudaStream_t st) {
  if (!wt.ok || n_bricks < 1 || n_bricks > SQ_WIN_MAX_BRICKS) {
    sq_set_error("window launch with %d bricks (max %d) or without tables", n_bricks, SQ_WIN_MAX_BRICKS);
    return SQ_ERR_INVALID;
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
  const dim3 grid((unsigned)wt.B.n_groups, (unsigned)wt.A.n_groups);
  const WinSideDev A = side_dev(wt.A), B = side_dev(wt.B);
  cudaError_t e = cudaSuccess;
  switch (wt.logK) {
    case 0: e = launch_win_k<0>(grid, wt.smem, st, state, sp->NB, A, B, P, wt.tile_doubles); break;
    case 1: e = launch_win_k<1>(grid, wt.smem, st, state, sp->NB, A, B, P, wt.tile_doubles); break;
    case 2: e = launch_win_k<2>(grid, wt.smem, st, state, sp->NB, A, B, P, wt.tile_doubles); break;
    case 3: e = launch_win_k<3>(grid, wt.smem, st, state, sp->NB, A, B, P, wt.tile_doubles); break;
    default: e = launch_win_k<4>(grid, wt.smem, st, state, sp->NB, A, B, P, wt.tile_doubles); break;
  }
  if (e != cudaSuccess) {
    sq_set_error("win_kernel launch failed: %s", cudaGetErrorString(e));
    return SQ_ERR_CUDA;
  }
  g_sq_launches.fetch_add(1);
  return SQ_OK;
}
