// CI space: alpha/beta string lists, ranking tables, determinant <-> index maps, and the closed-form
// action of a normal-ordered ladder string on a determinant.
//
// Replaces get_indexing (reference ci_spaces.py:76-116): instead of the idx2det array and the det2idx
// hash map the space is the product of two string lists; idx = Ia*NB + Ib reproduces the reference's
// ordering exactly (alpha outer loop, beta inner loop, both in itertools.combinations order).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "sqsv_internal.h"

std::atomic<int64_t> g_sq_launches{0};
static thread_local char g_err[1024] = "";

void sq_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* sq_last_error(void) { return g_err; }
extern "C" int sq_version(void) { return 100; }
extern "C" int64_t sq_launch_count(void) { return g_sq_launches.load(); }

// Enumerate k-subsets of {0..n-1} in lexicographic order of the sorted index tuple, which is the
// order itertools.combinations(range(n), k) yields (ci_spaces.py:56-73).
static void enumerate_strings(int n, int k, std::vector<uint32_t>& out) {
  out.clear();
  if (k < 0 || k > n) return;
  std::vector<int> c(k);
  for (int i = 0; i < k; ++i) c[i] = i;
  while (true) {
    uint32_t m = 0;
    for (int i = 0; i < k; ++i) m |= (1u << c[i]);
    out.push_back(m);
    int i = k - 1;
    while (i >= 0 && c[i] == n - k + i) --i;
    if (i < 0) break;
    ++c[i];
    for (int j = i + 1; j < k; ++j) c[j] = c[j - 1] + 1;
  }
}

int sq_rank_mask(const sq_space* sp, int spin, uint32_t mask) {
  const std::vector<int32_t>& r = spin ? sp->rankB : sp->rankA;
  if (mask >= r.size()) return -1;
  return r[mask];
}

static int space_create_impl(int n_orb, int n_alpha, int n_beta, int device, int64_t row_begin, int64_t row_end,
                             uint32_t alpha_cmask, uint32_t alpha_cpat, sq_space** out);

extern "C" int sq_space_create(int n_orb, int n_alpha, int n_beta, int device, int64_t row_begin,
                               int64_t row_end, sq_space** out) {
  return space_create_impl(n_orb, n_alpha, n_beta, device, row_begin, row_end, 0u, 0u, out);
}

// A space whose alpha list holds only the strings with (mask & alpha_cmask) == alpha_cpat, in the order they have in the full
// itertools.combinations list (the second layout of a re-sharded vector: rows grouped by the occupation of the LAST log2(G)
// orbitals).  All rows are local; operators that move an alpha electron on a constrained orbital cannot run in this space.
extern "C" int sq_space_create_constrained(int n_orb, int n_alpha, int n_beta, int device, uint32_t alpha_cmask,
                                           uint32_t alpha_cpat, sq_space** out) {
  if (n_orb >= 1 && n_orb < 32 && ((alpha_cmask >> n_orb) != 0u || (alpha_cpat & ~alpha_cmask) != 0u)) {
    sq_set_error("sq_space_create_constrained: mask 0x%x / pattern 0x%x do not fit %d orbitals", alpha_cmask, alpha_cpat, n_orb);
    if (out) *out = nullptr;
    return SQ_ERR_INVALID;
  }
  return space_create_impl(n_orb, n_alpha, n_beta, device, 0, -1, alpha_cmask, alpha_cpat, out);
}

static int space_create_impl(int n_orb, int n_alpha, int n_beta, int device, int64_t row_begin, int64_t row_end,
                             uint32_t alpha_cmask, uint32_t alpha_cpat, sq_space** out) {
  if (!out) return SQ_ERR_INVALID;
  *out = nullptr;
  if (n_orb < 1 || n_orb > 26 || n_alpha < 0 || n_beta < 0 || n_alpha > n_orb || n_beta > n_orb) {
    sq_set_error("sq_space_create: need 1 <= n_orb <= 26 and 0 <= n_alpha,n_beta <= n_orb (got %d,%d,%d)",
                 n_orb, n_alpha, n_beta);
    return SQ_ERR_INVALID;
  }
  sq_space* sp = new sq_space();
  sp->n_orb = n_orb;
  sp->n_alpha = n_alpha;
  sp->n_beta = n_beta;
  sp->device = device;
  memset(sp->binom, 0, sizeof(sp->binom));
  for (int i = 0; i <= SQ_MAX_ORB + 1; ++i) {
    sp->binom[i][0] = 1;
    for (int j = 1; j <= i; ++j)
      sp->binom[i][j] = sp->binom[i - 1][j - 1] + (j <= i - 1 ? sp->binom[i - 1][j] : 0);
  }
  enumerate_strings(n_orb, n_alpha, sp->strA);
  enumerate_strings(n_orb, n_beta, sp->strB);
  sp->alpha_cmask = alpha_cmask;
  sp->alpha_cpat = alpha_cpat;
  if (alpha_cmask) {
    std::vector<uint32_t> keep;
    for (uint32_t m : sp->strA)
      if ((m & alpha_cmask) == alpha_cpat) keep.push_back(m);
    sp->strA.swap(keep);
  }
  sp->NA = (int64_t)sp->strA.size();
  sp->NB = (int64_t)sp->strB.size();
  sp->ndet = sp->NA * sp->NB;
  if (row_end < 0) row_end = sp->NA;
  if (row_begin < 0 || row_begin > row_end || row_end > sp->NA) {
    sq_set_error("sq_space_create: bad row range [%lld,%lld) for %lld alpha strings", (long long)row_begin,
                 (long long)row_end, (long long)sp->NA);
    delete sp;
    return SQ_ERR_INVALID;
  }
  sp->row_begin = row_begin;
  sp->row_end = row_end;
  size_t nmask = (size_t)1 << n_orb;
  sp->rankA.assign(nmask, -1);
  sp->rankB.assign(nmask, -1);
  for (int64_t i = 0; i < sp->NA; ++i) sp->rankA[sp->strA[i]] = (int32_t)i;
  for (int64_t i = 0; i < sp->NB; ++i) sp->rankB[sp->strB[i]] = (int32_t)i;

  if (device < 0) {  // host-only space: integer tables for CPU-side checks, no device buffers
    *out = sp;
    return SQ_OK;
  }
  if (sp->row_end - sp->row_begin > (int64_t)65535 * 8) {
    // the brick / generic kernels put chunks of 8 rows into gridDim.y (limit 65535)
    sq_set_error("sq_space_create: %lld local alpha rows exceed the 524280 rows one device can take: shard the vector by alpha string "
                 "(row_begin / row_end)", (long long)(sp->row_end - sp->row_begin));
    delete sp;
    return SQ_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    sq_set_error("sq_space_create: cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e));
    delete sp;
    return SQ_ERR_CUDA;
  }
#define SP_CUDA(call)                                                            \
  do {                                                                           \
    cudaError_t e_ = (call);                                                     \
    if (e_ != cudaSuccess) {                                                     \
      sq_set_error("sq_space_create: %s -> %s", #call, cudaGetErrorString(e_)); \
      sq_space_destroy(sp);                                                      \
      return SQ_ERR_CUDA;                                                        \
    }                                                                            \
  } while (0)
  SP_CUDA(cudaMalloc(&sp->d_strA, sizeof(uint32_t) * (sp->NA > 0 ? sp->NA : 1)));
  SP_CUDA(cudaMalloc(&sp->d_strB, sizeof(uint32_t) * sp->NB));
  SP_CUDA(cudaMalloc(&sp->d_rankA, sizeof(int32_t) * nmask));
  SP_CUDA(cudaMalloc(&sp->d_rankB, sizeof(int32_t) * nmask));
  SP_CUDA(cudaMemcpy(sp->d_strA, sp->strA.data(), sizeof(uint32_t) * sp->NA, cudaMemcpyHostToDevice));
  SP_CUDA(cudaMemcpy(sp->d_strB, sp->strB.data(), sizeof(uint32_t) * sp->NB, cudaMemcpyHostToDevice));
  SP_CUDA(cudaMemcpy(sp->d_rankA, sp->rankA.data(), sizeof(int32_t) * nmask, cudaMemcpyHostToDevice));
  SP_CUDA(cudaMemcpy(sp->d_rankB, sp->rankB.data(), sizeof(int32_t) * nmask, cudaMemcpyHostToDevice));
  SP_CUDA(cudaMallocHost(&sp->h_pinned, sizeof(double) * 4096));
#undef SP_CUDA
  *out = sp;
  return SQ_OK;
}

extern "C" int sq_space_destroy(sq_space* sp) {
  if (!sp) return SQ_OK;
  if (sp->device < 0) {
    delete sp;
    return SQ_OK;
  }
  cudaSetDevice(sp->device);
  sq_hamiltonian_release(sp);
  cudaFree(sp->d_strA);
  cudaFree(sp->d_gwordB);
  cudaFree(sp->d_strB);
  cudaFree(sp->d_rankA);
  cudaFree(sp->d_rankB);
  cudaFree(sp->d_partial);
  cudaFree(sp->d_peer_tab);
  for (int i = 0; i < 3; ++i) cudaFree(sp->d_work[i]);
  if (sp->h_pinned) cudaFreeHost(sp->h_pinned);
  delete sp;
  return SQ_OK;
}

extern "C" int64_t sq_space_num_det(const sq_space* sp) { return sp ? sp->ndet : -1; }
extern "C" int64_t sq_space_num_strings(const sq_space* sp, int spin) {
  return sp ? (spin ? sp->NB : sp->NA) : -1;
}
extern "C" int64_t sq_space_local_rows(const sq_space* sp) { return sp ? sp->row_end - sp->row_begin : -1; }

extern "C" int sq_space_export_strings(const sq_space* sp, int spin, uint32_t* out_host) {
  if (!sp || !out_host) return SQ_ERR_INVALID;
  const std::vector<uint32_t>& s = spin ? sp->strB : sp->strA;
  memcpy(out_host, s.data(), sizeof(uint32_t) * s.size());
  return SQ_OK;
}

// spread the low n bits of an occupation mask so that orbital o lands on bit 2*(n-1-o)
// (the reference's determinant integer puts orbital 0 in the most significant bit pair).
static inline uint64_t spread_mask(uint32_t m, int n) {
  uint64_t r = 0;
  for (int o = 0; o < n; ++o)
    if (m & (1u << o)) r |= (uint64_t)1 << (2 * (n - 1 - o));
  return r;
}

extern "C" int sq_space_export_idx2det(const sq_space* sp, int64_t first, int64_t count, int64_t* out_host) {
  if (!sp || !out_host || first < 0 || count < 0 || first + count > sp->ndet) return SQ_ERR_INVALID;
  const int n = sp->n_orb;
  std::vector<uint64_t> sb(sp->NB);
  for (int64_t b = 0; b < sp->NB; ++b) sb[b] = spread_mask(sp->strB[b], n);
  for (int64_t k = 0; k < count; ++k) {
    int64_t idx = first + k;
    int64_t ia = idx / sp->NB, ib = idx % sp->NB;
    out_host[k] = (int64_t)((spread_mask(sp->strA[ia], n) << 1) | sb[ib]);
  }
  return SQ_OK;
}

extern "C" int sq_space_det2idx(const sq_space* sp, int64_t n, const int64_t* dets_host, int64_t* idx_host) {
  if (!sp || !dets_host || !idx_host) return SQ_ERR_INVALID;
  const int no = sp->n_orb;
  for (int64_t k = 0; k < n; ++k) {
    uint64_t d = (uint64_t)dets_host[k];
    int64_t res = -1;
    if (dets_host[k] >= 0 && (no == 32 || (d >> (2 * no)) == 0)) {
      uint32_t a = 0, b = 0;
      for (int o = 0; o < no; ++o) {
        int sh = 2 * (no - 1 - o);
        if ((d >> (sh + 1)) & 1) a |= 1u << o;
        if ((d >> sh) & 1) b |= 1u << o;
      }
      int ra = sp->rankA[a], rb = sp->rankB[b];
      if (ra >= 0 && rb >= 0) res = (int64_t)ra * sp->NB + rb;
    }
    idx_host[k] = res;
  }
  return SQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Closed-form action of one ladder string.  `ops` holds the label in order, entry = 2*spin_orb + dagger.
// The reference applies annihilators in label order, then creators in label order
// (a_string = anni_idx + create_idx, operator_state_algebra.py:612), and after each flip of
// spin-orbital k adds popcount(det & parity_check[k]) = number of occupied spin-orbitals with index < k
// (operator_state_algebra.py:128-130, parity_check built at :518-522).
// ---------------------------------------------------------------------------------------------
static inline void below_masks(int so, uint32_t* ma, uint32_t* mb) {
  // spin-orbitals with index < so, split into alpha / beta spatial-orbital masks
  int o = so >> 1;
  uint32_t lt = (o >= 32) ? 0xffffffffu : ((1u << o) - 1u);
  if (so & 1) {  // beta operator: alpha orbitals <= o, beta orbitals < o
    *ma = (o + 1 >= 32) ? 0xffffffffu : ((1u << (o + 1)) - 1u);
    *mb = lt;
  } else {       // alpha operator: alpha orbitals < o, beta orbitals < o
    *ma = lt;
    *mb = lt;
  }
}

int sq_make_string_action(const sq_space* sp, const int32_t* ops, int n_ops, StringAction* out) {
  if (n_ops < 0 || n_ops > SQ_MAX_STRING_OPS) {
    sq_set_error("operator string with %d ladder operators (max %d)", n_ops, SQ_MAX_STRING_OPS);
    return SQ_ERR_INVALID;
  }
  if (!sp || !out || (n_ops > 0 && !ops)) return SQ_ERR_INVALID;
  std::vector<int> anni, crea;
  for (int k = 0; k < n_ops; ++k) {
    int so = ops[k] >> 1;
    if (so < 0 || so >= 2 * sp->n_orb) {
      sq_set_error("spin-orbital index %d outside the active space (0..%d)", so, 2 * sp->n_orb - 1);
      return SQ_ERR_INVALID;
    }
    if (ops[k] & 1) crea.push_back(so); else anni.push_back(so);
  }
  StringAction a;
  memset(&a, 0, sizeof(a));
  uint32_t anniA = 0, anniB = 0, creaA = 0, creaB = 0;
  // a repeated index inside the annihilator (or creator) block makes the product vanish; the reference
  // never emits such a label (normal ordering drops it), so flag it as a caller error.
  for (int so : anni) {
    uint32_t bit = 1u << (so >> 1);
    uint32_t& m = (so & 1) ? anniB : anniA;
    if (m & bit) { sq_set_error("repeated annihilator %d in string", so); return SQ_ERR_INVALID; }
    m |= bit;
  }
  for (int so : crea) {
    uint32_t bit = 1u << (so >> 1);
    uint32_t& m = (so & 1) ? creaB : creaA;
    if (m & bit) { sq_set_error("repeated creator %d in string", so); return SQ_ERR_INVALID; }
    m |= bit;
  }
  // source screen (operator_state_algebra.py:112-123): annihilated orbitals occupied, created orbitals
  // that are not also annihilated empty.
  a.occA = anniA; a.occB = anniB;
  a.empA = creaA & ~anniA; a.empB = creaB & ~anniB;
  a.flipA = anniA ^ creaA; a.flipB = anniB ^ creaB;
  // target screen (operator_state_algebra.py:198-209)
  a.toccA = creaA; a.toccB = creaB;
  a.tempA = anniA & ~creaA; a.tempB = anniB & ~creaB;
  // phase: sum_k popc((src ^ F_k) & below_k)
  uint32_t FA = 0, FB = 0, PA = 0, PB = 0;
  int c0 = 0;
  std::vector<int> seq(anni);
  seq.insert(seq.end(), crea.begin(), crea.end());
  for (int so : seq) {
    if (so & 1) FB ^= 1u << (so >> 1); else FA ^= 1u << (so >> 1);
    uint32_t ma, mb;
    below_masks(so, &ma, &mb);
    PA ^= ma; PB ^= mb;
    c0 += __builtin_popcount(FA & ma) + __builtin_popcount(FB & mb);
  }
  a.parA = PA; a.parB = PB;
  a.s0 = (c0 & 1) ? -1 : 1;
  a.conserving = (__builtin_popcount(anniA) == __builtin_popcount(creaA)) &&
                 (__builtin_popcount(anniB) == __builtin_popcount(creaB));
  *out = a;
  return SQ_OK;
}

int sq_ensure_work(sq_space* sp, int which) {
  if (which < 0 || which >= 3) return SQ_ERR_INVALID;
  if (!sp->d_work[which]) {
    cudaError_t e = cudaMalloc(&sp->d_work[which], sizeof(double) * (size_t)sp->local_len());
    if (e != cudaSuccess) {
      sq_set_error("work buffer allocation of %lld doubles failed: %s", (long long)sp->local_len(),
                   cudaGetErrorString(e));
      sp->d_work[which] = nullptr;
      return SQ_ERR_NOMEM;
    }
  }
  return SQ_OK;
}

int sq_ensure_partial(sq_space* sp, int64_t n) {
  if (sp->n_partial >= n) return SQ_OK;
  if (sp->d_partial) cudaFree(sp->d_partial);
  sp->d_partial = nullptr;
  sp->n_partial = 0;
  SQ_CUDA(cudaMalloc(&sp->d_partial, sizeof(double) * (size_t)n));
  sp->n_partial = n;
  return SQ_OK;
}

// Introspection: action of one ladder string on the determinant (A,B) through the closed form used by
// every kernel.  Host-only; lets CPU tests check signs/targets against the reference's literal
// bit-flip loop (operator_state_algebra.py:118-135) without a GPU.
extern "C" int sq_debug_string_action(const sq_space* sp, const int32_t* ops, int n_ops, uint32_t A, uint32_t B,
                                      int* valid, uint32_t* tgtA, uint32_t* tgtB, int* sign) {
  if (!sp || !valid || !tgtA || !tgtB || !sign) return SQ_ERR_INVALID;
  StringAction a;
  SQ_CHECK(sq_make_string_action(sp, ops, n_ops, &a));
  *valid = ((A & a.occA) == a.occA) && ((A & a.empA) == 0u) && ((B & a.occB) == a.occB) && ((B & a.empB) == 0u);
  *tgtA = A ^ a.flipA;
  *tgtB = B ^ a.flipB;
  const int par = (__builtin_popcount(A & a.parA) + __builtin_popcount(B & a.parB)) & 1;
  *sign = par ? -a.s0 : a.s0;
  return SQ_OK;
}
