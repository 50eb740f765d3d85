/*
 * sq_oracle.c -- CPU restatement of the reference's state-vector hot loops.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product (slowquant_b200/) never does.
 *
 * Each function follows the reference file:line it cites (paths relative to the reference root,
 * slowquant/unitary_coupled_cluster/):
 *   orc_indexing            ci_spaces.py:56-116      (generate_spin_strings + get_indexing)
 *   orc_apply_serial        operator_state_algebra.py:53-136   (apply_operator_serial)
 *   orc_apply_threaded      operator_state_algebra.py:139-219  (apply_operator_threaded; OpenMP for prange)
 *   orc_electronic_energy   density_matrix.py:5-178  (RDM1, RDM2, get_electronic_energy)
 *   orc_orbital_gradient    density_matrix.py:181-230
 *
 * det2idx: the reference uses a hash map keyed by the determinant integer (ci_spaces.py:47-52).  Here the
 * same map is evaluated by de-interleaving the determinant into its alpha/beta occupation patterns and
 * looking both up in per-spin tables -- identical results (-1 = "key not in dict"), no multi-GB map.
 *
 * Pinning: checked against the reference itself (imported from /root/reference in the build container)
 * through the golden vectors in tests/golden/ (generator: tests/golden/make_golden.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct {
  int n, na, nb;
  int64_t NA, NB;
  uint32_t* strA; /* occupation pattern, bit (n-1-o) = orbital o: numeric value of the 0/1 list read as binary */
  uint32_t* strB;
  int32_t* rankA; /* pattern -> position in the string list, -1 if absent */
  int32_t* rankB;
} orc_space;

/* itertools.combinations(range(n), k) order (ci_spaces.py:70-73) */
static int64_t gen_strings(int n, int k, uint32_t** out) {
  int64_t cap = 16, cnt = 0;
  uint32_t* v = (uint32_t*)malloc(sizeof(uint32_t) * cap);
  if (k < 0 || k > n) { *out = v; return 0; }
  int c[64];
  for (int i = 0; i < k; ++i) c[i] = i;
  for (;;) {
    uint32_t m = 0;
    for (int i = 0; i < k; ++i) m |= 1u << (n - 1 - c[i]);
    if (cnt == cap) { cap *= 2; v = (uint32_t*)realloc(v, sizeof(uint32_t) * cap); }
    v[cnt++] = m;
    int i = k - 1;
    while (i >= 0 && c[i] == n - k + i) --i;
    if (i < 0) break;
    ++c[i];
    for (int j = i + 1; j < k; ++j) c[j] = c[j - 1] + 1;
  }
  *out = v;
  return cnt;
}

orc_space* orc_space_create(int n, int na, int nb) {
  if (n < 1 || n > 26) return NULL;
  orc_space* s = (orc_space*)calloc(1, sizeof(orc_space));
  s->n = n; s->na = na; s->nb = nb;
  s->NA = gen_strings(n, na, &s->strA);
  s->NB = gen_strings(n, nb, &s->strB);
  size_t nm = (size_t)1 << n;
  s->rankA = (int32_t*)malloc(sizeof(int32_t) * nm);
  s->rankB = (int32_t*)malloc(sizeof(int32_t) * nm);
  for (size_t i = 0; i < nm; ++i) { s->rankA[i] = -1; s->rankB[i] = -1; }
  for (int64_t i = 0; i < s->NA; ++i) s->rankA[s->strA[i]] = (int32_t)i;
  for (int64_t i = 0; i < s->NB; ++i) s->rankB[s->strB[i]] = (int32_t)i;
  return s;
}

void orc_space_destroy(orc_space* s) {
  if (!s) return;
  free(s->strA); free(s->strB); free(s->rankA); free(s->rankB); free(s);
}

int64_t orc_num_det(const orc_space* s) { return s->NA * s->NB; }

/* det_str = a0 b0 a1 b1 ... read as a binary number (ci_spaces.py:99-104) */
static inline int64_t interleave(uint32_t a, uint32_t b, int n) {
  int64_t d = 0;
  for (int o = 0; o < n; ++o) {
    int sh = n - 1 - o;
    d = (d << 2) | (int64_t)((((a >> sh) & 1u) << 1) | ((b >> sh) & 1u));
  }
  return d;
}

/* ci_spaces.py:93-107: alpha strings outer loop, beta strings inner loop */
void orc_indexing(const orc_space* s, int64_t* idx2det) {
  int64_t idx = 0;
  for (int64_t ia = 0; ia < s->NA; ++ia)
    for (int64_t ib = 0; ib < s->NB; ++ib) idx2det[idx++] = interleave(s->strA[ia], s->strB[ib], s->n);
}

static inline int64_t det2idx(const orc_space* s, int64_t det) {
  const int n = s->n;
  if (det < 0 || (n < 32 && (det >> (2 * n)) != 0)) return -1;
  uint32_t a = 0, b = 0;
#ifdef __BMI2__
  a = (uint32_t)__builtin_ia32_pext_di((unsigned long long)det, 0xAAAAAAAAAAAAAAAAull);
  b = (uint32_t)__builtin_ia32_pext_di((unsigned long long)det, 0x5555555555555555ull);
#else
  for (int o = 0; o < n; ++o) {
    int sh = 2 * (n - 1 - o);
    a = (a << 1) | (uint32_t)((det >> (sh + 1)) & 1);
    b = (b << 1) | (uint32_t)((det >> sh) & 1);
  }
#endif
  int32_t ra = s->rankA[a], rb = s->rankB[b];
  if (ra < 0 || rb < 0) return -1;
  return (int64_t)ra * s->NB + rb;
}

int64_t orc_det2idx(const orc_space* s, int64_t det) { return det2idx(s, det); }

/* Brian Kernighan popcount, as operator_state_algebra.py:33-50 */
static inline int bitcount(int64_t x) {
  /* same value as the Kernighan loop of the reference for x >= 0 */
  return __builtin_popcountll((unsigned long long)x);
}

/* parity_check[k] = the k most significant of the 2n determinant bits (osa.py:518-522) */
static void make_parity_check(int n, int64_t* pc) {
  int64_t num = 0;
  pc[0] = 0;
  for (int i = 2 * n - 1; i >= 0; --i) {
    num += (int64_t)1 << i;
    pc[2 * n - i] = num;
  }
}

/* apply_operator_serial (osa.py:53-136).  Returns 0, or 1 when a determinant left the space with
 * do_unsafe == 0 (the reference raises KeyError there). */
int orc_apply_serial(const orc_space* s, const int64_t* idx2det, const double* state, const int64_t* a_string,
                     int n_a, const int64_t* create_screen, int n_cs, const int64_t* anni_idx, int n_an,
                     int do_unsafe, double* tmp_state, double factor) {
  const int n = s->n;
  const int m1 = 2 * n - 1;
  int64_t pc[2 * 32 + 2];
  make_parity_check(n, pc);
  int64_t anni_mask = 0, create_mask = 0;
  for (int k = 0; k < n_an; ++k) anni_mask |= (int64_t)1 << (m1 - anni_idx[k]);
  for (int k = 0; k < n_cs; ++k) create_mask |= (int64_t)1 << (m1 - create_screen[k]);
  const int64_t nd = s->NA * s->NB;
  for (int64_t i = 0; i < nd; ++i) {
    int64_t det = idx2det[i];
    if ((det & anni_mask) != anni_mask) continue;
    if ((det & create_mask) != 0) continue;
    const double state_i = state[i];
    if (fabs(state_i) < 1e-28) continue;
    int phase = 0;
    for (int k = 0; k < n_a; ++k) {
      det ^= (int64_t)1 << (m1 - a_string[k]);
      phase += bitcount(det & pc[a_string[k]]);
    }
    const int64_t j = det2idx(s, det);
    if (j < 0) {
      if (do_unsafe) continue;
      return 1;
    }
    const double sign = 1.0 - 2.0 * (phase & 1);
    tmp_state[j] += sign * factor * state_i;
  }
  return 0;
}

/* apply_operator_threaded (osa.py:139-219): gather over target determinants; prange -> OpenMP */
int orc_apply_threaded(const orc_space* s, const int64_t* idx2det, const double* state, const int64_t* a_string,
                       int n_a, const int64_t* create_idx, int n_c, const int64_t* anni_screen, int n_as,
                       int do_unsafe, double* tmp_state, double factor) {
  const int n = s->n;
  const int m1 = 2 * n - 1;
  int64_t pc[2 * 32 + 2];
  make_parity_check(n, pc);
  int64_t create_mask = 0, anni_mask = 0;
  for (int k = 0; k < n_c; ++k) create_mask |= (int64_t)1 << (m1 - create_idx[k]);
  for (int k = 0; k < n_as; ++k) anni_mask |= (int64_t)1 << (m1 - anni_screen[k]);
  const int64_t nd = s->NA * s->NB;
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int64_t i = 0; i < nd; ++i) {
    int64_t det = idx2det[i];
    if ((det & create_mask) != create_mask) continue;
    if ((det & anni_mask) != 0) continue;
    int phase = 0;
    for (int k = 0; k < n_a; ++k) {
      det ^= (int64_t)1 << (m1 - a_string[k]);
      phase += bitcount(det & pc[a_string[k]]);
    }
    const int64_t j = det2idx(s, det);
    if (j < 0) {
      if (!do_unsafe) bad = 1;
      continue;
    }
    const double sign = 1.0 - 2.0 * (phase & 1);
    tmp_state[i] += sign * factor * state[j];
  }
  return bad;
}

/* ---- density_matrix.py ---------------------------------------------------------------------- */
/* RDM1 (density_matrix.py:5-43) */
static double RDM1(int p, int q, int nI, int nA, const double* rdm1) {
  const int virt = nI + nA;
  if (p >= virt || q >= virt) return 0.0;
  if (p >= nI && q >= nI) return rdm1[(p - nI) * nA + (q - nI)];
  if (p < nI && q < nI) return (p == q) ? 2.0 : 0.0;
  return 0.0;
}

/* RDM2 (density_matrix.py:46-136) */
static double RDM2(int p, int q, int r, int s, int nI, int nA, const double* rdm1, const double* rdm2) {
  const int virt = nI + nA;
  if (p >= virt || q >= virt || r >= virt || s >= virt) return 0.0;
  const int ap = p >= nI, aq = q >= nI, ar = r >= nI, as = s >= nI;
  if (ap && aq && ar && as)
    return rdm2[(((size_t)(p - nI) * nA + (q - nI)) * nA + (r - nI)) * nA + (s - nI)];
  if (!ap && aq && ar && !as) return (p == s) ? -rdm1[(q - nI) * nA + (r - nI)] : 0.0;
  if (ap && !aq && !ar && as) return (q == r) ? -rdm1[(p - nI) * nA + (s - nI)] : 0.0;
  if (ap && aq && !ar && !as) return (r == s) ? 2.0 * rdm1[(p - nI) * nA + (q - nI)] : 0.0;
  if (!ap && !aq && ar && as) return (p == q) ? 2.0 * rdm1[(r - nI) * nA + (s - nI)] : 0.0;
  if (!ap && !aq && !ar && !as) {
    double val = 0.0;
    if (p == q && r == s) val += 4.0;
    if (q == r && p == s) val -= 2.0;
    return val;
  }
  return 0.0;
}

/* get_electronic_energy (density_matrix.py:139-178); h [N][N], g [N][N][N][N] */
double orc_electronic_energy(const double* h, const double* g, int N, int nI, int nA, const double* rdm1,
                             const double* rdm2) {
  double energy = 0.0;
  const int M = nI + nA;
  for (int p = 0; p < M; ++p)
    for (int q = 0; q < M; ++q) energy += h[p * N + q] * RDM1(p, q, nI, nA, rdm1);
  for (int p = 0; p < M; ++p)
    for (int q = 0; q < M; ++q)
      for (int r = 0; r < M; ++r)
        for (int s = 0; s < M; ++s)
          energy += 1.0 / 2.0 * g[(((size_t)p * N + q) * N + r) * N + s] * RDM2(p, q, r, s, nI, nA, rdm1, rdm2);
  return energy;
}

/* get_orbital_gradient (density_matrix.py:181-230); kappa_idx [K][2] */
void orc_orbital_gradient(const double* h, const double* g, int N, const int64_t* kappa_idx, int K, int nI, int nA,
                          const double* rdm1, const double* rdm2, double* gradient) {
  const int M = nI + nA;
#define G4(a, b, c, d) g[(((size_t)(a) * N + (b)) * N + (c)) * N + (d)]
  for (int idx = 0; idx < K; ++idx) {
    const int m = (int)kappa_idx[2 * idx], n = (int)kappa_idx[2 * idx + 1];
    double acc = 0.0;
    for (int p = 0; p < M; ++p) {
      acc += 2.0 * h[n * N + p] * RDM1(m, p, nI, nA, rdm1);
      acc -= 2.0 * h[p * N + m] * RDM1(p, n, nI, nA, rdm1);
    }
    for (int p = 0; p < M; ++p)
      for (int q = 0; q < M; ++q)
        for (int r = 0; r < M; ++r) {
          acc += G4(n, p, q, r) * RDM2(m, p, q, r, nI, nA, rdm1, rdm2);
          acc -= G4(p, m, q, r) * RDM2(p, n, q, r, nI, nA, rdm1, rdm2);
          acc -= G4(m, p, q, r) * RDM2(n, p, q, r, nI, nA, rdm1, rdm2);
          acc += G4(p, n, q, r) * RDM2(p, m, q, r, nI, nA, rdm1, rdm2);
        }
    gradient[idx] = acc;
  }
#undef G4
}
