// extern "C" entry points: ansatz layout compilation, unitary-product-state application, gradient
// sweep, generic operator application.  See include/sqsv.h for the reference functions each replaces.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sqsv_internal.h"

int sq_build_quad_tables(sq_space* sp, const PairTables& p1, const PairTables& p2, int pair1, int pair2, QuadTables* qt);
void sq_free_quad_tables(QuadTables* qt);

// ---------------------------------------------------------------------------------------------
// table construction
// ---------------------------------------------------------------------------------------------
static inline int neg_of(uint32_t m, uint32_t par) { return __builtin_popcount(m & par) & 1; }

template <typename T>
static int upload(T** dptr, const std::vector<T>& v) {
  *dptr = nullptr;
  if (v.empty()) return SQ_OK;
  SQ_CUDA(cudaMalloc(dptr, sizeof(T) * v.size()));
  SQ_CUDA(cudaMemcpy(*dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return SQ_OK;
}

// Tables for the spatial orbital pair (i,a): generators
//   Ta = a+_{a,alpha} a_{i,alpha} - h.c.   Tb = a+_{a,beta} a_{i,beta} - h.c.     (sa_single; osa.py:1006-1007)
//   D  = a+_{a,alpha} a+_{a,beta} a_{i,beta} a_{i,alpha} - h.c.                   (pair double; util.py:705)
static int build_pair_tables(sq_space* sp, int i, int a, PairTables* pt) {
  if (i < 0 || a < 0 || i >= sp->n_orb || a >= sp->n_orb || i == a) {
    sq_set_error("orbital pair (%d,%d) outside the active space / degenerate", i, a);
    return SQ_ERR_INVALID;
  }
  pt->i = i;
  pt->a = a;
  if (sp->alpha_cmask & ((1u << i) | (1u << a))) {
    // constrained alpha list: partner rows do not exist in this space; the pair is never launched here
    pt->blocked = true;
    return SQ_OK;
  }
  StringAction actA, actB, actD;
  int32_t la[2] = {2 * (2 * a) + 1, 2 * (2 * i)};
  int32_t lb[2] = {2 * (2 * a + 1) + 1, 2 * (2 * i + 1)};
  // normal-ordered pair double: creators descending (2a+1, 2a), annihilators descending (2i+1, 2i), factor -1
  int32_t ld[4] = {2 * (2 * a + 1) + 1, 2 * (2 * a) + 1, 2 * (2 * i + 1), 2 * (2 * i)};
  SQ_CHECK(sq_make_string_action(sp, la, 2, &actA));
  SQ_CHECK(sq_make_string_action(sp, lb, 2, &actB));
  SQ_CHECK(sq_make_string_action(sp, ld, 4, &actD));
  const uint32_t bi = 1u << i, ba = 1u << a;
  std::vector<uint32_t> codeA(sp->NA), codeB(sp->NB);
  std::vector<int32_t> rows, crossSrc, crossTgt;
  int64_t nsrcA_local = 0, ninertA_local = 0;
  pt->cross_global = false;
  for (int64_t I = 0; I < sp->NA; ++I) {
    const uint32_t m = sp->strA[I];
    uint32_t cls = SQ_CLS_INERT, partner = 0;
    if ((m & bi) && !(m & ba)) {
      cls = SQ_CLS_SRC;
      partner = (uint32_t)sp->rankA[m ^ bi ^ ba];
    } else if (!(m & bi) && (m & ba)) {
      cls = SQ_CLS_TGT;
    }
    uint32_t code = cls;
    if (cls == SQ_CLS_SRC) {
      const int negS = ((actA.s0 < 0) ? 1 : 0) ^ neg_of(m, actA.parA);
      code |= (uint32_t)negS << 2;
      code |= (uint32_t)neg_of(m, actD.parA) << 4;
      code |= partner << 5;
    }
    code |= (uint32_t)neg_of(m, actB.parA) << 3;   // factor an alpha string gives the beta single
    codeA[I] = code;
    if (cls == SQ_CLS_SRC && sp->world > 1 && sq_row_owner(sp, I) != sq_row_owner(sp, partner)) pt->cross_global = true;
    if (I >= sp->row_begin && I < sp->row_end) {
      if (cls == SQ_CLS_SRC) {
        if ((int64_t)partner < sp->row_begin || (int64_t)partner >= sp->row_end) {
          if (sp->world <= 1) {
            sq_set_error("orbital pair (%d,%d): partner alpha row %u of row %lld is on another device and no "
                         "partition is set (sq_space_set_partition)", i, a, partner, (long long)I);
            return SQ_ERR_UNSUPPORTED;
          }
          crossSrc.push_back((int32_t)I);   // this rank owns the src row of a cross-device pair
        } else {
          rows.push_back((int32_t)I);
          ++nsrcA_local;
        }
      } else if (cls == SQ_CLS_INERT) {
        rows.push_back((int32_t)I);
        ++ninertA_local;
      } else {
        // tgt rows ride with their src row; if that one is remote this rank takes half of the tile's columns
        const int64_t src = sp->rankA[m ^ bi ^ ba];
        if (src < sp->row_begin || src >= sp->row_end) {
          if (sp->world <= 1) {
            sq_set_error("orbital pair (%d,%d): source alpha row of row %lld is on another device and no "
                         "partition is set (sq_space_set_partition)", i, a, (long long)I);
            return SQ_ERR_UNSUPPORTED;
          }
          crossTgt.push_back((int32_t)src);
        }
      }
    }
  }
  pt->n_cross_items = (int64_t)(crossSrc.size() + crossTgt.size());
  int64_t nsrcB = 0;
  for (int64_t I = 0; I < sp->NB; ++I) {
    const uint32_t m = sp->strB[I];
    uint32_t cls = SQ_CLS_INERT, partner = 0;
    if ((m & bi) && !(m & ba)) {
      cls = SQ_CLS_SRC;
      partner = (uint32_t)sp->rankB[m ^ bi ^ ba];
      ++nsrcB;
    } else if (!(m & bi) && (m & ba)) {
      cls = SQ_CLS_TGT;
    }
    uint32_t code = cls;
    if (cls == SQ_CLS_SRC) {
      const int negS = ((actB.s0 < 0) ? 1 : 0) ^ neg_of(m, actB.parB);
      code |= (uint32_t)negS << 2;
      // the pair-double constant sign (-1 from normal ordering times s0) is folded into the beta factor
      const int negD = ((-actD.s0 < 0) ? 1 : 0) ^ neg_of(m, actD.parB);
      code |= (uint32_t)negD << 4;
      code |= partner << 5;
    }
    code |= (uint32_t)neg_of(m, actA.parB) << 3;   // factor a beta string gives the alpha single
    codeB[I] = code;
  }
  pt->n_rows = (int64_t)rows.size();
  pt->n_src_rows = nsrcA_local;
  pt->n_src_cols = nsrcB;
  // amplitudes touched by a block containing sa_single: everything except inert x inert
  pt->touched = 2 * nsrcA_local * sp->NB + ninertA_local * 2 * nsrcB + pt->n_cross_items * sp->NB;
  // ---- class-homogeneous work lists (tile_kernel_v2) ----
  auto bitof = [](uint32_t code, int b) -> int { return (int)((code >> b) & 1u); };
  std::vector<int2> colItems;
  {
    std::vector<int2> src, inert;
    for (int64_t I = 0; I < sp->NB; ++I) {
      const uint32_t c = codeB[I], cls = c & 3u;
      if (cls == SQ_CLS_SRC) {
        const uint32_t ip = c >> 5;
        const int flags = bitof(c, 2) | (bitof(c, 3) << 1) | (bitof(codeB[ip], 3) << 2) | (bitof(c, 4) << 3);
        src.push_back(make_int2((int)I, (int)(ip | ((uint32_t)flags << 27))));
      } else if (cls == SQ_CLS_INERT) {
        const int flags = bitof(c, 3) << 1;
        inert.push_back(make_int2((int)I, (int)((uint32_t)flags << 27)));
      }
    }
    auto pad = [](std::vector<int2>& v, size_t mult) {
      while (v.size() % mult) v.push_back(make_int2(-1, 0));
    };
    pad(src, 256);
    pad(inert, 256);
    pt->n_colblk_src = (int)(src.size() / 256);
    pt->n_colblk_inert = (int)(inert.size() / 256);
    colItems = src;
    colItems.insert(colItems.end(), inert.begin(), inert.end());
  }
  std::vector<int4> rowItems;
  int sig_a = 0, sig_b = 0;   // gauge-invariant pair-double sign factors; 0 = not yet seen, 2 = mixed
  {
    // item = {row0 (local index on its owner), row1 (local index on its owner),
    //         flags | owner(row0) << 8 | owner(row1) << 16, column selector (0 all, 1 even / 2 odd column CTAs)}
    std::vector<int4> src, inert;
    const int me = sp->rank;
    auto pair_item = [&](int32_t S, int colsel) {   // S = global index of the src row of the pair
      const uint32_t c = codeA[S];
      const uint32_t ip = c >> 5;
      const int flags = bitof(c, 2) | (bitof(c, 3) << 1) | (bitof(codeA[ip], 3) << 2) | (bitof(c, 4) << 3);
      const int o0 = sq_row_owner(sp, S), o1 = sq_row_owner(sp, ip);
      src.push_back(make_int4((int)(S - sq_rank_start(sp, o0)), (int)((int64_t)ip - sq_rank_start(sp, o1)),
                              flags | (o0 << 8) | (o1 << 16), colsel));
      // row factor of sigma = dA * sSa * crossA(partner row)
      const int f = (bitof(c, 4) ^ bitof(c, 2) ^ bitof(codeA[ip], 3)) ? -1 : 1;
      sig_a = (sig_a == 0) ? f : (sig_a == f ? f : 2);
    };
    for (int32_t I : rows) {
      const uint32_t c = codeA[I], cls = c & 3u;
      if (cls == SQ_CLS_SRC)
        pair_item(I, 0);
      else
        inert.push_back(make_int4((int)(I - sp->row_begin), -1, (bitof(c, 3) << 1) | (me << 8) | (me << 16), 0));
    }
    // cross-device pairs: the src-row owner takes the even column CTAs, the tgt-row owner the odd ones
    for (int32_t S : crossSrc) pair_item(S, 1);
    for (int32_t S : crossTgt) pair_item(S, 2);
    const size_t TR = 8;
    auto pad = [&](std::vector<int4>& v) {
      while (v.size() % TR) v.push_back(make_int4(-1, -1, 0, 0));
    };
    pad(src);
    pad(inert);
    pt->n_rowchunk_src = (int)(src.size() / TR);
    pt->n_rowchunk_inert = (int)(inert.size() / TR);
    rowItems = src;
    rowItems.insert(rowItems.end(), inert.begin(), inert.end());
  }
  for (int64_t I = 0; I < sp->NB; ++I) {
    const uint32_t c = codeB[I];
    if ((c & 3u) != SQ_CLS_SRC) continue;
    // column factor of sigma = dB * crossB(this column) * sSb
    const int f = (bitof(c, 4) ^ bitof(c, 3) ^ bitof(c, 2)) ? -1 : 1;
    sig_b = (sig_b == 0) ? f : (sig_b == f ? f : 2);
  }
  pt->sigma = (sig_a == 2 || sig_b == 2) ? 0 : ((sig_a == 0 || sig_b == 0) ? 1 : sig_a * sig_b);
  pt->h_codeA = codeA;
  pt->h_codeB = codeB;
  if (sp->device < 0) return SQ_OK;   // host-only layout: plan / partition logic without device tables
  SQ_CUDA(cudaSetDevice(sp->device));
  SQ_CHECK(upload(&pt->d_codeA, codeA));
  SQ_CHECK(upload(&pt->d_codeB, codeB));
  SQ_CHECK(upload(&pt->d_rowsA, rows));
  SQ_CHECK(upload(&pt->d_colItems, colItems));
  SQ_CHECK(upload(&pt->d_rowItems, rowItems));
  return SQ_OK;
}

// normal-order a product  a+_{c0} a+_{c1} ... a_{d0} a_{d1} ...  of distinct operators: creators sorted
// descending, then annihilators sorted descending (fermionic_operator.py:27-104); returns the sign, or 0
// when an index repeats inside a block (the product vanishes).
static int normal_order_blocks(std::vector<int>& crea, std::vector<int>& anni) {
  int sign = 1;
  auto sort_desc = [&](std::vector<int>& v) -> bool {
    for (size_t x = 0; x < v.size(); ++x)
      for (size_t y = 0; y + 1 < v.size() - x; ++y) {
        if (v[y] == v[y + 1]) return false;
        if (v[y] < v[y + 1]) {
          std::swap(v[y], v[y + 1]);
          sign = -sign;
        }
      }
    for (size_t y = 0; y + 1 < v.size(); ++y)
      if (v[y] == v[y + 1]) return false;
    return true;
  };
  if (!sort_desc(crea)) return 0;
  if (!sort_desc(anni)) return 0;
  return sign;
}

// Tables for G = a+_{a} a+_{b} ... a_{j} a_{i}  (operators.py:145-359) given the reference's index tuple
// (i,j,..,a,b,..).
static int build_gen_tables(sq_space* sp, const std::vector<int>& idx, GenTables* gt, bool* null_op, bool* blocked) {
  const int rank = (int)idx.size() / 2;
  *null_op = false;
  *blocked = false;
  std::vector<int> crea(idx.begin() + rank, idx.end());            // a, b, c, ... in product order
  std::vector<int> anni(idx.begin(), idx.begin() + rank);          // i, j, k, ...
  std::reverse(anni.begin(), anni.end());                          // product order is ... a_k a_j a_i
  for (int so : idx)
    if (so < 0 || so >= 2 * sp->n_orb) {
      sq_set_error("spin-orbital index %d outside the active space", so);
      return SQ_ERR_INVALID;
    }
  const int sgnG = normal_order_blocks(crea, anni);
  if (sgnG == 0) {
    *null_op = true;
    return SQ_OK;
  }
  for (int c : crea)
    for (int d : anni)
      if (c == d) {
        sq_set_error("excitation generator with a shared creation/annihilation index %d is not a Givens rotation", c);
        return SQ_ERR_UNSUPPORTED;
      }
  std::vector<int32_t> label;
  for (int c : crea) label.push_back(2 * c + 1);
  for (int d : anni) label.push_back(2 * d);
  StringAction act;
  SQ_CHECK(sq_make_string_action(sp, label.data(), (int)label.size(), &act));
  gt->act = act;
  if (!act.conserving) {
    // spin-flipping generator leaves the (N_alpha, N_beta) sector: T|state> = 0 inside the space; the
    // reference would raise KeyError.  Mirror that.
    sq_set_error("excitation generator does not conserve N_alpha / N_beta");
    return SQ_ERR_OUTSIDE;
  }
  if (sp->alpha_cmask & act.flipA) {   // constrained alpha list: the targets are not in this space
    *blocked = true;
    return SQ_OK;
  }
  std::vector<int32_t> srcRows, tgtRows, colCode(sp->NB, -1);
  std::vector<int8_t> sgnRows;
  for (int64_t I = sp->row_begin; I < sp->row_end; ++I) {
    const uint32_t m = sp->strA[I];
    if ((m & act.occA) != act.occA || (m & act.empA) != 0u) continue;
    const int64_t t = sp->rankA[m ^ act.flipA];
    if (t < sp->row_begin || t >= sp->row_end) {
      sq_set_error("generator pairs alpha row %lld with row %lld on another device", (long long)I, (long long)t);
      return SQ_ERR_UNSUPPORTED;
    }
    srcRows.push_back((int32_t)I);
    tgtRows.push_back((int32_t)t);
    const int neg = ((sgnG * act.s0 < 0) ? 1 : 0) ^ neg_of(m, act.parA);
    sgnRows.push_back(neg ? -1 : 1);
  }
  int64_t nvalid = 0;
  for (int64_t I = 0; I < sp->NB; ++I) {
    const uint32_t m = sp->strB[I];
    if ((m & act.occB) != act.occB || (m & act.empB) != 0u) continue;
    const int32_t t = sp->rankB[m ^ act.flipB];
    colCode[I] = (t << 1) | neg_of(m, act.parB);
    ++nvalid;
  }
  gt->n_rows = (int64_t)srcRows.size();
  gt->n_cols_valid = nvalid;
  if (sp->device < 0) return SQ_OK;
  SQ_CUDA(cudaSetDevice(sp->device));
  SQ_CHECK(upload(&gt->d_srcRows, srcRows));
  SQ_CHECK(upload(&gt->d_tgtRows, tgtRows));
  SQ_CHECK(upload(&gt->d_sgnRows, sgnRows));
  SQ_CHECK(upload(&gt->d_colCode, colCode));
  return SQ_OK;
}

// ---------------------------------------------------------------------------------------------
// layout
// ---------------------------------------------------------------------------------------------
extern "C" int sq_layout_create(sq_space* sp, int n_ops, const int32_t* exc_type, const int32_t* idx_offsets,
                                const int32_t* idx_flat, sq_layout** out) {
  if (!sp || !out || n_ops < 0 || (n_ops > 0 && (!exc_type || !idx_offsets || !idx_flat))) return SQ_ERR_INVALID;
  *out = nullptr;
  for (int k = 0; k < n_ops; ++k)   // the offsets delimit the index tuples: non-negative, non-decreasing, at most 12 indices (sextuple)
    if (idx_offsets[k] < 0 || idx_offsets[k + 1] < idx_offsets[k] || idx_offsets[k + 1] - idx_offsets[k] > 12) {
      sq_set_error("sq_layout_create: bad index offsets for operator %d ([%d, %d))", k, idx_offsets[k], idx_offsets[k + 1]);
      return SQ_ERR_INVALID;
    }
  sq_layout* lay = new sq_layout();
  lay->sp = sp;
  lay->ops.resize(n_ops);
  int status = SQ_OK;
  for (int k = 0; k < n_ops && status == SQ_OK; ++k) {
    LayoutOp& op = lay->ops[k];
    op.type = exc_type[k];
    op.idx.assign(idx_flat + idx_offsets[k], idx_flat + idx_offsets[k + 1]);
    const int ni = (int)op.idx.size();
    auto get_pair = [&](int i, int a) -> int {
      auto key = std::make_pair(i, a);
      auto it = lay->pair_index.find(key);
      if (it != lay->pair_index.end()) return it->second;
      PairTables pt;
      status = build_pair_tables(sp, i, a, &pt);
      if (status != SQ_OK) return -1;
      lay->pairs.push_back(pt);
      lay->pair_index[key] = (int)lay->pairs.size() - 1;
      return (int)lay->pairs.size() - 1;
    };
    if (op.type == SQ_EXC_SA_SINGLE) {
      if (ni != 2) { sq_set_error("sa_single needs 2 indices, got %d", ni); status = SQ_ERR_INVALID; break; }
      op.pair = get_pair(op.idx[0], op.idx[1]);
    } else if (op.type >= SQ_EXC_SINGLE && op.type <= SQ_EXC_SEXTUPLE) {
      const int want = 2 * (op.type - SQ_EXC_SINGLE + 1);
      if (ni != want) {
        sq_set_error("excitation type %d needs %d indices, got %d", op.type, want, ni);
        status = SQ_ERR_INVALID;
        break;
      }
      if (op.type == SQ_EXC_DOUBLE && (op.idx[0] % 2 == 0) && op.idx[1] == op.idx[0] + 1 && (op.idx[2] % 2 == 0) &&
          op.idx[3] == op.idx[2] + 1 && op.idx[0] != op.idx[2]) {
        op.pair_double = true;
        op.pair = get_pair(op.idx[0] / 2, op.idx[2] / 2);
      } else {
        auto it = lay->gen_index.find(op.idx);
        if (it != lay->gen_index.end()) {
          op.gen = it->second;
          op.null_op = (op.gen < 0);
        } else {
          GenTables gt;
          bool null_op = false, blocked = false;
          status = build_gen_tables(sp, op.idx, &gt, &null_op, &blocked);
          if (status != SQ_OK) break;
          if (blocked) {
            op.blocked = true;   // not cached: the next operator with these indices is marked the same way
          } else if (null_op) {
            op.null_op = true;
            lay->gen_index[op.idx] = -1;
          } else {
            lay->gens.push_back(gt);
            op.gen = (int)lay->gens.size() - 1;
            lay->gen_index[op.idx] = op.gen;
          }
        }
      }
    } else if (op.type >= SQ_EXC_SA_DOUBLE_1 && op.type <= SQ_EXC_SA_DOUBLE_5) {
      if (ni != 4) { sq_set_error("sa_double needs 4 indices, got %d", ni); status = SQ_ERR_INVALID; break; }
      // generator strings are attached by the host (sq_layout_attach_generator)
    } else {
      sq_set_error("Got unknown excitation type code %d", op.type);
      status = SQ_ERR_INVALID;
    }
  }
  if (status != SQ_OK) {
    sq_layout_destroy(lay);
    return status;
  }
  *out = lay;
  return SQ_OK;
}

static int parse_strings(sq_space* sp, int n_strings, const int32_t* ops_flat, const int32_t* op_offsets,
                         const double* coeffs, int skip_outside, std::vector<StringAction>* acts,
                         std::vector<double>* cf) {
  acts->clear();
  cf->clear();
  if (n_strings < 0 || (n_strings > 0 && (!ops_flat || !op_offsets || !coeffs))) return SQ_ERR_INVALID;
  for (int s = 0; s < n_strings; ++s)
    if (op_offsets[s] < 0 || op_offsets[s + 1] < op_offsets[s]) {
      sq_set_error("operator string %d: bad offsets [%d, %d)", s, op_offsets[s], op_offsets[s + 1]);
      return SQ_ERR_INVALID;
    }
  for (int s = 0; s < n_strings; ++s) {
    StringAction a;
    SQ_CHECK(sq_make_string_action(sp, ops_flat + op_offsets[s], op_offsets[s + 1] - op_offsets[s], &a));
    if (!a.conserving) {
      if (skip_outside) continue;
      sq_set_error("operator string %d takes determinants outside the CI space (N_alpha/N_beta not conserved)", s);
      return SQ_ERR_OUTSIDE;
    }
    acts->push_back(a);
    cf->push_back(coeffs[s]);
  }
  return SQ_OK;
}

extern "C" int sq_layout_attach_generator(sq_layout* lay, int k, int n_strings, const int32_t* ops_flat,
                                          const int32_t* op_offsets, const double* coeffs) {
  if (!lay || k < 0 || k >= (int)lay->ops.size() || n_strings < 0) return SQ_ERR_INVALID;
  LayoutOp& op = lay->ops[k];
  if (op.type < SQ_EXC_SA_DOUBLE_1 || op.type > SQ_EXC_SA_DOUBLE_5) {
    sq_set_error("sq_layout_attach_generator: operator %d is not an sa_double", k);
    return SQ_ERR_INVALID;
  }
  GenOp* g = new GenOp();
  int st = parse_strings(lay->sp, n_strings, ops_flat, op_offsets, coeffs, 0, &g->strings, &g->coeffs);
  if (st != SQ_OK) {
    delete g;
    return st;
  }
  delete op.multi;
  op.multi = g;
  op.blocked = false;
  for (const StringAction& a : g->strings)
    if (lay->sp->alpha_cmask & a.flipA) op.blocked = true;
  return SQ_OK;
}

extern "C" int sq_layout_destroy(sq_layout* lay) {
  if (!lay) return SQ_OK;
  if (lay->sp->device >= 0) cudaSetDevice(lay->sp->device);
  for (auto& pt : lay->pairs) {
    cudaFree(pt.d_codeA);
    cudaFree(pt.d_codeB);
    cudaFree(pt.d_rowsA);
    cudaFree(pt.d_colItems);
    cudaFree(pt.d_rowItems);
  }
  for (auto& kv : lay->quads) sq_free_quad_tables(&kv.second);
  for (auto& kv : lay->wins) sq_free_win_tables(kv.second);
  for (auto& gt : lay->gens) {
    cudaFree(gt.d_srcRows);
    cudaFree(gt.d_tgtRows);
    cudaFree(gt.d_sgnRows);
    cudaFree(gt.d_colCode);
  }
  for (auto& op : lay->ops) delete op.multi;
  delete lay;
  return SQ_OK;
}

extern "C" int sq_layout_num_ops(const sq_layout* lay) { return lay ? (int)lay->ops.size() : -1; }

extern "C" int sq_layout_op_blocked(const sq_layout* lay, int k) {
  if (!lay || k < 0 || k >= (int)lay->ops.size()) return -1;
  const LayoutOp& op = lay->ops[k];
  return (op.blocked || (op.pair >= 0 && lay->pairs[op.pair].blocked)) ? 1 : 0;
}

// work-list statistics of operator k on this rank: out6 = {orbital-pair id or -1, local row pairs,
// local inert rows, cross-device row pairs this rank works on, cross_global, touched amplitudes}
extern "C" int sq_layout_op_stats(const sq_layout* lay, int k, int64_t* out6) {
  if (!lay || !out6 || k < 0 || k >= (int)lay->ops.size()) return SQ_ERR_INVALID;
  const LayoutOp& op = lay->ops[k];
  for (int i = 0; i < 6; ++i) out6[i] = 0;
  out6[0] = op.pair;
  if (op.pair >= 0) {
    const PairTables& pt = lay->pairs[op.pair];
    out6[1] = pt.n_src_rows;
    out6[2] = pt.n_rows - pt.n_src_rows;
    out6[3] = pt.n_cross_items;
    out6[4] = pt.cross_global ? 1 : 0;
    out6[5] = pt.touched;
  }
  return SQ_OK;
}

// ---------------------------------------------------------------------------------------------
// execution plan: consecutive operators on the same orbital pair fuse into one tile launch
// ---------------------------------------------------------------------------------------------
struct Run {
  int kind;          // 0 tile run, 1 generic single-string, 2 multi-string (sa_double), 3 skip
  int first, last;   // layout ops [first,last) in *execution* order (already reversed for dagger)
};

static inline bool is_tile_op(const LayoutOp& op) { return op.pair >= 0; }
static inline int tile_steps_of(const LayoutOp& op) { return op.pair_double ? 1 : 2; }

// order[] lists the operators in execution order; runs group them
static void plan_runs(const sq_layout* lay, const std::vector<int>& order, const double* thetas,
                      std::vector<std::vector<int>>* runs) {
  runs->clear();
  std::vector<int> cur;
  int cur_pair = -1, cur_steps = 0;
  auto flush = [&]() {
    if (!cur.empty()) runs->push_back(cur);
    cur.clear();
    cur_pair = -1;
    cur_steps = 0;
  };
  for (int k : order) {
    const LayoutOp& op = lay->ops[k];
    if (std::fabs(thetas[k]) < 1e-28) continue;   // operator_state_algebra.py:998
    if (is_tile_op(op)) {
      const int st = tile_steps_of(op);
      if (op.pair != cur_pair || cur_steps + st > SQ_MAX_PROGRAM) flush();
      cur.push_back(k);
      cur_pair = op.pair;
      cur_steps += st;
    } else {
      flush();
      runs->push_back(std::vector<int>{k});
    }
  }
  flush();
}

static void exec_order(int first, int last, int dagger, std::vector<int>* order) {
  order->clear();
  if (!dagger)
    for (int k = first; k < last; ++k) order->push_back(k);
  else
    for (int k = last - 1; k >= first; --k) order->push_back(k);
}

// ---- quad fusion: two consecutive bricks on disjoint orbital pairs in one sweep (sqsv_quad.cu) ----
int sq_build_quad_tables(sq_space* sp, const PairTables& p1, const PairTables& p2, int pair1, int pair2, QuadTables* qt);
void sq_free_quad_tables(QuadTables* qt);

static bool quad_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SQ_QUAD");   // SQ_QUAD=0 disables the fusion (A/B comparison, debugging)
    const char* t = getenv("SQ_TILE_KERNEL");
    v = ((e && e[0] == '0') || (t && t[0] == '1')) ? 0 : 1;
  }
  return v == 1;
}

static int get_quad(sq_layout* lay, int pA, int pB, const QuadTables** out) {
  *out = nullptr;
  if (pA < 0 || pB < 0 || pA == pB) return SQ_OK;
  if (lay->pairs[pA].blocked || lay->pairs[pB].blocked) return SQ_OK;
  auto key = std::make_pair(pA, pB);
  auto it = lay->quads.find(key);
  if (it == lay->quads.end()) {
    QuadTables qt;
    SQ_CHECK(sq_build_quad_tables(lay->sp, lay->pairs[pA], lay->pairs[pB], pA, pB, &qt));
    it = lay->quads.emplace(key, qt).first;
  }
  *out = &it->second;
  return SQ_OK;
}

// ---- launch plan: window sweeps (sqsv_win.cu), quad fusion, single bricks, generic operators ----
// A run is one brick (consecutive operators on one orbital pair) or one non-tile operator.  Bricks on disjoint
// orbital pairs commute (even products of ladder operators on disjoint spin orbitals), so a brick may be
// executed early if it commutes with every not-yet-executed run in front of it.  The planner repeatedly takes
// the window that can absorb the most such bricks; non-tile operators are barriers.

static int g_plan_version = 0;   // bumped by sq_set_option: cached plans of older versions are dropped
static int g_win_grad = 0;       // sq_set_option("wingrad", "1"): gradient sweep through the window kernel
static int g_quad_grad = 1;      // sq_set_option("quadgrad", "0"): one brick per launch in the gradient sweep (no quad_grad_kernel)

struct WinConfig {
  bool enabled = true;
  int widths[3] = {6, 5, 4};     // window widths tried (orbitals)
  int smem_kb = 72;              // largest CTA footprint admitted (3 CTAs per SM)
  int min_suffix = 3;            // windows below the top one need at least this many orbitals above them
  int max_bricks = SQ_WIN_MAX_BRICKS;
  int min_bricks = 3;            // smaller isolated groups go to the tile / quad kernels
};

static void parse_win_config(const char* e, WinConfig* cfg) {
  // "0" disables, "1" / "" restores the defaults, otherwise
  // "w1:w2:w3,smem_kb,min_suffix,max_bricks,min_bricks" (trailing fields optional)
  *cfg = WinConfig();
  if (!e || !e[0] || (e[0] == '1' && e[1] == 0)) return;
  if (e[0] == '0' && e[1] == 0) {
    cfg->enabled = false;
    return;
  }
  int w[3] = {0, 0, 0}, v[4] = {cfg->smem_kb, cfg->min_suffix, cfg->max_bricks, cfg->min_bricks};
  const int got = sscanf(e, "%d:%d:%d,%d,%d,%d,%d", &w[0], &w[1], &w[2], &v[0], &v[1], &v[2], &v[3]);
  if (got >= 3) {
    for (int i = 0; i < 3; ++i) cfg->widths[i] = w[i];
    cfg->smem_kb = v[0];
    cfg->min_suffix = v[1];
    cfg->max_bricks = std::max(1, std::min(v[2], SQ_WIN_MAX_BRICKS));
    cfg->min_bricks = std::max(1, v[3]);
  }
}

static WinConfig& win_config() {
  static WinConfig cfg;
  static bool init = false;
  if (!init) {
    init = true;
    parse_win_config(getenv("SQ_WIN"), &cfg);
    const char* t = getenv("SQ_TILE_KERNEL");
    if (t && t[0] == '1') cfg.enabled = false;
  }
  return cfg;
}

static WinConfig& grad_win_config();
// run-time switches for A/B comparisons and tests: name "win" takes the SQ_WIN syntax
extern "C" int sq_set_option(const char* name, const char* value) {
  if (!name) return SQ_ERR_INVALID;
  if (strcmp(name, "win") == 0) {
    parse_win_config(value, &win_config());
    ++g_plan_version;
    return SQ_OK;
  }
  if (strcmp(name, "win3") == 0) {   // window sweeps: "0" (default) win_kernel, "1" win3_kernel (orbital-triple register blocks, merged tiles; measured slower)
    sq_win3_set_enabled(value && value[0] == '1');
    return SQ_OK;
  }
  if (strcmp(name, "wingrad") == 0) {
    g_win_grad = (value && value[0] == '0') ? 0 : 1;
    return SQ_OK;
  }
  if (strcmp(name, "quadgrad") == 0) {   // gradient sweep: "1" (default) two commuting bricks per launch, "0" one brick per launch
    g_quad_grad = (value && value[0] == '0') ? 0 : 1;
    return SQ_OK;
  }
  if (strcmp(name, "wingrad_win") == 0) {   // window configuration of the gradient sweep (SQ_WIN syntax)
    parse_win_config(value, &grad_win_config());
    ++g_plan_version;
    return SQ_OK;
  }
  if (strcmp(name, "panel") == 0) {   // determinants per sigma / RDM panel for spaces created afterwards ("0": default)
    sq_hamiltonian_set_panel_width(value ? atoll(value) : 0);
    return SQ_OK;
  }
  if (strcmp(name, "rows_cfg") == 0) {   // "threads,chunks" of the row kernels (chunks 0: automatic)
    int t = 1024, c = 0;
    if (value) sscanf(value, "%d,%d", &t, &c);
    sq_hamiltonian_set_rows_cfg(t, c);
    return SQ_OK;
  }
  if (strcmp(name, "rows") == 0) {   // sigma / RDM panel kernels: "0" (default) determinant-per-thread, "1" row-per-CTA (slower, kept as evidence)
    sq_hamiltonian_set_rows_mode(value && value[0] == '1');
    return SQ_OK;
  }
  if (strcmp(name, "pipeline") == 0) {   // sigma / RDM panels: "1" (default) overlaps gather, DGEMM and scatter of neighbouring panels
    sq_hamiltonian_set_pipeline(!(value && value[0] == '0'));
    return SQ_OK;
  }
  if (strcmp(name, "rdm_tri") == 0) {   // RDMs with bra == ket: "1" three half-size DGEMMs (3/4 of the flops), "0" (default) one DGEMM
    sq_hamiltonian_set_rdm_tri(value && value[0] == '1');
    return SQ_OK;
  }
  if (strcmp(name, "etab") == 0) {   // E_pq table of the sigma / RDM panel kernels: "smem" (default) or "const"
    sq_hamiltonian_set_etab_mode(value && strcmp(value, "const") == 0);
    sq_hamiltonian_set_etab_alu(value && strcmp(value, "alu") == 0);
    sq_hamiltonian_set_etab_tab(value && strcmp(value, "tab") == 0);
    return SQ_OK;
  }
  if (strcmp(name, "rdm_sym") == 0) {   // sq_rdm12 with bra == ket: "1" (default) two symmetric S / A Gram matrices, "0" the plain n^2 x n^2 one
    sq_hamiltonian_set_rdm_sym(!(value && value[0] == '0'));
    return SQ_OK;
  }
  if (strcmp(name, "sigma_spinsym") == 0) {   // sigma of a spin-flip symmetric vector from the determinants above the diagonal: "1" (default) / "0"
    sq_hamiltonian_set_sigma_spinsym(!(value && value[0] == '0'));
    sq_hamiltonian_set_spinsym_blk(!(value && strcmp(value, "tri") == 0));   // "tri": determinant-per-thread kernels instead of the 32 x 32 blocks
    return SQ_OK;
  }
  if (strcmp(name, "sigma_fused") == 0) {   // sigma: "0" (default) three-kernel panel pipeline, "1" fused gather -> DMMA -> scatter kernel (slower)
    sq_hamiltonian_set_sigma_fused(value && value[0] == '1');
    return SQ_OK;
  }
  if (strcmp(name, "sgemm_wm") == 0) {   // sigma DMMA kernel: "2" (default) four warps per CTA, "3" six warps (slower)
    sq_sigma_gemm_set_row_parts(value ? atoi(value) : 2);
    return SQ_OK;
  }
  if (strcmp(name, "sgemm_cta") == 0) {   // sigma DMMA kernel: "2" (default) or "1" CTAs of 4 warps per SM
    sq_sigma_gemm_set_residency(value ? atoi(value) : 2);
    return SQ_OK;
  }
  if (strcmp(name, "reshard") == 0) {   // re-shard kernel of sharded vectors: "tma" (default, bulk-copy engine) or "lsu"
    sq_reshard_set_mode(value && strcmp(value, "lsu") == 0);
    return SQ_OK;
  }
  sq_set_error("sq_set_option: unknown option '%s'", name);
  return SQ_ERR_INVALID;
}

struct WinCand {
  int w0, H;
  size_t smem;
  bool dead = false;
};

// The gradient sweep stages TWO vectors per batch, so it plans with its own window configuration (narrower windows: the tiles
// of bra and ket together must leave room for two CTAs per SM) and keeps its own plan cache (sq_layout::grad_plans).
static WinConfig& grad_win_config() {
  static WinConfig cfg;
  static bool init = false;
  if (!init) {
    init = true;
    const char* e = getenv("SQ_WINGRAD_WIN");
    parse_win_config(e ? e : "5:4:0,40,3,16,2", &cfg);
  }
  return cfg;
}
static const WinConfig* g_active_cfg = nullptr;   // configuration the planner runs with (nullptr: win_config())
static inline const WinConfig& active_config() { return g_active_cfg ? *g_active_cfg : win_config(); }

static void win_candidates(const sq_space* sp, std::vector<WinCand>* out) {
  out->clear();
  const WinConfig& cfg = active_config();
  if (!cfg.enabled) return;
  const int n = sp->n_orb;
  for (int wi = 0; wi < 3; ++wi) {
    int H = cfg.widths[wi];
    if (H <= 1) continue;
    if (H > n) H = n;
    bool dup = false;
    for (int wj = 0; wj < wi; ++wj) dup |= (std::min(cfg.widths[wj], n) == H);
    if (dup) continue;
    for (int w0 = 0; w0 + H <= n; ++w0) {
      const int suffix = n - w0 - H;
      if (suffix != 0 && suffix < cfg.min_suffix) continue;   // short suffix runs would not coalesce
      WinCand c{w0, H, 0};
      // work lists are bounded by (strings per side)^2; the exact size comes with the tables
      const int ma = sq_win_max_class(n, sp->n_alpha, w0, H), mb = sq_win_max_class(n, sp->n_beta, w0, H);
      c.smem = sq_win_smem_bytes(ma, mb, suffix ? 16 : 18, 1, ma, mb, (ma * mb) / 8 + 1, (ma * mb) / 3 + 1, 8);
      if (c.smem <= (size_t)cfg.smem_kb * 1024) out->push_back(c);
    }
  }
}

static uint32_t run_orbitals(const sq_layout* lay, const std::vector<int>& run) {
  const LayoutOp& op = lay->ops[run[0]];
  if (!is_tile_op(op)) return 0xffffffffu;
  const PairTables& pt = lay->pairs[op.pair];
  return (1u << pt.i) | (1u << pt.a);
}


static int plan_launches_uncached(sq_layout* lay, const std::vector<std::vector<int>>& runs, std::vector<Launch>* out);

// plans are cached on the layout (the beam search costs tens of milliseconds for a 720-operator circuit)
static int plan_launches(sq_layout* lay, const std::vector<std::vector<int>>& runs, std::vector<Launch>* out, bool grad = false) {
  if (lay->plan_version != g_plan_version) {
    lay->plans.clear();
    lay->grad_plans.clear();
    lay->plan_version = g_plan_version;
  }
  std::vector<PlanCache>& cache = grad ? lay->grad_plans : lay->plans;
  for (auto& pc : cache)
    if (pc.runs == runs) {
      *out = pc.launches;
      return SQ_OK;
    }
  g_active_cfg = grad ? &grad_win_config() : nullptr;
  const int rc = plan_launches_uncached(lay, runs, out);
  g_active_cfg = nullptr;
  SQ_CHECK(rc);
  if (cache.size() >= 64) cache.erase(cache.begin());   // the re-sharding driver plans one run list per phase
  cache.push_back({runs, *out});
  return SQ_OK;
}

static int plan_launches_uncached(sq_layout* lay, const std::vector<std::vector<int>>& runs, std::vector<Launch>* out) {
  out->clear();
  sq_space* sp = lay->sp;
  const WinConfig& cfg = active_config();
  const int m = (int)runs.size();
  std::vector<WinCand> cands;
  win_candidates(sp, &cands);
  std::vector<char> done(m, 0);
  std::vector<uint32_t> orb(m);
  std::vector<int> pair(m);
  for (int t = 0; t < m; ++t) {
    orb[t] = run_orbitals(lay, runs[t]);
    pair[t] = lay->ops[runs[t][0]].pair;
  }
  auto simulate = [&](const WinCand& c, int head, std::vector<int>* sel) {
    sel->clear();
    uint32_t blocked = 0;
    const int wlo = c.w0, whi = c.w0 + c.H;
    const uint32_t full = ((whi >= 32) ? 0xffffffffu : ((1u << whi) - 1u)) & ~((1u << wlo) - 1u);
    for (int t = head; t < m && (int)sel->size() < cfg.max_bricks; ++t) {
      if (done[t]) continue;
      if (pair[t] < 0) break;   // barrier
      if (!(orb[t] & blocked) && sq_win_pair_ok(lay, pair[t], c.w0, c.H)) sel->push_back(t);
      else blocked |= orb[t];
      if ((blocked & full) == full) break;
    }
  };
  // usable candidates (tables are built once per layout and window)
  for (WinCand& c : cands) {
    bool any = false;
    for (int t = 0; t < m && !any; ++t) any = pair[t] >= 0 && sq_win_pair_ok(lay, pair[t], c.w0, c.H);
    if (!any) { c.dead = true; continue; }
    const WinTables* wt = nullptr;
    SQ_CHECK(sq_get_win(sp, lay, c.w0, c.H, &wt));
    if (!wt || !wt->ok) c.dead = true;
  }
  auto emit_single = [&](int head, bool front) -> int {
    // one brick alone; the front brick (everything before it executed) may be fused with the next pending brick
    // when they commute (quad kernel)
    int nxt = head + 1;
    while (nxt < m && done[nxt]) ++nxt;
    if (front && pair[head] >= 0 && nxt < m && pair[nxt] >= 0 && quad_enabled()) {
      const QuadTables* qt = nullptr;
      SQ_CHECK(get_quad(lay, pair[head], pair[nxt], &qt));
      if (qt && qt->ok) {
        Launch l;
        l.kind = 1;
        l.runs = {head, nxt};
        l.qt = qt;
        out->push_back(l);
        done[head] = done[nxt] = 1;
        return SQ_OK;
      }
    }
    Launch l;
    l.kind = 0;
    l.runs = {head};
    out->push_back(l);
    done[head] = 1;
    return SQ_OK;
  };
  // Beam search over window sequences for every stretch of bricks between two barriers: a state is the set of
  // executed runs; one step = one sweep (the window with its greedy closure of executable bricks).  States of
  // equal depth are ranked by the number of executed bricks.
  struct Node {
    std::vector<char> done;
    int head, parent, cand, n_done;
    std::vector<int> sel;
  };
  const int BEAM = 24;
  int head = 0;
  std::vector<int> sel;
  while (head < m) {
    if (done[head]) { ++head; continue; }
    bool coverable = false;
    if (pair[head] >= 0)
      for (const WinCand& c : cands) coverable |= !c.dead && sq_win_pair_ok(lay, pair[head], c.w0, c.H);
    if (!coverable) { SQ_CHECK(emit_single(head, true)); continue; }
    int seg_end = head;   // bricks up to the next barrier
    while (seg_end < m && pair[seg_end] >= 0) ++seg_end;
    std::vector<Node> nodes;
    nodes.push_back({done, head, -1, -1, 0, {}});
    std::vector<int> frontier = {0};
    int goal = -1;
    while (goal < 0) {
      std::vector<int> next;
      std::map<std::vector<char>, int> seen;
      for (int ni : frontier) {
        std::vector<char> cur = nodes[ni].done;   // copy: nodes may reallocate
        int h = nodes[ni].head;
        while (h < seg_end && cur[h]) ++h;
        if (h >= seg_end) { goal = ni; break; }
        done.swap(cur);   // simulate() reads `done`
        bool head_covered = false;
        std::vector<std::pair<int, std::vector<int>>> moves;
        for (size_t ci = 0; ci < cands.size(); ++ci) {
          if (cands[ci].dead) continue;
          simulate(cands[ci], h, &sel);
          while (!sel.empty() && sel.back() >= seg_end) sel.pop_back();
          if (sel.empty()) continue;
          head_covered |= (sel[0] == h);
          moves.push_back({(int)ci, sel});
        }
        done.swap(cur);
        if (!head_covered) moves.push_back({-1, std::vector<int>{h}});   // a brick no window takes: single launch
        for (auto& mv : moves) {
          std::vector<char> d2 = cur;
          for (int t : mv.second) d2[t] = 1;
          if (seen.count(d2)) continue;
          seen[d2] = 1;
          nodes.push_back({d2, h, ni, mv.first, nodes[ni].n_done + (int)mv.second.size(), mv.second});
          next.push_back((int)nodes.size() - 1);
        }
      }
      if (goal >= 0) break;
      std::sort(next.begin(), next.end(), [&](int a, int b) {
        return nodes[a].n_done != nodes[b].n_done ? nodes[a].n_done > nodes[b].n_done : nodes[a].head > nodes[b].head;
      });
      if ((int)next.size() > BEAM) next.resize(BEAM);
      frontier.swap(next);
      if (frontier.empty()) {
        sq_set_error("launch planner: no progress");
        return SQ_ERR_INVALID;
      }
    }
    std::vector<int> path;
    for (int ni = goal; nodes[ni].parent >= 0; ni = nodes[ni].parent) path.push_back(ni);
    std::reverse(path.begin(), path.end());
    // a short window sweep pays two extra gauge sweeps unless a neighbouring launch is a window sweep as well
    std::vector<char> as_win(path.size(), 0);
    for (size_t k = 0; k < path.size(); ++k) as_win[k] = nodes[path[k]].cand >= 0 && (int)nodes[path[k]].sel.size() >= cfg.min_bricks;
    for (size_t k = 0; k < path.size(); ++k)
      if (!as_win[k] && nodes[path[k]].cand >= 0 && ((k > 0 && as_win[k - 1]) || (k + 1 < path.size() && as_win[k + 1]))) as_win[k] = 2;
    for (size_t k = 0; k < path.size(); ++k) {
      const Node& nd = nodes[path[k]];
      if (as_win[k]) {
        Launch l;
        l.kind = 2;
        l.runs = nd.sel;
        SQ_CHECK(sq_get_win(sp, lay, cands[nd.cand].w0, cands[nd.cand].H, &l.wt));
        out->push_back(l);
        for (int t : nd.sel) done[t] = 1;
      } else {
        for (int t : nd.sel)
          if (!done[t]) SQ_CHECK(emit_single(t, false));
      }
    }
    head = seg_end;
  }
  return SQ_OK;
}

static int launch_cost(const sq_layout* lay, const std::vector<std::vector<int>>& runs, const Launch& l, int* n_kernels,
                       int64_t* touched) {
  const sq_space* sp = lay->sp;
  const std::vector<int>& r = runs[l.runs[0]];
  const LayoutOp& op = lay->ops[r[0]];
  *n_kernels = 0;
  *touched = 0;
  if (l.kind == 2) {
    *n_kernels = 1;
    *touched = l.wt->touched;
  } else if (l.kind == 1) {
    *n_kernels = 1;
    *touched = l.qt->touched;
  } else if (is_tile_op(op)) {
    const PairTables& pt = lay->pairs[op.pair];
    bool has_single = false;
    for (int k : r) has_single |= !lay->ops[k].pair_double;
    *n_kernels = 1;
    *touched = has_single ? pt.touched : 2 * pt.n_src_rows * pt.n_src_cols;
  } else if (op.null_op) {
  } else if (op.gen >= 0) {
    *n_kernels = 1;
    *touched = 2 * lay->gens[op.gen].n_rows * lay->gens[op.gen].n_cols_valid;
  } else if (op.multi) {
    static const int napp[5] = {2, 4, 4, 8, 10};
    // each power: gather (read + write) and axpy (2 reads + write) over the whole vector
    *n_kernels = 2 * napp[op.type - SQ_EXC_SA_DOUBLE_1];
    *touched = (int64_t)napp[op.type - SQ_EXC_SA_DOUBLE_1] * 3 * sp->local_len();
  }
  return SQ_OK;
}

static int plan_range(sq_layout* lay, int first, int last, std::vector<std::vector<int>>* runs, std::vector<Launch>* launches) {
  std::vector<double> th(lay->ops.size(), 1.0);
  std::vector<int> order;
  exec_order(first, last, 0, &order);
  plan_runs(lay, order, th.data(), runs);
  return plan_launches(lay, *runs, launches);
}

extern "C" int sq_layout_num_launches(const sq_layout* lay_c, int first, int last) {
  sq_layout* lay = const_cast<sq_layout*>(lay_c);
  if (!lay || first < 0 || last > (int)lay->ops.size() || first > last) return -1;
  std::vector<std::vector<int>> runs;
  std::vector<Launch> launches;
  if (plan_range(lay, first, last, &runs, &launches) != SQ_OK) return -1;
  int n = 0;
  bool in_gauge = false;
  for (auto& l : launches) {
    int nk;
    int64_t t;
    launch_cost(lay, runs, l, &nk, &t);
    n += nk;
    if ((l.kind == 2) != in_gauge) { ++n; in_gauge = !in_gauge; }
  }
  return n + (in_gauge ? 1 : 0);
}

// amplitudes read+written by the launches of ops [first,last): the algorithmic traffic of sq_ups_apply is
// 16 bytes (one fp64 read + one fp64 write) per touched amplitude per launch.
extern "C" int64_t sq_layout_touched_amplitudes(const sq_layout* lay_c, int first, int last) {
  sq_layout* lay = const_cast<sq_layout*>(lay_c);
  if (!lay || first < 0 || last > (int)lay->ops.size() || first > last) return -1;
  std::vector<std::vector<int>> runs;
  std::vector<Launch> launches;
  if (plan_range(lay, first, last, &runs, &launches) != SQ_OK) return -1;
  int64_t total = 0;
  bool in_gauge = false;
  for (auto& l : launches) {
    int nk;
    int64_t t;
    launch_cost(lay, runs, l, &nk, &t);
    total += t;
    if ((l.kind == 2) != in_gauge) { total += lay->sp->local_len(); in_gauge = !in_gauge; }   // gauge sweep
  }
  return total + (in_gauge ? lay->sp->local_len() : 0);
}

// plan summary for tools / tests: out[0] = launches, out[1] = window sweeps, out[2] = bricks inside window sweeps,
// out[3] = quad launches, out[4] = single-brick launches, out[5] = other launches
extern "C" int sq_layout_plan_stats(const sq_layout* lay_c, int first, int last, int64_t* out6) {
  sq_layout* lay = const_cast<sq_layout*>(lay_c);
  if (!lay || !out6 || first < 0 || last > (int)lay->ops.size() || first > last) return SQ_ERR_INVALID;
  std::vector<std::vector<int>> runs;
  std::vector<Launch> launches;
  SQ_CHECK(plan_range(lay, first, last, &runs, &launches));
  for (int i = 0; i < 6; ++i) out6[i] = 0;
  for (auto& l : launches) {
    ++out6[0];
    if (l.kind == 2) { ++out6[1]; out6[2] += (int64_t)l.runs.size(); }
    else if (l.kind == 1) ++out6[3];
    else if (is_tile_op(lay->ops[runs[l.runs[0]][0]])) ++out6[4];
    else ++out6[5];
  }
  if (getenv("SQ_PLAN_DEBUG")) {
    for (auto& l : launches) {
      if (l.kind == 2) {
        fprintf(stderr, "win [%d,%d) smem=%zu :", l.wt->w0, l.wt->w0 + l.wt->H, l.wt->smem);
        for (int t : l.runs) fprintf(stderr, " (%d,%d)", lay->pairs[lay->ops[runs[t][0]].pair].i, lay->pairs[lay->ops[runs[t][0]].pair].a);
        fprintf(stderr, "\n");
        int pidx[SQ_WIN_MAX_BRICKS], nb = 0;
        for (int t : l.runs) pidx[nb++] = lay->ops[runs[t][0]].pair;
        sq_win3_print_stats(*l.wt, pidx, nb);
      } else if (l.kind == 1) {
        fprintf(stderr, "quad\n");
      } else {
        fprintf(stderr, "single\n");
      }
    }
  }
  return SQ_OK;
}

// the same summary for an explicit operator list in execution order (one phase of the re-sharding driver):
// out8 = {launches, window sweeps, bricks inside window sweeps, quad launches, single-brick launches, other launches,
//         kernels (without gauge sweeps), amplitudes those kernels read and write}
extern "C" int sq_layout_plan_stats_list(const sq_layout* lay_c, int n_list, const int32_t* op_list, int64_t* out8) {
  sq_layout* lay = const_cast<sq_layout*>(lay_c);
  if (!lay || !out8 || n_list < 0 || (n_list > 0 && !op_list)) return SQ_ERR_INVALID;
  for (int i = 0; i < 8; ++i) out8[i] = 0;
  std::vector<int> order;
  for (int i = 0; i < n_list; ++i) {
    if (op_list[i] < 0 || op_list[i] >= (int)lay->ops.size()) return SQ_ERR_INVALID;
    const LayoutOp& op = lay->ops[op_list[i]];
    if (op.blocked || (op.pair >= 0 && lay->pairs[op.pair].blocked)) {
      sq_set_error("sq_layout_plan_stats_list: operator %d cannot run in this space", op_list[i]);
      return SQ_ERR_UNSUPPORTED;
    }
    order.push_back(op_list[i]);
  }
  std::vector<double> th(lay->ops.size(), 1.0);
  std::vector<std::vector<int>> runs;
  std::vector<Launch> launches;
  plan_runs(lay, order, th.data(), &runs);
  SQ_CHECK(plan_launches(lay, runs, &launches));
  for (auto& l : launches) {
    ++out8[0];
    if (l.kind == 2) { ++out8[1]; out8[2] += (int64_t)l.runs.size(); }
    else if (l.kind == 1) ++out8[3];
    else if (is_tile_op(lay->ops[runs[l.runs[0]][0]])) ++out8[4];
    else ++out8[5];
    int nk;
    int64_t t;
    launch_cost(lay, runs, l, &nk, &t);
    out8[6] += nk;
    out8[7] += t;
  }
  return SQ_OK;
}

// the plan itself, for tests of the planner: operators of [first,last) in EXECUTION order (dagger: reversed circuit) with the
// index of the launch each one rides in.  thetas_host may be NULL (all operators active) -- zero angles are skipped as in
// sq_ups_apply.  Returns SQ_ERR_INVALID if cap is too small; *n_out = number of entries.
extern "C" int sq_layout_plan_export(const sq_layout* lay_c, const double* thetas_host, int first, int last, int dagger,
                                     int32_t* ops_out, int32_t* launch_out, int cap, int* n_out) {
  sq_layout* lay = const_cast<sq_layout*>(lay_c);
  if (!lay || !ops_out || !launch_out || !n_out || first < 0 || last > (int)lay->ops.size() || first > last) return SQ_ERR_INVALID;
  std::vector<double> th(lay->ops.size(), 1.0);
  if (thetas_host)
    for (size_t k = 0; k < th.size(); ++k) th[k] = thetas_host[k];
  std::vector<int> order;
  exec_order(first, last, dagger ? 1 : 0, &order);
  std::vector<std::vector<int>> runs;
  std::vector<Launch> launches;
  plan_runs(lay, order, th.data(), &runs);
  SQ_CHECK(plan_launches(lay, runs, &launches));
  int n = 0, li = 0;
  for (auto& l : launches) {
    for (int t : l.runs)
      for (int k : runs[t]) {
        if (n >= cap) return SQ_ERR_INVALID;
        ops_out[n] = k;
        launch_out[n] = li;
        ++n;
      }
    ++li;
  }
  *n_out = n;
  return SQ_OK;
}

// closed forms of exp(theta T) for the spin-adapted doubles (operator_state_algebra.py:1086-1409):
// out += sum_m w_m(theta) T^m out with the reference's coefficient tables.
static int sa_double_poly(sq_space* sp, const GenOp& g, int type, double theta, double* state, cudaStream_t st) {
  SQ_CHECK(sq_ensure_work(sp, 0));
  SQ_CHECK(sq_ensure_work(sp, 1));
  double* tmp[2] = {sp->d_work[0], sp->d_work[1]};
  std::vector<double> w;   // weight of T^m out, m = 1..
  const double r2 = std::sqrt(2.0), r3 = std::sqrt(3.0);
  if (type == SQ_EXC_SA_DOUBLE_1) {
    w = {std::sin(theta), 1.0 - std::cos(theta)};   // osa.py:1069-1085
  } else if (type == SQ_EXC_SA_DOUBLE_2 || type == SQ_EXC_SA_DOUBLE_3) {
    const double S[2] = {1.0, r2 / 2};
    const double k1[2] = {-1, 2 * r2}, k3[2] = {-2, 2 * r2}, k2[2] = {1, -4}, k4[2] = {2, -4};
    const double* ks[4] = {k1, k2, k3, k4};
    for (int m = 0; m < 4; ++m) {
      double v = 0;
      for (int f = 0; f < 2; ++f) v += ks[m][f] * ((m % 2 == 0) ? std::sin(S[f] * theta) : (std::cos(S[f] * theta) - 1));
      w.push_back(v);
    }
  } else if (type == SQ_EXC_SA_DOUBLE_4) {
    const double S[4] = {1.0, r2, r2 / 2, 0.5};
    const double k1[4] = {2.0 / 3, -r2 / 42, -8 * r2 / 3, 128.0 / 21};
    const double k3[4] = {13.0 / 3, -r2 / 6, -44 * r2 / 3, 64.0 / 3};
    const double k5[4] = {22.0 / 3, -r2 / 3, -52 * r2 / 3, 64.0 / 3};
    const double k7[4] = {8.0 / 3, -4 * r2 / 21, -16 * r2 / 3, 128.0 / 21};
    const double k2[4] = {-2.0 / 3, 1.0 / 42, 16.0 / 3, -256.0 / 21};
    const double k4[4] = {-13.0 / 3, 1.0 / 6, 88.0 / 3, -128.0 / 3};
    const double k6[4] = {-22.0 / 3, 1.0 / 3, 104.0 / 3, -128.0 / 3};
    const double k8[4] = {-8.0 / 3, 4.0 / 21, 32.0 / 3, -256.0 / 21};
    const double* ks[8] = {k1, k2, k3, k4, k5, k6, k7, k8};
    for (int m = 0; m < 8; ++m) {
      double v = 0;
      for (int f = 0; f < 4; ++f) v += ks[m][f] * ((m % 2 == 0) ? std::sin(S[f] * theta) : (std::cos(S[f] * theta) - 1));
      w.push_back(v);
    }
  } else if (type == SQ_EXC_SA_DOUBLE_5) {
    const double S[5] = {r2, r2 / 2, r3 / 3, r3 / 2, r3 / 6};
    const double k1[5] = {r2 / 1150, 8 * r2 / 5, -54 * r3 / 25, -16 * r3 / 75, 432 * r3 / 115};
    const double k3[5] = {11 * r2 / 690, 404 * r2 / 15, -171 * r3 / 5, -56 * r3 / 15, 2952 * r3 / 115};
    const double k5[5] = {133 * r2 / 1725, 308 * r2 / 3, -2718 * r3 / 25, -1192 * r3 / 75, 1368 * r3 / 23};
    const double k7[5] = {16 * r2 / 115, 608 * r2 / 5, -576 * r3 / 5, -112 * r3 / 5, 6192 * r3 / 115};
    const double k9[5] = {48 * r2 / 575, 192 * r2 / 5, -864 * r3 / 25, -192 * r3 / 25, 1728 * r3 / 115};
    const double k2[5] = {-1.0 / 1150, -16.0 / 5, 162.0 / 25, 32.0 / 75, -2592.0 / 115};
    const double k4[5] = {-11.0 / 690, -808.0 / 15, 513.0 / 5, 112.0 / 15, -17712.0 / 115};
    const double k6[5] = {-133.0 / 1725, -616.0 / 3, 8154.0 / 25, 2384.0 / 75, -8208.0 / 23};
    const double k8[5] = {-16.0 / 115, -1216.0 / 5, 1728.0 / 5, 224.0 / 5, -37152.0 / 115};
    const double k10[5] = {-48.0 / 575, -384.0 / 5, 2592.0 / 25, 384.0 / 25, -10368.0 / 115};
    const double* ks[10] = {k1, k2, k3, k4, k5, k6, k7, k8, k9, k10};
    for (int m = 0; m < 10; ++m) {
      double v = 0;
      for (int f = 0; f < 5; ++f) v += ks[m][f] * ((m % 2 == 0) ? std::sin(S[f] * theta) : (std::cos(S[f] * theta) - 1));
      w.push_back(v);
    }
  } else {
    return SQ_ERR_INVALID;
  }
  // sa_double_1 in the reference evaluates T and T^2 on the *old* out before updating; the higher
  // cases update out progressively -- both are the same polynomial in T applied to the old vector,
  // because every power is generated from the previous power (tmp), never from the updated out.
  const double* src = state;
  for (size_t m = 0; m < w.size(); ++m) {
    double* dst = tmp[m & 1];
    if (m == 0) {
      // T^1: source is the state itself; it must not be modified until tmp holds T state
      SQ_CHECK(sq_launch_gather(sp, g.strings, g.coeffs, src, dst, 0, st));
    } else {
      SQ_CHECK(sq_launch_gather(sp, g.strings, g.coeffs, tmp[(m - 1) & 1], dst, 0, st));
    }
    SQ_CHECK(sq_launch_axpy(sp, w[m], dst, state, st));
  }
  return SQ_OK;
}

static int run_tile(sq_space* sp, sq_layout* lay, const std::vector<int>& run, const double* thetas, int dagger,
                    TileStep* steps, int* n_steps, int* step_op) {
  int n = 0;
  for (int k : run) {
    const LayoutOp& op = lay->ops[k];
    const double th = dagger ? -thetas[k] : thetas[k];
    const double c = std::cos(th), s = std::sin(th);
    if (op.pair_double) {
      steps[n] = {2, c, s};
      step_op[n++] = k;
    } else {
      // sa_single: alpha rotation then beta rotation (osa.py:1009-1042); they commute
      steps[n] = {0, c, s};
      step_op[n++] = k;
      steps[n] = {1, c, s};
      step_op[n++] = k;
    }
  }
  *n_steps = n;
  return SQ_OK;
}

static int ups_apply_impl(sq_space* sp, sq_layout* lay, const double* thetas_host, int first, int last, int dagger,
                          double* state_dev, const PeerPtrs* peers, void* stream);
static int ups_apply_order(sq_space* sp, sq_layout* lay, const double* thetas_host, const std::vector<int>& order, int dagger,
                           double* state_dev, const PeerPtrs* peers, void* stream, int gauge_flags, int n_states = 1,
                           int64_t state_stride = 0);

extern "C" int sq_ups_apply(sq_space* sp, sq_layout* lay, const double* thetas_host, int first, int last, int dagger,
                            double* state_dev, void* stream) {
  return ups_apply_impl(sp, lay, thetas_host, first, last, dagger, state_dev, nullptr, stream);
}

// ---- alpha-sharded vectors (one process per GPU; shards peer-mapped over NVLink with CUDA IPC) ----
extern "C" int sq_space_set_partition(sq_space* sp, int world, int rank, const int64_t* row_starts) {
  if (!sp || !row_starts || world < 1 || world > SQ_MAX_WORLD || rank < 0 || rank >= world) return SQ_ERR_INVALID;
  if (row_starts[0] != 0 || row_starts[world] != sp->NA) {
    sq_set_error("sq_space_set_partition: row_starts must run from 0 to %lld", (long long)sp->NA);
    return SQ_ERR_INVALID;
  }
  for (int r = 0; r < world; ++r)
    if (row_starts[r] > row_starts[r + 1]) return SQ_ERR_INVALID;
  if (row_starts[rank] != sp->row_begin || row_starts[rank + 1] != sp->row_end) {
    sq_set_error("sq_space_set_partition: rank %d owns rows [%lld,%lld) but the space was created for [%lld,%lld)", rank,
                 (long long)row_starts[rank], (long long)row_starts[rank + 1], (long long)sp->row_begin,
                 (long long)sp->row_end);
    return SQ_ERR_INVALID;
  }
  sp->world = world;
  sp->rank = rank;
  sp->row_starts.assign(row_starts, row_starts + world + 1);
  return SQ_OK;
}

// Default partition: alpha strings grouped by the occupation of the first log2(world) orbitals.  In
// itertools.combinations order these groups are contiguous row ranges, and every orbital pair (p, p+1) with
// p >= log2(world) keeps both rows of a pair on the same device.
extern "C" int sq_partition_prefix(int n_orb, int n_alpha, int world, int64_t* row_starts_out) {
  if (!row_starts_out || world < 1 || (world & (world - 1)) != 0 || world > SQ_MAX_WORLD) {
    sq_set_error("sq_partition_prefix: world must be a power of two <= %d", SQ_MAX_WORLD);
    return SQ_ERR_INVALID;
  }
  int k = 0;
  while ((1 << k) < world) ++k;
  if (k > n_orb || n_orb < 1 || n_alpha < 0 || n_alpha > n_orb) {
    sq_set_error("sq_partition_prefix: need 0 <= n_alpha <= n_orb and log2(world) <= n_orb");
    return SQ_ERR_INVALID;
  }
  auto binom = [](int n, int r) -> int64_t {
    if (r < 0 || r > n) return 0;
    long double v = 1;
    for (int i = 1; i <= r; ++i) v = v * (n - r + i) / i;
    return (int64_t)(v + 0.5L);
  };
  // lexicographic order of sorted index tuples: prefix patterns with an occupied early orbital come first,
  // i.e. prefix bit patterns (orbital 0 = most significant) in DEcreasing numeric order
  int64_t acc = 0;
  row_starts_out[0] = 0;
  for (int g = 0; g < world; ++g) {
    const int pattern = world - 1 - g;
    const int occ = __builtin_popcount((unsigned)pattern);
    acc += binom(n_orb - k, n_alpha - occ);
    row_starts_out[g + 1] = acc;
  }
  return SQ_OK;
}

extern "C" int sq_dist_alloc(int device, int64_t n_doubles, double** out) {
  if (!out || n_doubles < 0) return SQ_ERR_INVALID;
  SQ_CUDA(cudaSetDevice(device));
  // at least one element so that the allocation has a valid IPC handle even for an empty shard
  SQ_CUDA(cudaMalloc(out, sizeof(double) * (size_t)(n_doubles > 0 ? n_doubles : 1)));
  return SQ_OK;
}

extern "C" int sq_dist_free(double* ptr) {
  if (ptr) SQ_CUDA(cudaFree(ptr));
  return SQ_OK;
}

extern "C" int sq_ipc_export(const double* ptr, unsigned char* handle64) {
  if (!ptr || !handle64) return SQ_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  SQ_CUDA(cudaIpcGetMemHandle(&h, const_cast<double*>(ptr)));
  memcpy(handle64, &h, 64);
  return SQ_OK;
}

extern "C" int sq_ipc_import(int device, const unsigned char* handle64, double** out) {
  if (!handle64 || !out) return SQ_ERR_INVALID;
  SQ_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  SQ_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *out = (double*)p;
  return SQ_OK;
}

extern "C" int sq_ipc_close(double* ptr) {
  if (ptr) SQ_CUDA(cudaIpcCloseMemHandle(ptr));
  return SQ_OK;
}

// 1 if some row pair of the runs in [first,last) spans two devices: every rank must then separate these
// operators from their neighbours by a device-wide barrier (same answer on every rank).
extern "C" int sq_layout_needs_exchange(const sq_layout* lay, int first, int last) {
  if (!lay || first < 0 || last > (int)lay->ops.size() || first > last) return -1;
  for (int k = first; k < last; ++k) {
    const LayoutOp& op = lay->ops[k];
    if (op.pair >= 0 && lay->pairs[op.pair].cross_global) return 1;
  }
  return 0;
}

extern "C" int sq_ups_apply_dist(sq_space* sp, sq_layout* lay, const double* thetas_host, int first, int last,
                                 int dagger, double* const* shard_ptrs_host, void* stream) {
  if (!sp || !shard_ptrs_host) return SQ_ERR_INVALID;
  PeerPtrs peers;
  for (int r = 0; r < SQ_MAX_WORLD; ++r) peers.p[r] = (r < sp->world) ? shard_ptrs_host[r] : nullptr;
  return ups_apply_impl(sp, lay, thetas_host, first, last, dagger, peers.p[sp->rank], &peers, stream);
}

static int ups_apply_impl(sq_space* sp, sq_layout* lay, const double* thetas_host, int first, int last, int dagger,
                          double* state_dev, const PeerPtrs* peers, void* stream) {
  if (!sp || !lay || lay->sp != sp || !state_dev) return SQ_ERR_INVALID;
  const int P = (int)lay->ops.size();
  if (first < 0 || last > P || first > last || (first < last && !thetas_host)) {
    sq_set_error("sq_ups_apply: bad operator range [%d,%d) for %d operators", first, last, P);
    return SQ_ERR_INVALID;
  }
  if (sp->device < 0) {
    sq_set_error("sq_ups_apply: host-only space (device = -1) cannot run kernels");
    return SQ_ERR_INVALID;
  }
  std::vector<int> order;
  exec_order(first, last, dagger, &order);
  return ups_apply_order(sp, lay, thetas_host, order, dagger, state_dev, peers, stream, 0);
}

// Operators `op_list` (indices into the layout) in the given EXECUTION order; dagger != 0 only negates the angles.  The caller
// vouches that the order is equivalent to the circuit order (operators that are moved past each other commute) -- this is how
// the re-sharding driver (slowquant_b200/distributed.py) runs the part of a circuit that is executable in the current row
// layout.  gauge_flags bit 0: the vector is in the sign-free gauge of the window kernel on entry; bit 1: leave it in that
// gauge on return (the gauge is a property of the determinant, not of the row layout, so it survives a re-shard).
extern "C" int sq_ups_apply_list(sq_space* sp, sq_layout* lay, const double* thetas_host, int n_list, const int32_t* op_list,
                                 int dagger, int gauge_flags, double* state_dev, void* stream) {
  if (!sp || !lay || lay->sp != sp || !state_dev || n_list < 0 || (n_list > 0 && (!op_list || !thetas_host))) return SQ_ERR_INVALID;
  if (sp->device < 0) {
    sq_set_error("sq_ups_apply_list: host-only space (device = -1) cannot run kernels");
    return SQ_ERR_INVALID;
  }
  const int P = (int)lay->ops.size();
  std::vector<int> order(op_list, op_list + n_list);
  std::vector<char> seen((size_t)P, 0);
  for (int k : order) {
    if (k < 0 || k >= P || seen[k]) {
      sq_set_error("sq_ups_apply_list: operator index %d out of range or repeated", k);
      return SQ_ERR_INVALID;
    }
    seen[k] = 1;
  }
  return ups_apply_order(sp, lay, thetas_host, order, dagger, state_dev, nullptr, stream, gauge_flags);
}

// The same circuit on a BATCH of vectors states_dev + s * state_stride, s < n_states (the *_SA twins of the reference,
// osa.py:1415-1864 / 2312-2754, and the common tail of RotoSolve's shifted states, ups_wavefunction.py:1183-1187): window sweeps
// and gauge sweeps take the batch as one launch whose CTAs stage their tables once for all states; the other kernels run per state.
extern "C" int sq_ups_apply_batch(sq_space* sp, sq_layout* lay, const double* thetas_host, int first, int last, int dagger,
                                  double* states_dev, int n_states, int64_t state_stride, void* stream) {
  if (!sp || !lay || lay->sp != sp || !states_dev || n_states < 1 || (n_states > 1 && state_stride < sp->local_len())) {
    sq_set_error("sq_ups_apply_batch: need n_states >= 1 and a stride of at least the vector length");
    return SQ_ERR_INVALID;
  }
  const int P = (int)lay->ops.size();
  if (first < 0 || last > P || first > last || (first < last && !thetas_host)) {
    sq_set_error("sq_ups_apply_batch: bad operator range [%d,%d) for %d operators", first, last, P);
    return SQ_ERR_INVALID;
  }
  if (sp->device < 0) {
    sq_set_error("sq_ups_apply_batch: host-only space (device = -1) cannot run kernels");
    return SQ_ERR_INVALID;
  }
  std::vector<int> order;
  exec_order(first, last, dagger, &order);
  return ups_apply_order(sp, lay, thetas_host, order, dagger, states_dev, nullptr, stream, 0, n_states, state_stride);
}

static int ups_apply_order(sq_space* sp, sq_layout* lay, const double* thetas_host, const std::vector<int>& order, int dagger,
                           double* state_dev0, const PeerPtrs* peers, void* stream, int gauge_flags, int n_states, int64_t state_stride) {
  SqRange nvtx_range("sq_ups_apply");
  cudaStream_t st = (cudaStream_t)stream;
  SQ_CUDA(cudaSetDevice(sp->device));
  for (int k : order) {
    const LayoutOp& op = lay->ops[k];
    if (std::fabs(thetas_host[k]) < 1e-28) continue;
    if (op.blocked || (op.pair >= 0 && lay->pairs[op.pair].blocked)) {
      sq_set_error("operator %d moves an alpha electron on a constrained orbital of this space (re-shard first)", k);
      return SQ_ERR_UNSUPPORTED;
    }
  }
  std::vector<std::vector<int>> runs;
  plan_runs(lay, order, thetas_host, &runs);
  std::vector<Launch> launches;
  SQ_CHECK(plan_launches(lay, runs, &launches));
  static const bool timing = getenv("SQ_LAUNCH_TIMING") != nullptr;   // debug: per-launch device time on stderr
  static cudaEvent_t tev0 = nullptr, tev1 = nullptr;   // created once (debug switch only)
  if (timing && !tev0) {
    cudaEventCreate(&tev0);
    cudaEventCreate(&tev1);
  }
  struct TimingScope {
    cudaEvent_t a, b;
    cudaStream_t st;
    const Launch* l;
    TimingScope(cudaEvent_t a_, cudaEvent_t b_, cudaStream_t st_, const Launch* l_) : a(a_), b(b_), st(st_), l(l_) {
      if (a) cudaEventRecord(a, st);
    }
    ~TimingScope() {
      if (!a) return;
      cudaEventRecord(b, st);
      cudaEventSynchronize(b);
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      if (l->kind == 2)
        fprintf(stderr, "launch win [%d,%d) smem=%zu grid=%dx%d bricks=%zu : %.3f ms\n", l->wt->w0, l->wt->w0 + l->wt->H,
                l->wt->smem, l->wt->n_ranges_b, l->wt->n_groups_a, l->runs.size(), ms);
      else
        fprintf(stderr, "launch kind=%d : %.3f ms\n", l->kind, ms);
    }
  };
  // window sweeps work in the sign-free gauge (sqsv_win.cu); every other kernel in the reference's sign convention
  bool in_gauge = (gauge_flags & 1) != 0;
  for (const Launch& l : launches) {
    if ((l.kind == 2) != in_gauge) {
      SQ_CHECK(sq_launch_gauge(sp, state_dev0, st, n_states, state_stride));
      in_gauge = !in_gauge;
    }
    TimingScope tscope(tev0, tev1, st, &l);
    auto& run = runs[l.runs[0]];
    const LayoutOp& op = lay->ops[run[0]];
    TileStep steps[SQ_MAX_PROGRAM];
    int step_op[SQ_MAX_PROGRAM], n_steps = 0;
    if (l.kind == 2) {
      // window sweep: the rotation steps of every brick; the launcher folds them into constant matrices
      int pair_idx[SQ_WIN_MAX_BRICKS], nst[SQ_WIN_MAX_BRICKS];
      TileStep wsteps[SQ_WIN_MAX_BRICKS][SQ_MAX_PROGRAM];
      const TileStep* sptr[SQ_WIN_MAX_BRICKS];
      int nb = 0;
      for (int t : l.runs) {
        SQ_CHECK(run_tile(sp, lay, runs[t], thetas_host, dagger, wsteps[nb], &nst[nb], step_op));
        pair_idx[nb] = lay->ops[runs[t][0]].pair;
        sptr[nb] = wsteps[nb];
        ++nb;
      }
      SQ_CHECK(sq_launch_win(sp, *l.wt, pair_idx, sptr, nst, nb, state_dev0, st, n_states, state_stride));
      continue;
    }
    for (int si = 0; si < n_states; ++si) {
    double* const state_dev = state_dev0 + (int64_t)si * state_stride;
    if (l.kind == 1) {
      // two bricks on disjoint orbital pairs commute: one sweep for both (sqsv_quad.cu)
      const int pA = op.pair, pB = lay->ops[runs[l.runs[1]][0]].pair;
      SQ_CHECK(run_tile(sp, lay, run, thetas_host, dagger, steps, &n_steps, step_op));
      TileStep steps2[SQ_MAX_PROGRAM];
      int step_op2[SQ_MAX_PROGRAM], n_steps2 = 0;
      SQ_CHECK(run_tile(sp, lay, runs[l.runs[1]], thetas_host, dagger, steps2, &n_steps2, step_op2));
      SQ_CHECK(sq_launch_quad(sp, *l.qt, steps, n_steps, lay->pairs[pA].sigma, steps2, n_steps2, lay->pairs[pB].sigma,
                              state_dev, st));
    } else if (is_tile_op(op)) {
      SQ_CHECK(run_tile(sp, lay, run, thetas_host, dagger, steps, &n_steps, step_op));
      SQ_CHECK(sq_launch_tile(sp, lay->pairs[op.pair], steps, n_steps, state_dev, peers, st));
    } else if (op.null_op) {
      break;
    } else if (op.gen >= 0) {
      const double th = dagger ? -thetas_host[run[0]] : thetas_host[run[0]];
      SQ_CHECK(sq_launch_gen_rot(sp, lay->gens[op.gen], std::cos(th), std::sin(th), state_dev, st));
    } else if (op.type >= SQ_EXC_SA_DOUBLE_1 && op.type <= SQ_EXC_SA_DOUBLE_5) {
      if (!op.multi) {
        sq_set_error("sa_double operator %d has no generator attached (sq_layout_attach_generator)", run[0]);
        return SQ_ERR_INVALID;
      }
      const double th = dagger ? -thetas_host[run[0]] : thetas_host[run[0]];
      SQ_CHECK(sa_double_poly(sp, *op.multi, op.type, th, state_dev, st));
    } else {
      sq_set_error("Got unknown excitation type code %d", op.type);
      return SQ_ERR_INVALID;
    }
    }   // states of the batch
  }
  if (in_gauge != ((gauge_flags & 2) != 0)) SQ_CHECK(sq_launch_gauge(sp, state_dev0, st, n_states, state_stride));
  return SQ_OK;
}

// TEST INFRASTRUCTURE (no product call reaches it): run the launch plan of operators [first,last) on a HOST vector, every window
// sweep through the host emulation of win3_kernel (sqsv_win3.cu) -- the same tables, step grouping, expanded item lists and block
// algebra the kernel uses.  The CPU tests compare the result with the oracle on host-only spaces.  A plan that contains anything
// but window sweeps is refused.
extern "C" int sq_debug_win3_emulate(sq_space* sp, sq_layout* lay, const double* thetas_host, int first, int last, int dagger,
                                     double* host_state) {
  if (!sp || !lay || lay->sp != sp || !host_state || !thetas_host) return SQ_ERR_INVALID;
  const int P = (int)lay->ops.size();
  if (first < 0 || last > P || first > last) return SQ_ERR_INVALID;
  std::vector<int> order;
  exec_order(first, last, dagger, &order);
  std::vector<std::vector<int>> runs;
  plan_runs(lay, order, thetas_host, &runs);
  std::vector<Launch> launches;
  SQ_CHECK(plan_launches(lay, runs, &launches));
  for (const Launch& l : launches)
    if (l.kind != 2) {
      sq_set_error("sq_debug_win3_emulate: the plan holds a launch that is not a window sweep");
      return SQ_ERR_UNSUPPORTED;
    }
  if (launches.empty()) return SQ_OK;
  sq_gauge_host(sp, host_state);
  for (const Launch& l : launches) {
    int pair_idx[SQ_WIN_MAX_BRICKS], nst[SQ_WIN_MAX_BRICKS], step_op[SQ_MAX_PROGRAM];
    TileStep wsteps[SQ_WIN_MAX_BRICKS][SQ_MAX_PROGRAM];
    const TileStep* sptr[SQ_WIN_MAX_BRICKS];
    int nb = 0;
    for (int t : l.runs) {
      SQ_CHECK(run_tile(sp, lay, runs[t], thetas_host, dagger, wsteps[nb], &nst[nb], step_op));
      pair_idx[nb] = lay->ops[runs[t][0]].pair;
      sptr[nb] = wsteps[nb];
      ++nb;
    }
    Win3Program P3;
    SQ_CHECK(sq_win3_program(*l.wt, pair_idx, sptr, nst, nb, &P3));
    SQ_CHECK(sq_win3_emulate_host(sp, *l.wt, P3, host_state));
  }
  sq_gauge_host(sp, host_state);
  return SQ_OK;
}

extern "C" int sq_grad_action(sq_space* sp, sq_layout* lay, int k, const double* in_dev, double* out_dev, void* stream) {
  SqRange nvtx_range("sq_grad_action");
  if (!sp || !lay || lay->sp != sp || !in_dev || !out_dev || k < 0 || k >= (int)lay->ops.size()) return SQ_ERR_INVALID;
  if (in_dev == out_dev) {
    sq_set_error("sq_grad_action: in and out must not alias");
    return SQ_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  SQ_CUDA(cudaSetDevice(sp->device));
  const LayoutOp& op = lay->ops[k];
  if (is_tile_op(op)) {
    // T|in> through the gather kernel with the generator's strings (Ta + Tb for sa_single, osa.py:2794-2806)
    const PairTables& pt = lay->pairs[op.pair];
    const int i = pt.i, a = pt.a;
    std::vector<std::vector<int32_t>> labels;
    std::vector<double> cf;
    if (op.pair_double) {
      labels.push_back({2 * (2 * a + 1) + 1, 2 * (2 * a) + 1, 2 * (2 * i + 1), 2 * (2 * i)});
      cf.push_back(-1.0);
      labels.push_back({2 * (2 * i + 1) + 1, 2 * (2 * i) + 1, 2 * (2 * a + 1), 2 * (2 * a)});
      cf.push_back(1.0);
    } else {
      labels.push_back({2 * (2 * a) + 1, 2 * (2 * i)});
      cf.push_back(1.0);
      labels.push_back({2 * (2 * i) + 1, 2 * (2 * a)});
      cf.push_back(-1.0);
      labels.push_back({2 * (2 * a + 1) + 1, 2 * (2 * i + 1)});
      cf.push_back(1.0);
      labels.push_back({2 * (2 * i + 1) + 1, 2 * (2 * a + 1)});
      cf.push_back(-1.0);
    }
    std::vector<StringAction> acts(labels.size());
    for (size_t s = 0; s < labels.size(); ++s)
      SQ_CHECK(sq_make_string_action(sp, labels[s].data(), (int)labels[s].size(), &acts[s]));
    return sq_launch_gather(sp, acts, cf, in_dev, out_dev, 0, st);
  }
  if (op.null_op) {
    SQ_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double) * (size_t)sp->local_len(), st));
    return SQ_OK;
  }
  if (op.gen >= 0) return sq_launch_gen_apply(sp, lay->gens[op.gen], in_dev, out_dev, st);
  if (op.multi) return sq_launch_gather(sp, op.multi->strings, op.multi->coeffs, in_dev, out_dev, 0, st);
  sq_set_error("sq_grad_action: operator %d has no generator", k);
  return SQ_ERR_INVALID;
}

static int grad_sweep_order(sq_space* sp, sq_layout* lay, const double* thetas_host, const std::vector<int>& order,
                            const std::vector<int>& out_pos, double* bra_dev, double* ket_dev, double* grad_host, void* stream,
                            int dagger);
static int grad_sweep_list_impl(sq_space* sp, sq_layout* lay, const double* thetas_host, int n_list, const int32_t* op_list,
                                double* bra_dev, double* ket_dev, double* grad_host, void* stream, int dagger);

extern "C" int sq_ups_grad_sweep(sq_space* sp, sq_layout* lay, const double* thetas_host, int first, int last,
                                 double* bra_dev, double* ket_dev, double* grad_host, void* stream) {
  if (!sp || !lay || lay->sp != sp || !bra_dev || !ket_dev || !grad_host) return SQ_ERR_INVALID;
  const int P = (int)lay->ops.size();
  if (first < 0 || last > P || first > last) return SQ_ERR_INVALID;
  std::vector<int> order, out_pos((size_t)P, -1);
  exec_order(first, last, 0, &order);
  for (int k = first; k < last; ++k) out_pos[k] = k - first;
  return grad_sweep_order(sp, lay, thetas_host, order, out_pos, bra_dev, ket_dev, grad_host, stream, 0);
}

// The same sweep over an explicit operator list in execution order (one phase of the re-sharding driver; the caller vouches that
// the order is equivalent to the circuit order -- operators that changed places commute, which leaves every <bra|T_k|ket>
// unchanged).  grad_host[i] belongs to operator op_list[i].
extern "C" int sq_ups_grad_sweep_list(sq_space* sp, sq_layout* lay, const double* thetas_host, int n_list, const int32_t* op_list,
                                      double* bra_dev, double* ket_dev, double* grad_host, void* stream) {
  return grad_sweep_list_impl(sp, lay, thetas_host, n_list, op_list, bra_dev, ket_dev, grad_host, stream, 0);
}

static int grad_sweep_list_impl(sq_space* sp, sq_layout* lay, const double* thetas_host, int n_list, const int32_t* op_list,
                                double* bra_dev, double* ket_dev, double* grad_host, void* stream, int dagger) {
  if (!sp || !lay || lay->sp != sp || !bra_dev || !ket_dev || n_list < 0 || (n_list > 0 && (!op_list || !grad_host || !thetas_host)))
    return SQ_ERR_INVALID;
  const int P = (int)lay->ops.size();
  std::vector<int> order, out_pos((size_t)P, -1);
  for (int i = 0; i < n_list; ++i) {
    const int k = op_list[i];
    if (k < 0 || k >= P || out_pos[k] >= 0) {
      sq_set_error("sq_ups_grad_sweep_list: operator index %d out of range or repeated", k);
      return SQ_ERR_INVALID;
    }
    const LayoutOp& op = lay->ops[k];
    if (op.blocked || (op.pair >= 0 && lay->pairs[op.pair].blocked)) {
      sq_set_error("operator %d moves an alpha electron on a constrained orbital of this space (re-shard first)", k);
      return SQ_ERR_UNSUPPORTED;
    }
    out_pos[k] = i;
    order.push_back(k);
  }
  return grad_sweep_order(sp, lay, thetas_host, order, out_pos, bra_dev, ket_dev, grad_host, stream, dagger);
}

// The sweep run BACKWARDS: op_list is in the execution order of the adjoint circuit; for every operator g = 2 <bra|T_k|ket> is
// taken first, then both vectors <- U_k^dagger.  Started from (H|psi>, |psi>) it yields the gradient of ups_wavefunction.py:1114-1138
// without the adjoint pass over the circuit (T_k commutes with its own rotation).
extern "C" int sq_ups_grad_sweep_list_rev(sq_space* sp, sq_layout* lay, const double* thetas_host, int n_list, const int32_t* op_list,
                                          double* bra_dev, double* ket_dev, double* grad_host, void* stream) {
  return grad_sweep_list_impl(sp, lay, thetas_host, n_list, op_list, bra_dev, ket_dev, grad_host, stream, 1);
}

// dagger = 1: `order` runs backwards through the circuit and every rotation is undone (theta -> -theta) after its gradient
// has been taken: the sweep of ups_wavefunction.py:1114-1138 started from (H|psi>, |psi>) instead of (U^d H|psi>, |ref>) --
// the same numbers <bra_k|T_k|ket_k> (T_k commutes with its own rotation), without the adjoint pass over the circuit.
static int grad_sweep_order(sq_space* sp, sq_layout* lay, const double* thetas_host, const std::vector<int>& order,
                            const std::vector<int>& out_pos, double* bra_dev, double* ket_dev, double* grad_host, void* stream,
                            int dagger) {
  SqRange nvtx_range("sq_ups_grad_sweep");
  const int P = (int)lay->ops.size();
  cudaStream_t st = (cudaStream_t)stream;
  SQ_CUDA(cudaSetDevice(sp->device));
  for (int k : order) grad_host[out_pos[k]] = 0.0;
  // zero-theta operators still contribute a gradient; only the rotation is skipped.  Plan with all
  // operators (thetas replaced by 1 for the planner), rotations with c=1,s=0 are exact identities.
  std::vector<double> plan_th(P, 1.0);
  std::vector<std::vector<int>> runs;
  plan_runs(lay, order, plan_th.data(), &runs);
  // every launch writes its <bra|T|ket> values into a device array; ONE copy back at the end of the sweep
  std::vector<int> slot_op;            // layout operator each device slot belongs to
  std::vector<int> run_slot0(runs.size(), -1);
  for (size_t ri = 0; ri < runs.size(); ++ri) {
    const LayoutOp& op = lay->ops[runs[ri][0]];
    if (is_tile_op(op)) {
      run_slot0[ri] = (int)slot_op.size();
      for (int k : runs[ri])
        for (int s = 0; s < tile_steps_of(lay->ops[k]); ++s) slot_op.push_back(k);
    } else if (!op.null_op && op.gen >= 0) {
      run_slot0[ri] = (int)slot_op.size();
      slot_op.push_back(runs[ri][0]);
    }
  }
  double* d_grad = nullptr;
  const int n_repl = sq_win_grad_replicas();   // the window kernel spreads its global atomics over replicas of the slot array
  const size_t n_slot = slot_op.size();
  if (n_slot) {
    SQ_CUDA(cudaMallocAsync(&d_grad, sizeof(double) * n_slot * n_repl, st));
    const cudaError_t e0 = cudaMemsetAsync(d_grad, 0, sizeof(double) * n_slot * n_repl, st);
    if (e0 != cudaSuccess) {
      cudaFreeAsync(d_grad, st);
      sq_set_error("sq_ups_grad_sweep: %s", cudaGetErrorString(e0));
      return SQ_ERR_CUDA;
    }
  }
  // the same launch plan as sq_ups_apply: bricks of a window sweep are differentiated and applied inside the sweep
  // (commuting bricks may be reordered: <bra|T_k|ket> does not change); quad launches run as two single bricks
  std::vector<Launch> launches;
  int status = SQ_OK;
  if (g_win_grad) {
    status = plan_launches(lay, runs, &launches, true);
  } else {
    // default: one fused brick per launch (tile_grad_kernel_v2 runs at 0.7 of the HBM roofline; the window gradient
    // kernel is correct but not yet faster per brick)
    // two consecutive bricks on disjoint orbital pairs share one launch (quad_grad_kernel: one read + one write of bra and ket
    // for both bricks; every half layer of a tUPS brick wall is such a set)
    for (size_t ri = 0; ri < runs.size();) {
      Launch l;
      l.runs = {(int)ri};
      if (g_quad_grad && quad_enabled() && ri + 1 < runs.size()) {
        const LayoutOp &o1 = lay->ops[runs[ri][0]], &o2 = lay->ops[runs[ri + 1][0]];
        if (is_tile_op(o1) && is_tile_op(o2) && o1.pair != o2.pair) {
          const QuadTables* qt = nullptr;
          status = get_quad(lay, o1.pair, o2.pair, &qt);
          if (status != SQ_OK) break;
          if (qt && qt->ok) {
            l.kind = 1;
            l.runs.push_back((int)ri + 1);
            l.qt = qt;
          }
        }
      }
      ri += l.runs.size();
      launches.push_back(l);
    }
  }
  bool in_gauge = false;
  auto set_gauge = [&](bool want) -> int {
    if (want == in_gauge) return SQ_OK;
    SQ_CHECK(sq_launch_gauge(sp, bra_dev, st));
    SQ_CHECK(sq_launch_gauge(sp, ket_dev, st));
    in_gauge = want;
    return SQ_OK;
  };
  auto tile_program = [&](const std::vector<int>& run, TileStep* steps, int* n_steps) -> int {
    int step_op[SQ_MAX_PROGRAM];
    SQ_CHECK(run_tile(sp, lay, run, thetas_host, dagger, steps, n_steps, step_op));
    for (int s = 0; s < *n_steps; ++s)
      if (std::fabs(thetas_host[step_op[s]]) < 1e-28) { steps[s].c = 1.0; steps[s].s = 0.0; }
    return SQ_OK;
  };
  for (const Launch& l : launches) {
    if (status != SQ_OK) break;
    if (l.kind == 2) {
      status = set_gauge(true);
      if (status != SQ_OK) break;
      int pair_idx[SQ_WIN_MAX_BRICKS], nst[SQ_WIN_MAX_BRICKS], slot0[SQ_WIN_MAX_BRICKS];
      TileStep wsteps[SQ_WIN_MAX_BRICKS][SQ_MAX_PROGRAM];
      const TileStep* sptr[SQ_WIN_MAX_BRICKS];
      int nb = 0;
      for (int t : l.runs) {
        status = tile_program(runs[t], wsteps[nb], &nst[nb]);
        if (status != SQ_OK) break;
        pair_idx[nb] = lay->ops[runs[t][0]].pair;
        slot0[nb] = run_slot0[t];
        sptr[nb] = wsteps[nb];
        ++nb;
      }
      if (status == SQ_OK) status = sq_launch_win_grad(sp, *l.wt, pair_idx, sptr, nst, slot0, nb, bra_dev, ket_dev, d_grad, (int)n_slot, st);
      continue;
    }
    status = set_gauge(false);
    if (status != SQ_OK) break;
    if (l.kind == 1) {
      TileStep s1[SQ_MAX_PROGRAM], s2[SQ_MAX_PROGRAM];
      int n1 = 0, n2 = 0;
      const int t1 = l.runs[0], t2 = l.runs[1];
      status = tile_program(runs[t1], s1, &n1);
      if (status == SQ_OK) status = tile_program(runs[t2], s2, &n2);
      if (status == SQ_OK)
        status = sq_launch_quad_grad(sp, *l.qt, s1, n1, lay->pairs[lay->ops[runs[t1][0]].pair].sigma, s2, n2,
                                     lay->pairs[lay->ops[runs[t2][0]].pair].sigma, bra_dev, ket_dev, d_grad + run_slot0[t1],
                                     d_grad + run_slot0[t2], st);
      continue;
    }
    for (int t : l.runs) {
      if (status != SQ_OK) break;
      const std::vector<int>& run = runs[t];
      const LayoutOp& op = lay->ops[run[0]];
      if (is_tile_op(op)) {
        TileStep steps[SQ_MAX_PROGRAM];
        int n_steps = 0;
        status = tile_program(run, steps, &n_steps);
        if (status != SQ_OK) break;
        status = sq_launch_tile_grad(sp, lay->pairs[op.pair], steps, n_steps, bra_dev, ket_dev, d_grad + run_slot0[t], st);
      } else if (op.null_op) {
        continue;
      } else if (op.gen >= 0) {
        const int k = run[0];
        double th = dagger ? -thetas_host[k] : thetas_host[k];
        double c = std::cos(th), s = std::sin(th);
        if (std::fabs(th) < 1e-28) { c = 1.0; s = 0.0; }
        status = sq_launch_gen_grad(sp, lay->gens[op.gen], c, s, bra_dev, ket_dev, d_grad + run_slot0[t], st);
      } else if (op.multi) {
        const int k = run[0];
        status = sq_ensure_work(sp, 2);
        if (status == SQ_OK) status = sq_launch_gather(sp, op.multi->strings, op.multi->coeffs, ket_dev, sp->d_work[2], 0, st);
        double g = 0.0;
        if (status == SQ_OK) status = sq_launch_dot(sp, bra_dev, sp->d_work[2], &g, st);
        grad_host[out_pos[k]] = 2.0 * g;
        if (status == SQ_OK && std::fabs(thetas_host[k]) >= 1e-28) {
          const double th = dagger ? -thetas_host[k] : thetas_host[k];
          status = sa_double_poly(sp, *op.multi, op.type, th, bra_dev, st);
          if (status == SQ_OK) status = sa_double_poly(sp, *op.multi, op.type, th, ket_dev, st);
        }
      } else {
        sq_set_error("sq_ups_grad_sweep: operator %d has no generator", run[0]);
        status = SQ_ERR_INVALID;
      }
    }
  }
  if (status == SQ_OK) status = set_gauge(false);
  if (status == SQ_OK && n_slot) {
    std::vector<double> g(n_slot * n_repl);
    cudaError_t e = cudaMemcpyAsync(g.data(), d_grad, sizeof(double) * g.size(), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      sq_set_error("sq_ups_grad_sweep: %s", cudaGetErrorString(e));
      status = SQ_ERR_CUDA;
    } else {
      for (size_t i = 0; i < n_slot; ++i) {
        double v = 0.0;
        for (int r = 0; r < n_repl; ++r) v += g[(size_t)r * n_slot + i];   // replicas in a fixed order
        grad_host[out_pos[slot_op[i]]] += 2.0 * v;
      }
    }
  }
  if (d_grad) cudaFreeAsync(d_grad, st);
  return status;
}


// Gradient sweep over operators [first,last) of an ALPHA-SHARDED (bra, ket) pair whose row pairs may live on two GPUs (an exchange
// stretch of the plan: sa_single / pair-double operators only).  Same arithmetic as sq_ups_grad_sweep, one fused launch per brick;
// cross-device tiles are read and written in place through the peer mappings.  grad_host receives THIS rank's partial
// 2 <bra|T_k|ket> (the tiles it processed): the caller adds the ranks' results and puts a device-wide barrier before and after.
// Written without GPU time (compiled, not yet run); the Python side uses it only on request (peer_gradient=True).
extern "C" int sq_ups_grad_sweep_dist(sq_space* sp, sq_layout* lay, const double* thetas_host, int first, int last,
                                      double* const* bra_ptrs_host, double* const* ket_ptrs_host, double* grad_host, void* stream) {
  if (!sp || !lay || lay->sp != sp || !bra_ptrs_host || !ket_ptrs_host || !grad_host) return SQ_ERR_INVALID;
  const int P = (int)lay->ops.size();
  if (first < 0 || last > P || first > last) return SQ_ERR_INVALID;
  if (sp->device < 0 || sp->world < 1 || sp->world > SQ_MAX_WORLD) return SQ_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  SQ_CUDA(cudaSetDevice(sp->device));
  for (int k = first; k < last; ++k) grad_host[k - first] = 0.0;
  double* bra_dev = bra_ptrs_host[sp->rank];
  double* ket_dev = ket_ptrs_host[sp->rank];
  if (!bra_dev || !ket_dev) return SQ_ERR_INVALID;
  std::vector<double> plan_th(P, 1.0);
  std::vector<int> order;
  exec_order(first, last, 0, &order);
  std::vector<std::vector<int>> runs;
  plan_runs(lay, order, plan_th.data(), &runs);
  std::vector<int> slot_op;
  std::vector<int> run_slot0(runs.size(), -1);
  for (size_t ri = 0; ri < runs.size(); ++ri) {
    if (!is_tile_op(lay->ops[runs[ri][0]])) {
      sq_set_error("sq_ups_grad_sweep_dist: operator %d is not a brick operator (sa_single / pair double)", runs[ri][0]);
      return SQ_ERR_UNSUPPORTED;
    }
    run_slot0[ri] = (int)slot_op.size();
    for (int k : runs[ri])
      for (int s = 0; s < tile_steps_of(lay->ops[k]); ++s) slot_op.push_back(k);
  }
  if (slot_op.empty()) return SQ_OK;
  // device copies of the two base-pointer tables (pageable source: the copy is staged before the call returns)
  unsigned long long tabs[2 * SQ_MAX_WORLD];
  for (int r = 0; r < SQ_MAX_WORLD; ++r) {
    tabs[r] = r < sp->world ? (unsigned long long)(uintptr_t)bra_ptrs_host[r] : 0ull;
    tabs[SQ_MAX_WORLD + r] = r < sp->world ? (unsigned long long)(uintptr_t)ket_ptrs_host[r] : 0ull;
  }
  unsigned long long* d_tabs = nullptr;
  double* d_grad = nullptr;
  int status = SQ_OK;
  {
    // both buffers are released on every exit path (the frees at the end of the function see whatever was allocated)
    cudaError_t e = cudaMallocAsync(&d_tabs, sizeof(tabs), st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tabs, tabs, sizeof(tabs), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_grad, sizeof(double) * slot_op.size(), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_grad, 0, sizeof(double) * slot_op.size(), st);
    if (e != cudaSuccess) {
      sq_set_error("sq_ups_grad_sweep_dist: %s", cudaGetErrorString(e));
      status = SQ_ERR_CUDA;
    }
  }
  for (size_t ri = 0; ri < runs.size() && status == SQ_OK; ++ri) {
    const std::vector<int>& run = runs[ri];
    const LayoutOp& op = lay->ops[run[0]];
    TileStep steps[SQ_MAX_PROGRAM];
    int n_steps = 0, step_op[SQ_MAX_PROGRAM];
    status = run_tile(sp, lay, run, thetas_host, 0, steps, &n_steps, step_op);
    if (status != SQ_OK) break;
    for (int s = 0; s < n_steps; ++s)   // zero angles: the gradient is taken, the rotation is the identity
      if (std::fabs(thetas_host[step_op[s]]) < 1e-28) { steps[s].c = 1.0; steps[s].s = 0.0; }
    const PairTables& pt = lay->pairs[op.pair];
    if (pt.n_cross_items > 0)
      status = sq_launch_tile_grad_peer(sp, pt, steps, n_steps, bra_dev, ket_dev, d_tabs, d_tabs + SQ_MAX_WORLD, d_grad + run_slot0[ri], st);
    else
      status = sq_launch_tile_grad(sp, pt, steps, n_steps, bra_dev, ket_dev, d_grad + run_slot0[ri], st);
  }
  if (status == SQ_OK) {
    std::vector<double> g(slot_op.size());
    cudaError_t e = cudaMemcpyAsync(g.data(), d_grad, sizeof(double) * g.size(), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      sq_set_error("sq_ups_grad_sweep_dist: %s", cudaGetErrorString(e));
      status = SQ_ERR_CUDA;
    } else {
      for (size_t i = 0; i < g.size(); ++i) grad_host[slot_op[i] - first] += 2.0 * g[i];
    }
  }
  if (d_grad) cudaFreeAsync(d_grad, st);
  if (d_tabs) cudaFreeAsync(d_tabs, st);
  return status;
}

// ---------------------------------------------------------------------------------------------
// generic operators and BLAS-1
// ---------------------------------------------------------------------------------------------
extern "C" int sq_apply_strings(sq_space* sp, int n_strings, const int32_t* ops_flat, const int32_t* op_offsets,
                                const double* coeffs, const double* in_dev, double* out_dev, int accumulate,
                                int skip_outside, void* stream) {
  SqRange nvtx_range("sq_apply_strings");
  if (!sp || n_strings < 0 || !in_dev || !out_dev || (n_strings > 0 && (!ops_flat || !op_offsets || !coeffs)))
    return SQ_ERR_INVALID;
  if (in_dev == out_dev) {
    sq_set_error("sq_apply_strings: in and out must not alias");
    return SQ_ERR_INVALID;
  }
  if (sp->device < 0) {
    sq_set_error("sq_apply_strings: host-only space (device = -1) cannot run kernels");
    return SQ_ERR_INVALID;
  }
  SQ_CUDA(cudaSetDevice(sp->device));
  std::vector<StringAction> acts;
  std::vector<double> cf;
  SQ_CHECK(parse_strings(sp, n_strings, ops_flat, op_offsets, coeffs, skip_outside, &acts, &cf));
  return sq_launch_gather(sp, acts, cf, in_dev, out_dev, accumulate, (cudaStream_t)stream);
}

// Energy and theta gradient in one call (ups_wavefunction.py:1019-1142 with the state path of :345-365):
//   |psi> = U(theta)|ref>,  E = <psi|H|psi>,  g_k = 2 <bra_k|T_k|ket_k>  (reverse sweep of :1114-1138 started from
//   bra = U^d H|psi>, ket = |ref>).  The caller supplies two work vectors of the state's length.
extern "C" int sq_ups_energy_grad(sq_space* sp, sq_layout* lay, const double* thetas_host, double e_core,
                                  const double* h_act_host, const double* g_act_host, const double* ref_dev,
                                  double* work_ket_dev, double* work_bra_dev, double* energy_host, double* grad_host,
                                  void* stream) {
  if (!sp || !lay || lay->sp != sp || !thetas_host || !h_act_host || !g_act_host || !ref_dev || !work_ket_dev ||
      !work_bra_dev || !energy_host)
    return SQ_ERR_INVALID;
  if (work_ket_dev == work_bra_dev || work_ket_dev == ref_dev || work_bra_dev == ref_dev) {
    sq_set_error("sq_ups_energy_grad: the reference state and the two work vectors must be distinct");
    return SQ_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  SQ_CUDA(cudaSetDevice(sp->device));
  const int P = (int)lay->ops.size();
  const size_t bytes = sizeof(double) * (size_t)sp->local_len();
  SQ_CUDA(cudaMemcpyAsync(work_ket_dev, ref_dev, bytes, cudaMemcpyDeviceToDevice, st));
  SQ_CHECK(sq_ups_apply(sp, lay, thetas_host, 0, P, 0, work_ket_dev, stream));                       // |psi>
  SQ_CHECK(sq_sigma(sp, e_core, h_act_host, g_act_host, work_ket_dev, work_bra_dev, stream));        // H|psi>
  SQ_CHECK(sq_dot(sp, work_ket_dev, work_bra_dev, energy_host, stream));
  if (!grad_host) return SQ_OK;
  // backwards through the circuit from (H|psi>, |psi>): no adjoint pass (one state construction less per evaluation)
  std::vector<int> order, out_pos((size_t)P, -1);
  exec_order(0, P, 1, &order);
  for (int k = 0; k < P; ++k) out_pos[k] = k;
  return grad_sweep_order(sp, lay, thetas_host, order, out_pos, work_bra_dev, work_ket_dev, grad_host, stream, 1);
}

extern "C" int sq_dot(sq_space* sp, const double* a_dev, const double* b_dev, double* out_host, void* stream) {
  if (!sp || !a_dev || !b_dev || !out_host) return SQ_ERR_INVALID;
  SQ_CUDA(cudaSetDevice(sp->device));
  if (sp->local_len() == 0) {
    *out_host = 0.0;
    return SQ_OK;
  }
  return sq_launch_dot(sp, a_dev, b_dev, out_host, (cudaStream_t)stream);
}

extern "C" int sq_axpy(sq_space* sp, double alpha, const double* x_dev, double* y_dev, void* stream) {
  if (!sp || !x_dev || !y_dev) return SQ_ERR_INVALID;
  SQ_CUDA(cudaSetDevice(sp->device));
  return sq_launch_axpy(sp, alpha, x_dev, y_dev, (cudaStream_t)stream);
}

extern "C" int sq_scale_copy(sq_space* sp, double alpha, const double* x_dev, double* y_dev, void* stream) {
  if (!sp || !x_dev || !y_dev) return SQ_ERR_INVALID;
  SQ_CUDA(cudaSetDevice(sp->device));
  return sq_launch_scale_copy(sp, alpha, x_dev, y_dev, (cudaStream_t)stream);
}
