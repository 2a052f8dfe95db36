// See topology.h.  Restates, once per (pairing, n), what the reference rebuilds for every
// coarse cell: DoF enumeration (ned_rt_basis.cc:181-222), boundary data of the k local
// problems (:225-359, via include/functions/basis_*.tpp closed forms on a cube), the
// QGauss<3>(2) local-matrix structure (:463-513) and the B^T A B operator (:850-948).
#include "topology.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <map>
#include <stdexcept>

#include "../../include/msfec.h"

namespace msfec {

void gauss_points(double qp[8][3]) {
  const double g[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
  for (int q = 0; q < 8; ++q) {
    qp[q][0] = g[q & 1];
    qp[q][1] = g[(q >> 1) & 1];
    qp[q][2] = g[q >> 2];
  }
}

namespace {

inline double w(int bit, double s) { return bit ? s : 1.0 - s; }
inline double dw(int bit) { return bit ? 1.0 : -1.0; }

// deal.II GeometryInfo<3> line table: direction and the corner bits on the two
// transverse axes.
struct LineDef { int d; int a0, b0, a1, b1; };   // transverse axes a0<a1 with bits b0,b1
const LineDef kLines[12] = {
    {1, 0, 0, 2, 0}, {1, 0, 1, 2, 0}, {0, 1, 0, 2, 0}, {0, 1, 1, 2, 0},
    {1, 0, 0, 2, 1}, {1, 0, 1, 2, 1}, {0, 1, 0, 2, 1}, {0, 1, 1, 2, 1},
    {2, 0, 0, 1, 0}, {2, 0, 1, 1, 0}, {2, 0, 0, 1, 1}, {2, 0, 1, 1, 1}};
struct FaceDef { int d, s; };
const FaceDef kFaces[6] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}, {2, 0}, {2, 1}};

// Reference shape functions on the unit cube at xi.
void q1_ref(const double xi[3], double val[8], double grad[8][3]) {
  for (int v = 0; v < 8; ++v) {
    const int b[3] = {v & 1, (v >> 1) & 1, v >> 2};
    const double ww[3] = {w(b[0], xi[0]), w(b[1], xi[1]), w(b[2], xi[2])};
    val[v] = ww[0] * ww[1] * ww[2];
    grad[v][0] = dw(b[0]) * ww[1] * ww[2];
    grad[v][1] = ww[0] * dw(b[1]) * ww[2];
    grad[v][2] = ww[0] * ww[1] * dw(b[2]);
  }
}
void ned_ref(const double xi[3], double val[12][3], double curl[12][3]) {
  for (int l = 0; l < 12; ++l) {
    const LineDef &L = kLines[l];
    const double f = w(L.b0, xi[L.a0]) * w(L.b1, xi[L.a1]);
    double g[3] = {0, 0, 0};
    g[L.a0] = dw(L.b0) * w(L.b1, xi[L.a1]);
    g[L.a1] = w(L.b0, xi[L.a0]) * dw(L.b1);
    for (int c = 0; c < 3; ++c) val[l][c] = 0.0;
    val[l][L.d] = f;
    double e[3] = {0, 0, 0};
    e[L.d] = 1.0;   // curl(f e_d) = grad f x e_d
    curl[l][0] = g[1] * e[2] - g[2] * e[1];
    curl[l][1] = g[2] * e[0] - g[0] * e[2];
    curl[l][2] = g[0] * e[1] - g[1] * e[0];
  }
}
void rt_ref(const double xi[3], double val[6][3], double div[6]) {
  for (int f = 0; f < 6; ++f) {
    for (int c = 0; c < 3; ++c) val[f][c] = 0.0;
    val[f][kFaces[f].d] = w(kFaces[f].s, xi[kFaces[f].d]);
    div[f] = dw(kFaces[f].s);
  }
}

// Per-kind local tables at the 8 Gauss points (unit h).
struct LocalTables {
  int ldofs;
  double vec[12][8][3];   // V: grad, E: value, F: value
  double vec2[12][8][3];  // E: curl
  double sc[12][8];       // V: value, F: div, C: 1
};

LocalTables local_tables(int kind) {
  LocalTables t{};
  double qp[8][3];
  gauss_points(qp);
  for (int q = 0; q < 8; ++q) {
    if (kind == ENT_V) {
      t.ldofs = 8;
      double v[8], g[8][3];
      q1_ref(qp[q], v, g);
      for (int i = 0; i < 8; ++i) { t.sc[i][q] = v[i]; for (int c = 0; c < 3; ++c) t.vec[i][q][c] = g[i][c]; }
    } else if (kind == ENT_E) {
      t.ldofs = 12;
      double v[12][3], cu[12][3];
      ned_ref(qp[q], v, cu);
      for (int i = 0; i < 12; ++i) for (int c = 0; c < 3; ++c) { t.vec[i][q][c] = v[i][c]; t.vec2[i][q][c] = cu[i][c]; }
    } else if (kind == ENT_F) {
      t.ldofs = 6;
      double v[6][3], d[6];
      rt_ref(qp[q], v, d);
      for (int i = 0; i < 6; ++i) { t.sc[i][q] = d[i]; for (int c = 0; c < 3; ++c) t.vec[i][q][c] = v[i][c]; }
    } else {
      t.ldofs = 1;
      t.sc[0][q] = 1.0;
    }
  }
  return t;
}

BlockTopo make_block(int kind, int n) {
  BlockTopo b;
  b.kind = kind;
  const int n1 = n + 1, nC = n * n * n;
  std::vector<double> pos;
  std::vector<int32_t> axis;
  std::vector<int32_t> nat_cell_dofs;
  auto add_grid = [&](int nx, int ny, int nz, double ox, double oy, double oz, int ax) {
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      pos.push_back(i + ox); pos.push_back(j + oy); pos.push_back(k + oz);
      axis.push_back(ax);
    }
  };
  const int nEx = n * n1 * n1, nFx = n1 * n * n;
  if (kind == ENT_V) { b.ldofs = 8; add_grid(n1, n1, n1, 0, 0, 0, -1); }
  if (kind == ENT_E) {
    b.ldofs = 12;
    add_grid(n, n1, n1, 0.5, 0, 0, 0); add_grid(n1, n, n1, 0, 0.5, 0, 1); add_grid(n1, n1, n, 0, 0, 0.5, 2);
  }
  if (kind == ENT_F) {
    b.ldofs = 6;
    add_grid(n1, n, n, 0, 0.5, 0.5, 0); add_grid(n, n1, n, 0.5, 0, 0.5, 1); add_grid(n, n, n1, 0.5, 0.5, 0, 2);
  }
  if (kind == ENT_C) { b.ldofs = 1; add_grid(n, n, n, 0.5, 0.5, 0.5, -1); }
  b.n_total = (int)axis.size();
  auto V = [&](int i, int j, int k) { return i + n1 * (j + n1 * k); };
  auto EX = [&](int i, int j, int k) { return i + n * (j + n1 * k); };
  auto EY = [&](int i, int j, int k) { return nEx + i + n1 * (j + n * k); };
  auto EZ = [&](int i, int j, int k) { return 2 * nEx + i + n1 * (j + n1 * k); };
  auto FX = [&](int i, int j, int k) { return i + n1 * (j + n * k); };
  auto FY = [&](int i, int j, int k) { return nFx + i + n * (j + n1 * k); };
  auto FZ = [&](int i, int j, int k) { return 2 * nFx + i + n * (j + n * k); };
  nat_cell_dofs.resize((size_t)nC * b.ldofs);
  for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
    int32_t *d = &nat_cell_dofs[(size_t)(i + n * (j + n * k)) * b.ldofs];
    if (kind == ENT_V) for (int v = 0; v < 8; ++v) d[v] = V(i + (v & 1), j + ((v >> 1) & 1), k + (v >> 2));
    if (kind == ENT_E) {
      d[0] = EY(i, j, k); d[1] = EY(i + 1, j, k); d[2] = EX(i, j, k); d[3] = EX(i, j + 1, k);
      d[4] = EY(i, j, k + 1); d[5] = EY(i + 1, j, k + 1); d[6] = EX(i, j, k + 1); d[7] = EX(i, j + 1, k + 1);
      d[8] = EZ(i, j, k); d[9] = EZ(i + 1, j, k); d[10] = EZ(i, j + 1, k); d[11] = EZ(i + 1, j + 1, k);
    }
    if (kind == ENT_F) {
      d[0] = FX(i, j, k); d[1] = FX(i + 1, j, k); d[2] = FY(i, j, k); d[3] = FY(i, j + 1, k);
      d[4] = FZ(i, j, k); d[5] = FZ(i, j, k + 1);
    }
    if (kind == ENT_C) d[0] = i + n * (j + n * k);
  }
  // boundary flags: vertices on dK; edges lying in dK (a transverse coordinate is 0 or n);
  // faces in dK (normal coordinate is 0 or n); cells never.
  std::vector<uint8_t> bnd(b.n_total, 0);
  for (int e = 0; e < b.n_total; ++e) {
    bool on = false;
    for (int c = 0; c < 3; ++c) {
      const double p = pos[3 * e + c];
      const bool at = (p == 0.0 || p == (double)n);
      if (kind == ENT_V) on |= at;
      if (kind == ENT_E && c != axis[e]) on |= at;
      if (kind == ENT_F && c == axis[e]) on |= at;
    }
    bnd[e] = on;
  }
  // interior-first permutation
  std::vector<int32_t> perm(b.n_total);
  int ni = 0;
  for (int e = 0; e < b.n_total; ++e) if (!bnd[e]) perm[e] = ni++;
  b.n_int = ni;
  int nb = 0;
  for (int e = 0; e < b.n_total; ++e) if (bnd[e]) perm[e] = ni + nb++;
  b.pos.resize(pos.size()); b.axis.resize(b.n_total); b.bnd.resize(b.n_total);
  for (int e = 0; e < b.n_total; ++e) {
    for (int c = 0; c < 3; ++c) b.pos[3 * perm[e] + c] = pos[3 * e + c];
    b.axis[perm[e]] = axis[e];
    b.bnd[perm[e]] = bnd[e];
  }
  b.cell_dofs.resize(nat_cell_dofs.size());
  for (size_t i = 0; i < nat_cell_dofs.size(); ++i) b.cell_dofs[i] = perm[nat_cell_dofs[i]];
  return b;
}

// ---- bilinear-form descriptors ------------------------------------------------------
enum FormType { FORM_TENSOR, FORM_SCALAR };

// coefficient channels per (fine cell, q): 0..5 = sym tensor (xx,xy,xz,yy,yz,zz), 6 = scalar
constexpr int kCoefPerQ = 7;
const int kSymIdx[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};

// Builds the slot numbering (unique unordered DoF pairs sharing a fine cell) and the
// gather-assembly table of a symmetric block form on block `B`.
void build_sym_form(const BlockTopo &B, int nC, FormType type, const double (*vec)[8][3],
                    const double (*sc)[8], AsmTable &tab, std::map<std::pair<int, int>, int> &slot_of) {
  const int L = B.ldofs;
  // pair lists for every local (i, j)
  tab.coef_stride = 8 * kCoefPerQ;
  tab.pair_ptr.assign(1, 0);
  for (int i = 0; i < L; ++i) for (int j = 0; j < L; ++j) {
    for (int q = 0; q < 8; ++q) {
      if (type == FORM_SCALAR) {
        const double wt = sc[i][q] * sc[j][q];
        if (wt != 0.0) { tab.pair_idx.push_back(q * kCoefPerQ + 6); tab.pair_w.push_back(wt); }
      } else {
        for (int a = 0; a < 3; ++a) for (int b = a; b < 3; ++b) {
          double wt = vec[i][q][a] * vec[j][q][b];
          if (a != b) wt += vec[i][q][b] * vec[j][q][a];
          if (wt != 0.0) { tab.pair_idx.push_back(q * kCoefPerQ + kSymIdx[a][b]); tab.pair_w.push_back(wt); }
        }
      }
    }
    tab.pair_ptr.push_back((int32_t)tab.pair_idx.size());
  }
  // slots
  std::vector<std::vector<std::pair<int, int>>> contribs;
  for (int T = 0; T < nC; ++T) {
    const int32_t *d = &B.cell_dofs[(size_t)T * L];
    for (int i = 0; i < L; ++i) for (int j = 0; j < L; ++j) {
      const int r = d[i], c = d[j];
      if (r > c) continue;
      if (tab.pair_ptr[i * L + j + 1] == tab.pair_ptr[i * L + j]) continue;   // structurally zero
      auto key = std::make_pair(r, c);
      auto it = slot_of.find(key);
      int s;
      if (it == slot_of.end()) { s = (int)contribs.size(); slot_of[key] = s; contribs.emplace_back(); }
      else s = it->second;
      contribs[s].push_back({T, i * L + j});
    }
  }
  tab.n_slots = (int)contribs.size();
  tab.contrib_ptr.assign(1, 0);
  for (auto &cl : contribs) {
    for (auto &c : cl) { tab.contrib_cell.push_back(c.first); tab.contrib_pair.push_back(c.second); }
    tab.contrib_ptr.push_back((int32_t)tab.contrib_cell.size());
  }
}

struct Triplet { int r, c, ref; double v; bool shared; };

RefOperator make_operator(int n_rows, int n_cols, std::vector<Triplet> &tr) {
  RefOperator op;
  op.n_rows = n_rows; op.n_cols = n_cols;
  std::stable_sort(tr.begin(), tr.end(), [](const Triplet &a, const Triplet &b) {
    return a.r != b.r ? a.r < b.r : a.c < b.c; });
  op.cptr.assign(n_rows + 1, 0); op.sptr.assign(n_rows + 1, 0);
  for (auto &t : tr) (t.shared ? op.sptr : op.cptr)[t.r + 1]++;
  for (int r = 0; r < n_rows; ++r) { op.cptr[r + 1] += op.cptr[r]; op.sptr[r + 1] += op.sptr[r]; }
  op.ccol.resize(op.cptr[n_rows]); op.cref.resize(op.cptr[n_rows]);
  op.scol.resize(op.sptr[n_rows]); op.sval.resize(op.sptr[n_rows]);
  std::vector<int32_t> cp(op.cptr.begin(), op.cptr.end() - 1), spp(op.sptr.begin(), op.sptr.end() - 1);
  for (auto &t : tr) {
    if (t.shared) { op.scol[spp[t.r]] = t.c; op.sval[spp[t.r]++] = t.v; }
    else { op.ccol[cp[t.r]] = t.c; op.cref[cp[t.r]++] = t.ref; }
  }
  return op;
}

}  // namespace

Topology build_topology(int pairing, int n) {
  if (n < 1 || (n & (n - 1))) throw std::runtime_error("n must be a power of two");
  Topology t;
  t.pairing = pairing; t.n = n; t.nC = n * n * n;
  int kind0, kind1 = -1;
  switch (pairing) {
    case MSFEC_Q: kind0 = ENT_V; t.k_solve = 8; t.k_gram = 8; t.k0 = 8; break;
    case MSFEC_Q_NED: kind0 = ENT_V; kind1 = ENT_E; t.k_solve = 20; t.k_gram = 20; t.k0 = 8; break;
    case MSFEC_NED_RT: kind0 = ENT_E; kind1 = ENT_F; t.k_solve = 18; t.k_gram = 18; t.k0 = 12; break;
    case MSFEC_RT_DQ: kind0 = ENT_F; kind1 = ENT_C; t.k_solve = 6; t.k_gram = 7; t.k0 = 6; break;
    default: throw std::runtime_error("unknown pairing");
  }
  t.two_blocks = kind1 >= 0;
  t.blk[0] = make_block(kind0, n);
  if (t.two_blocks) t.blk[1] = make_block(kind1, n);
  const BlockTopo &B0 = t.blk[0], &B1 = t.blk[1];
  const int NI0 = B0.n_int, NB0 = B0.n_total - B0.n_int;
  const int NI1 = t.two_blocks ? B1.n_int : 0, NB1 = t.two_blocks ? B1.n_total - B1.n_int : 0;
  t.NI = NI0 + NI1; t.NB = NB0 + NB1; t.NF = B0.n_total + (t.two_blocks ? B1.n_total : 0);
  const LocalTables L0 = local_tables(kind0);
  LocalTables L1{};
  if (t.two_blocks) L1 = local_tables(kind1);

  // ---- per-cell (coefficient dependent) symmetric blocks -> slots -----------------
  std::map<std::pair<int, int>, int> slot00, slot11;
  switch (pairing) {
    case MSFEC_Q:        // (grad phi_i, A grad phi_j), q_basis.cc:228-231
      build_sym_form(B0, t.nC, FORM_TENSOR, L0.vec, nullptr, t.asm00, slot00);
      t.asm00.h_exponent = 1; t.coef_tensor_is_inverse = 0; break;
    case MSFEC_Q_NED:    // (tau, B^-1 sigma) and (curl v, A curl u), q_ned_basis.cc:483-491
      build_sym_form(B0, t.nC, FORM_SCALAR, nullptr, L0.sc, t.asm00, slot00);
      build_sym_form(B1, t.nC, FORM_TENSOR, L1.vec2, nullptr, t.asm11, slot11);
      t.asm00.h_exponent = 3; t.asm11.h_exponent = -1;
      t.coef_tensor_is_inverse = 0; t.coef_scalar_is_inverse = 1; break;
    case MSFEC_NED_RT:   // (tau, A^-1 sigma) and (div v, B div u), ned_rt_basis.cc:488-496
      build_sym_form(B0, t.nC, FORM_TENSOR, L0.vec, nullptr, t.asm00, slot00);
      build_sym_form(B1, t.nC, FORM_SCALAR, nullptr, L1.sc, t.asm11, slot11);
      t.asm00.h_exponent = 1; t.asm11.h_exponent = -3;
      t.coef_tensor_is_inverse = 1; t.coef_scalar_is_inverse = 0; break;
    case MSFEC_RT_DQ:    // (phi, A^-1 psi), rt_dq_basis.cc:495-503 ; block (1,1) = R = 0
      build_sym_form(B0, t.nC, FORM_TENSOR, L0.vec, nullptr, t.asm00, slot00);
      t.asm00.h_exponent = -1; t.coef_tensor_is_inverse = 1; break;
  }
  t.n_slots0 = t.asm00.n_slots; t.n_slots1 = t.asm11.n_slots;

  // ---- coupling block K = block(1,0), cell independent up to h^p ------------------
  std::map<std::pair<int, int>, double> Kmap;   // (row in blk1 full numbering, col in blk0 full numbering)
  if (t.two_blocks) {
    for (int T = 0; T < t.nC; ++T) {
      const int32_t *d0 = &B0.cell_dofs[(size_t)T * B0.ldofs], *d1 = &B1.cell_dofs[(size_t)T * B1.ldofs];
      for (int i = 0; i < B1.ldofs; ++i) for (int j = 0; j < B0.ldofs; ++j) {
        double v = 0.0;
        for (int q = 0; q < 8; ++q) {
          if (pairing == MSFEC_Q_NED)        // (v_i, grad sigma_j)
            for (int c = 0; c < 3; ++c) v += L1.vec[i][q][c] * L0.vec[j][q][c];
          else if (pairing == MSFEC_NED_RT)  // (v_i, curl sigma_j)
            for (int c = 0; c < 3; ++c) v += L1.vec[i][q][c] * L0.vec2[j][q][c];
          else                               // (w_i, div psi_j)
            v += L1.sc[i][q] * L0.sc[j][q];
        }
        v /= 8.0;
        if (v != 0.0) Kmap[{d1[i], d0[j]}] += v;
      }
    }
    t.k_h_exponent = pairing == MSFEC_Q_NED ? 1 : (pairing == MSFEC_NED_RT ? -1 : 0);
  }

  // ---- operators ---------------------------------------------------------------------
  // index helpers: interior row/col index in the stacked interior numbering; boundary col
  auto int_index = [&](int blk, int dof) { return blk == 0 ? dof : NI0 + dof; };
  auto bnd_index = [&](int blk, int dof) { return blk == 0 ? dof - NI0 : NB0 + (dof - NI1); };
  auto full_index = [&](int blk, int dof) { return blk == 0 ? dof : B0.n_total + dof; };
  std::vector<Triplet> sys, lift, full, kint;
  auto add_sym_block = [&](const std::map<std::pair<int, int>, int> &slots, int blk, int slot_off, int nint,
                           bool negate_in_sys) {
    for (auto &kv : slots) {
      const int r = kv.first.first, c = kv.first.second, s = kv.second + slot_off;
      for (int rep = 0; rep < (r == c ? 1 : 2); ++rep) {
        const int rr = rep ? c : r, cc = rep ? r : c;
        full.push_back({full_index(blk, rr), full_index(blk, cc), s << 1, 0.0, false});
        if (rr < nint) {
          if (cc < nint) sys.push_back({int_index(blk, rr), int_index(blk, cc), (s << 1) | (negate_in_sys ? 1 : 0), 0.0, false});
          // lift: b0 = -A00_IB g0 ; b1(sym form) = +A11_IB g1
          else lift.push_back({int_index(blk, rr), bnd_index(blk, cc), (s << 1) | (blk == 0 ? 1 : 0), 0.0, false});
        }
      }
    }
  };
  add_sym_block(slot00, 0, 0, NI0, false);
  if (t.two_blocks) add_sym_block(slot11, 1, t.n_slots0, NI1, true);
  for (auto &kv : Kmap) {
    const int r1 = kv.first.first, c0 = kv.first.second;
    const double v = kv.second;
    full.push_back({full_index(1, r1), full_index(0, c0), 0, v, true});     // +K
    full.push_back({full_index(0, c0), full_index(1, r1), 0, -v, true});    // -K^T
    const bool ri = r1 < NI1, ci = c0 < NI0;
    if (ri && ci) {
      sys.push_back({int_index(1, r1), int_index(0, c0), 0, -v, true});
      sys.push_back({int_index(0, c0), int_index(1, r1), 0, -v, true});
      kint.push_back({r1, c0, 0, v, true});
    }
    if (ri && !ci) lift.push_back({int_index(1, r1), bnd_index(0, c0), 0, v, true});    // b1 = +K g0
    if (!ri && ci) lift.push_back({int_index(0, c0), bnd_index(1, r1), 0, v, true});    // b0 = +K^T g1
  }
  t.sys = make_operator(t.NI, t.NI, sys);
  t.lift = make_operator(t.NI, t.NB, lift);
  t.full = make_operator(t.NF, t.NF, full);
  t.kint = make_operator(NI1, NI0, kint);
  t.diag_slot0.assign(NI0, -1); t.diag_slot1.assign(NI1, -1);
  for (int r = 0; r < NI0; ++r) { auto it = slot00.find({r, r}); if (it != slot00.end()) t.diag_slot0[r] = it->second; }
  for (int r = 0; r < NI1; ++r) { auto it = slot11.find({r, r}); if (it != slot11.end()) t.diag_slot1[r] = it->second + t.n_slots0; }

  // ---- global rhs gather table --------------------------------------------------------
  {
    const int rb = t.two_blocks ? 1 : 0;
    t.rhs_block = rb;
    const BlockTopo &BR = t.blk[rb];
    const LocalTables &LR = rb ? L1 : L0;
    const bool vector_rhs = (pairing == MSFEC_Q_NED || pairing == MSFEC_NED_RT);
    t.rhs_ncomp = vector_rhs ? 3 : 1;
    AsmTable &tab = t.asm_rhs;
    tab.coef_stride = 8 * t.rhs_ncomp;
    tab.pair_ptr.assign(1, 0);
    for (int i = 0; i < BR.ldofs; ++i) {
      for (int q = 0; q < 8; ++q) {
        if (vector_rhs) {
          for (int c = 0; c < 3; ++c) if (LR.vec[i][q][c] != 0.0) {
            tab.pair_idx.push_back(q * 3 + c); tab.pair_w.push_back(LR.vec[i][q][c]);
          }
        } else {
          tab.pair_idx.push_back(q); tab.pair_w.push_back(LR.sc[i][q]);
        }
      }
      tab.pair_ptr.push_back((int32_t)tab.pair_idx.size());
    }
    std::vector<std::vector<std::pair<int, int>>> contribs(BR.n_total);
    for (int T = 0; T < t.nC; ++T)
      for (int i = 0; i < BR.ldofs; ++i) contribs[BR.cell_dofs[(size_t)T * BR.ldofs + i]].push_back({T, i});
    tab.n_slots = BR.n_total;
    tab.contrib_ptr.assign(1, 0);
    for (auto &cl : contribs) {
      for (auto &c : cl) { tab.contrib_cell.push_back(c.first); tab.contrib_pair.push_back(c.second); }
      tab.contrib_ptr.push_back((int32_t)tab.contrib_cell.size());
    }
    // (phi,f): h^3 ; (N,f): h^2 ; (Psi,f): h ; (1,f): h^3     (all x 1/8)
    tab.h_exponent = pairing == MSFEC_Q ? 3 : pairing == MSFEC_Q_NED ? 2 : pairing == MSFEC_NED_RT ? 1 : 3;
  }

  // ---- boundary data G and basis-specific volume rhs F1 on the unit coarse cell -------
  t.G.assign((size_t)t.k_solve * t.NB, 0.0);
  t.F1.assign((size_t)t.k_solve * t.NI, 0.0);
  const double h = 1.0 / n;
  auto coarse_xi = [&](const BlockTopo &B, int dof, double xi[3]) { for (int c = 0; c < 3; ++c) xi[c] = B.pos[3 * dof + c] * h; };
  for (int m = 0; m < t.k_solve; ++m) {
    double *Gm = &t.G[(size_t)m * t.NB];
    // block 0 boundary data
    for (int d = NI0; d < B0.n_total; ++d) {
      double xi[3]; coarse_xi(B0, d, xi);
      double val = 0.0;
      if (kind0 == ENT_V) {          // nodal value of coarse Q1_m (q_basis.cc:143-147, q_ned_basis.cc:263-267)
        if (m < 8) { double v[8], g[8][3]; q1_ref(xi, v, g); val = v[m]; }
      } else if (kind0 == ENT_E) {   // h * (Ned_m . t) at the edge midpoint (ned_rt_basis.cc:261-268)
        if (m < 12) { double v[12][3], cu[12][3]; ned_ref(xi, v, cu); val = h * v[m][B0.axis[d]]; }
      } else {                       // h^2 * (RT_m . n) at the face centre (rt_dq_basis.cc:364-369)
        double v[6][3], dv[6]; rt_ref(xi, v, dv); val = h * h * v[m][B0.axis[d]];
      }
      Gm[d - NI0] = val;
    }
    // block 1 boundary data
    if (t.two_blocks) for (int d = NI1; d < B1.n_total; ++d) {
      double xi[3]; coarse_xi(B1, d, xi);
      double val = 0.0;
      if (pairing == MSFEC_Q_NED) {
        if (m < 8) { double v[8], g[8][3]; q1_ref(xi, v, g); val = h * g[m][B1.axis[d]]; }       // grad Q1_m . t (:268-273)
        else { double v[12][3], cu[12][3]; ned_ref(xi, v, cu); val = h * v[m - 8][B1.axis[d]]; } // Ned . t (:335-340)
      } else if (pairing == MSFEC_NED_RT) {
        if (m >= 12) { double v[6][3], dv[6]; rt_ref(xi, v, dv); val = h * h * v[m - 12][B1.axis[d]]; }   // (:335-342)
      }
      Gm[NB0 + (d - NI1)] = val;
    }
  }
  // F1: rows of block 1 in the symmetric-form system carry -(volume rhs)
  if (t.two_blocks) {
    double qp[8][3]; gauss_points(qp);
    for (int T = 0; T < t.nC; ++T) {
      const int ci = T % n, cj = (T / n) % n, ck = T / (n * n);
      const int32_t *d1 = &B1.cell_dofs[(size_t)T * B1.ldofs];
      for (int q = 0; q < 8; ++q) {
        const double xi[3] = {(ci + qp[q][0]) * h, (cj + qp[q][1]) * h, (ck + qp[q][2]) * h};
        double cq1v[8], cq1g[8][3], cnv[12][3], cnc[12][3];
        if (pairing == MSFEC_Q_NED) q1_ref(xi, cq1v, cq1g);
        if (pairing == MSFEC_NED_RT) ned_ref(xi, cnv, cnc);
        for (int i = 0; i < B1.ldofs; ++i) {
          const int r = d1[i];
          if (r >= NI1) continue;
          if (pairing == MSFEC_Q_NED) {          // (v_i, grad Q1_m): (1/h)(1/H) h^3/8  (q_ned_basis.cc:498-505)
            for (int m = 0; m < 8; ++m) {
              double s = 0; for (int c = 0; c < 3; ++c) s += L1.vec[i][q][c] * cq1g[m][c];
              t.F1[(size_t)m * t.NI + NI0 + r] -= s * h * h / 8.0;
            }
          } else if (pairing == MSFEC_NED_RT) {  // (v_i, curl Ned_m): (1/h^2)(1/H^2) h^3/8  (ned_rt_basis.cc:503-511)
            for (int m = 0; m < 12; ++m) {
              double s = 0; for (int c = 0; c < 3; ++c) s += L1.vec[i][q][c] * cnc[m][c];
              t.F1[(size_t)m * t.NI + NI0 + r] -= s * h / 8.0;
            }
          } else if (pairing == MSFEC_RT_DQ) {   // -+ (w, 1/|K|)  (rt_dq_basis.cc:517-537)
            for (int m = 0; m < 6; ++m) t.F1[(size_t)m * t.NI + NI0 + r] -= ((m & 1) ? 1.0 : -1.0) * h * h * h / 8.0;
          }
        }
      }
    }
    t.f1_H_exponent = pairing == MSFEC_Q_NED ? 1 : (pairing == MSFEC_NED_RT ? -1 : 0);
  }
  return t;
}

std::vector<NormOperator> build_norm_operators(const Topology &t) {
  std::vector<NormOperator> out(4);
  for (int b = 0; b < (t.two_blocks ? 2 : 1); ++b) {
    const BlockTopo &B = t.blk[b];
    const LocalTables L = local_tables(B.kind);
    for (int which = 0; which < 2; ++which) {
      // which = 0: (u, v);  which = 1: (grad u, grad v) / (curl u, curl v) / (div u, div v)
      FormType type;
      const double (*vec)[8][3] = nullptr;
      const double (*sc)[8] = nullptr;
      int hexp;
      if (B.kind == ENT_V) { if (!which) { type = FORM_SCALAR; sc = L.sc; hexp = 3; } else { type = FORM_TENSOR; vec = L.vec; hexp = 1; } }
      else if (B.kind == ENT_E) { type = FORM_TENSOR; if (!which) { vec = L.vec; hexp = 1; } else { vec = L.vec2; hexp = -1; } }
      else if (B.kind == ENT_F) { if (!which) { type = FORM_TENSOR; vec = L.vec; hexp = -1; } else { type = FORM_SCALAR; sc = L.sc; hexp = -3; } }
      else { if (which) continue; type = FORM_SCALAR; sc = L.sc; hexp = 3; }
      AsmTable tab;
      std::map<std::pair<int, int>, int> slots;
      build_sym_form(B, t.nC, type, vec, sc, tab, slots);
      // slot values for the unit coefficient (A = I: channels xx, yy, zz = 1; scalar channel = 1)
      std::vector<double> v(tab.n_slots, 0.0);
      for (int s = 0; s < tab.n_slots; ++s) {
        double acc = 0.0;
        for (int c = tab.contrib_ptr[s]; c < tab.contrib_ptr[s + 1]; ++c) {
          const int pr = tab.contrib_pair[c];
          for (int e = tab.pair_ptr[pr]; e < tab.pair_ptr[pr + 1]; ++e) {
            const int comp = tab.pair_idx[e] % kCoefPerQ;
            if (comp == 0 || comp == 3 || comp == 5 || comp == 6) acc += tab.pair_w[e];
          }
        }
        v[s] = acc / 8.0;
      }
      NormOperator &op = out[2 * b + which];
      op.n = B.n_total; op.h_exponent = hexp;
      op.ptr.assign(op.n + 1, 0);
      for (auto &kv : slots) { op.ptr[kv.first.first + 1]++; if (kv.first.first != kv.first.second) op.ptr[kv.first.second + 1]++; }
      for (int r = 0; r < op.n; ++r) op.ptr[r + 1] += op.ptr[r];
      op.col.resize(op.ptr[op.n]); op.val.resize(op.ptr[op.n]);
      std::vector<int32_t> fill(op.ptr.begin(), op.ptr.end() - 1);
      for (auto &kv : slots) {            // std::map iterates (r, c) in lexicographic order: columns ascend per row
        const int r = kv.first.first, c = kv.first.second;
        op.col[fill[r]] = c; op.val[fill[r]++] = v[kv.second];
        if (r != c) { op.col[fill[c]] = r; op.val[fill[c]++] = v[kv.second]; }
      }
    }
  }
  return out;
}

}  // namespace msfec

namespace msfec {

DirectPlan build_direct_plan(const Topology &t, int ordering) {
  DirectPlan P;
  const int n = t.n, NI0 = t.blk[0].n_int;
  const int NI1 = t.two_blocks ? t.blk[1].n_int : 0;
  constexpr int PW = DirectPlan::kPanel;
  // block key along z in elimination order: layer k -> 2k, plane k -> 2k-1 (split) ; slab -> ceil(z)-1
  auto key_of = [&](int blk, int dof) {
    const double z = t.blk[blk].pos[3 * dof + 2];
    if (ordering == 1) return std::min(std::max((int)std::ceil(z) - 1, 0), n - 1);
    const bool on_plane = (z == std::floor(z));
    const int kz = (int)std::floor(z);
    return on_plane ? 2 * kz - 1 : 2 * kz;
  };
  std::vector<std::vector<int>> slabs;
  if (ordering == 2) {
    // Geometric nested dissection (doubled integer coordinates 0..2n): a box with more than kMinCells fine cells along
    // its longest axis is cut by the mesh plane through its middle; the two halves are eliminated first (recursively),
    // then the DoFs ON the cut (split into pieces, see below).  Every leading set of blocks is again a union of sub-box
    // problems with essential conditions on the cuts, so the no-pivot argument of the layer/plane ordering carries over
    // (RT_DQ needs the cell hand-up described below).  Pays off from n = 16 (per cell: Q_Ned 14 instead of 34 GFLOP,
    // Ned_RT 20 instead of 51, RT_DQ 3.5 instead of 34); at n = 8 the 32-padding of the small blocks costs more than the
    // fill it saves (profiles/ordering_model.py).
    constexpr int kMinCells = 4;
    int sep_piece = 8;                                     // min edge (fine cells) of a separator piece
    if (const char *e = std::getenv("MSFEC_ND_SEP_PIECE")) sep_piece = std::max(1, std::atoi(e));
    struct Dof { int row; int p[3]; };
    std::vector<Dof> all;
    for (int b = 0; b < (t.two_blocks ? 2 : 1); ++b) {
      const int ni = b == 0 ? NI0 : NI1;
      for (int d = 0; d < ni; ++d) {
        Dof q; q.row = (b == 0 ? 0 : NI0) + d;
        for (int c = 0; c < 3; ++c) q.p[c] = (int)std::lround(2.0 * t.blk[b].pos[3 * d + c]);
        all.push_back(q);
      }
    }
    auto emit = [&](std::vector<int> &idx) {
      if (idx.empty()) return;
      std::vector<int> rows;
      for (int i : idx) rows.push_back(all[i].row);
      std::sort(rows.begin(), rows.end());               // stacked numbering: sigma-type rows first
      slabs.push_back(rows);
    };
    // RT_DQ: with all faces of its boundary still uneliminated a box interior is a pure-Neumann problem (u is fixed up to
    // a constant) and the last u pivot of the box would vanish.  Every box therefore hands ONE cell DoF up to the block
    // of the plane that joins it with its sibling: there the first of the two handed-up cells is eliminated (the joined
    // box still lacks the second one) and the second moves on to the next plane; the root's leftover is the pinned DoF.
    const bool defer_u = t.pairing == MSFEC_RT_DQ;
    std::function<int(int[3], int[3], std::vector<int> &)> rec = [&](int lo[3], int hi[3], std::vector<int> &idx) -> int {
      int d = -1, best = kMinCells;
      for (int a : {2, 1, 0}) if ((hi[a] - lo[a]) / 2 > best) { best = (hi[a] - lo[a]) / 2; d = a; }
      if (d < 0 || (int)idx.size() <= PW) {
        int deferred = -1;
        if (defer_u) {
          size_t at = 0;
          for (size_t i = 0; i < idx.size(); ++i)
            if (all[idx[i]].row >= NI0 && (deferred < 0 || all[idx[i]].row > all[deferred].row)) { deferred = idx[i]; at = i; }
          if (deferred >= 0) idx.erase(idx.begin() + at);
        }
        emit(idx);
        return deferred;
      }
      const int mid = lo[d] + ((hi[d] - lo[d]) / 4) * 2;
      std::vector<int> L, R, Sp;
      for (int i : idx) (all[i].p[d] < mid ? L : all[i].p[d] > mid ? R : Sp).push_back(i);
      int l2[3] = {lo[0], lo[1], lo[2]}, h2[3] = {hi[0], hi[1], hi[2]};
      h2[d] = mid;
      const int dl = rec(l2, h2, L);
      h2[d] = hi[d]; l2[d] = mid;
      const int dr = rec(l2, h2, R);
      // The plane is split into pieces of at least kSepPiece x kSepPiece fine cells along the cuts its two half boxes use
      // next (a 16 x 16 plane into 4 quadrants), so that a box couples only with the pieces it touches.  The DoFs ON a
      // dividing line go with the piece before it and the quadrants are taken in cyclic order: what is still fixed of the
      // plane then always hangs on the outer boundary in one connected part, the leading sets stay simply connected
      // (a separate block for the line would leave a slit spanning two faces -- the Ned_RT failure noted above).
      int ax[2], na = 0, cut[2] = {0, 0};
      for (int a = 0; a < 3; ++a) if (a != d) ax[na++] = a;
      bool sp[2];
      for (int i = 0; i < 2; ++i) {
        sp[i] = (hi[ax[i]] - lo[ax[i]]) / 2 >= 2 * sep_piece;
        cut[i] = lo[ax[i]] + ((hi[ax[i]] - lo[ax[i]]) / 4) * 2;
      }
      if (!sp[0] && !sp[1]) {
        if (dl >= 0) Sp.push_back(dl);
        emit(Sp);
        return dr;
      }
      std::vector<int> piece[4];
      for (int i : Sp) {
        const int u = sp[0] && all[i].p[ax[0]] > cut[0], v = sp[1] && all[i].p[ax[1]] > cut[1];
        piece[v ? (u ? 2 : 3) : (u ? 1 : 0)].push_back(i);          // cyclic: (0,0) (1,0) (1,1) (0,1)
      }
      int last = 0;
      for (int q = 0; q < 4; ++q) if (!piece[q].empty()) last = q;
      if (dl >= 0) piece[last].push_back(dl);
      for (int q = 0; q < 4; ++q) emit(piece[q]);
      return dr;
    };
    std::vector<int> idx(all.size());
    for (size_t i = 0; i < all.size(); ++i) idx[i] = (int)i;
    int lo[3] = {0, 0, 0}, hi[3] = {2 * n, 2 * n, 2 * n};
    const int left_over = rec(lo, hi, idx);
    if (left_over >= 0) { slabs.back().push_back(all[left_over].row); std::sort(slabs.back().begin(), slabs.back().end()); }
  } else {
    std::map<int, std::vector<int>> by_key;   // stacked interior row indices, sigma-type first
    for (int d = 0; d < NI0; ++d) by_key[key_of(0, d)].push_back(d);
    for (int d = 0; d < NI1; ++d) by_key[key_of(1, d)].push_back(NI0 + d);
    for (auto &kv : by_key) slabs.push_back(kv.second);
  }
  // Padding-aware balancing: a block whose size exceeds a multiple of 32 by a few DoFs hands its last
  // u-type DoFs to the next block if they fit into that block's padding (n = 8 Ned_RT: layers 161 -> 160).
  // Deferring u-type DoFs keeps every leading matrix non-singular (the leading Schur complement only loses
  // rows/columns), so the no-pivot factorisation is unaffected.
  for (size_t s = 0; ordering != 2 && s + 1 < slabs.size(); ++s) {
    const int excess = (int)slabs[s].size() % PW;
    const int next_fill = (int)slabs[s + 1].size() % PW;
    int n_u = 0;
    for (int r : slabs[s]) n_u += r >= NI0;
    const bool u_block = t.two_blocks;
    if (!u_block || excess == 0 || excess > 8 || n_u < excess) continue;
    if (next_fill == 0 || next_fill + excess > PW) continue;
    for (int q = 0; q < excess; ++q) { slabs[s + 1].push_back(slabs[s].back()); slabs[s].pop_back(); }
  }
  // RT_DQ: the system has the constant-u null space (rt_dq_basis.cc:683-709); fix the last u DoF of the
  // last block to zero.  sigma is unaffected, u is overwritten by 1 afterwards.
  if (t.pairing == MSFEC_RT_DQ) P.pinned_row = slabs.back().back();
  const int nB = (int)slabs.size();
  P.n_slabs = nB;
  P.bs.resize(nB); P.slab_off.resize(nB); P.ld.resize(nB); P.col_off.resize(nB); P.front_rows.resize(nB);
  int off = 0;
  for (int s = 0; s < nB; ++s) {
    P.bs[s] = ((int)slabs[s].size() + PW - 1) / PW * PW;
    P.slab_off[s] = off;
    off += P.bs[s];
  }
  P.NP = off;
  P.perm.assign(t.NI, -1); P.inv_perm.assign(P.NP, -1);
  std::vector<int> blk_of_p(P.NP, 0);
  for (int s = 0; s < nB; ++s) {
    for (int i = 0; i < (int)slabs[s].size(); ++i) {
      P.perm[slabs[s][i]] = P.slab_off[s] + i;
      P.inv_perm[P.slab_off[s] + i] = slabs[s][i];
    }
    for (int i = 0; i < P.bs[s]; ++i) blk_of_p[P.slab_off[s] + i] = s;
  }
  // block graph + symbolic factorisation
  const RefOperator &S = t.sys;
  std::vector<std::vector<char>> adj(nB, std::vector<char>(nB, 0));
  auto couple = [&](int r, int c) {
    const int a = blk_of_p[P.perm[r]], b = blk_of_p[P.perm[c]];
    if (a != b) adj[std::min(a, b)][std::max(a, b)] = 1;
  };
  for (int r = 0; r < S.n_rows; ++r) {
    for (int e = S.cptr[r]; e < S.cptr[r + 1]; ++e) couple(r, S.ccol[e]);
    for (int e = S.sptr[r]; e < S.sptr[r + 1]; ++e) couple(r, S.scol[e]);
  }
  std::vector<std::vector<int>> reach(nB);
  for (int s = 0; s < nB; ++s) {
    for (int b = s + 1; b < nB; ++b) if (adj[s][b]) reach[s].push_back(b);
    for (size_t i = 0; i < reach[s].size(); ++i)
      for (size_t j = i + 1; j < reach[s].size(); ++j) adj[reach[s][i]][reach[s][j]] = 1;
  }
  // fronts, chunk tables, band layout
  P.front_pos.assign((size_t)nB * nB, -1);
  P.chunk_off.assign(1, 0);
  int64_t boff = 0;
  for (int s = 0; s < nB; ++s) {
    int rows = 0;
    std::vector<int> front = {s};
    front.insert(front.end(), reach[s].begin(), reach[s].end());
    for (int b : front) {
      P.front_pos[(size_t)s * nB + b] = rows;
      for (int i = 0; i < P.bs[b]; i += PW) { P.chunk_blk.push_back(b); P.chunk_local.push_back(i); }
      rows += P.bs[b];
    }
    P.chunk_blk.push_back(-1); P.chunk_local.push_back(0);   // the rhs chunk
    P.chunk_off.push_back((int32_t)P.chunk_blk.size());
    P.front_rows[s] = rows;
    P.ld[s] = rows + DirectPlan::kRhsRows;
    P.col_off[s] = boff;
    boff += (int64_t)P.ld[s] * P.bs[s];
    // flops of the right-looking trailing updates of this block column (lower triangle of what lies behind each panel)
    for (int j0 = 0; j0 < P.bs[s]; j0 += PW) {
      const double R = P.ld[s] - (j0 + PW), Cn = rows - (j0 + PW);
      if (Cn > 0) P.update_flops += 2.0 * PW * (Cn * R - Cn * (Cn - 1) / 2.0);
    }
  }
  P.band_doubles = boff;
  if (boff >= (int64_t)1 << 31) throw std::runtime_error("direct solver band exceeds 2^31 entries per cell");
  auto dest_of = [&](int pr, int pc) -> int64_t {   // lower entry (pr >= pc)
    const int c = blk_of_p[pc], r = blk_of_p[pr];
    const int fp = P.front_pos[(size_t)c * nB + r];
    if (fp < 0) throw std::runtime_error("direct plan: entry outside the symbolic block structure");
    return P.col_off[c] + (int64_t)(pc - P.slab_off[c]) * P.ld[c] + fp + (pr - P.slab_off[r]);
  };
  for (int r = 0; r < S.n_rows; ++r) {
    const int pr = P.perm[r];
    for (int e = S.cptr[r]; e < S.cptr[r + 1]; ++e) {
      const int c = S.ccol[e], pc = P.perm[c];
      if (pr < pc || r == P.pinned_row || c == P.pinned_row) continue;
      P.cell_dest.push_back((int32_t)dest_of(pr, pc)); P.cell_ref.push_back(S.cref[e]);
    }
    for (int e = S.sptr[r]; e < S.sptr[r + 1]; ++e) {
      const int c = S.scol[e], pc = P.perm[c];
      if (pr < pc || r == P.pinned_row || c == P.pinned_row) continue;
      P.shared_dest.push_back((int32_t)dest_of(pr, pc)); P.shared_val.push_back(S.sval[e]);
    }
  }
  for (int p = 0; p < P.NP; ++p) {
    const int r = P.inv_perm[p];
    if (r < 0) { P.const_dest.push_back((int32_t)dest_of(p, p)); P.const_val.push_back(1.0); }
    else if (r == P.pinned_row) { P.const_dest.push_back((int32_t)dest_of(p, p)); P.const_val.push_back(-1.0); }
  }
  P.rhs_dest.assign(t.NI, -1);
  for (int r = 0; r < t.NI; ++r) {
    if (r == P.pinned_row) continue;
    const int p = P.perm[r], s = blk_of_p[p];
    P.rhs_dest[r] = (int32_t)(P.col_off[s] + (int64_t)(p - P.slab_off[s]) * P.ld[s] + P.front_rows[s]);
  }
  if (P.pinned_row >= 0) P.inv_perm[P.perm[P.pinned_row]] = -1;   // solution there stays 0
  return P;
}

}  // namespace msfec
