// CPU baseline of the Ned_RT multiscale basis build in compiled C++.  TEST / BENCH INFRASTRUCTURE ONLY: nothing in
// the product path (mpi-msfec_b200/) may include, link or execute this file; bench.py's cpu_baseline / --impl
// reference legs and tests/ do.  PARITY UNPINNED by the reference (deal.II / Trilinos cannot be built here); it is
// checked against oracle/msfec_oracle.py, which carries the analytic invariants (tests/test_oracle.py).
//
// A deal.II-free restatement of NedRTBasis::run() (reference source/Ned_RT/ned_rt_basis.cc:1280-1428) for axis-aligned
// cubes, one coarse cell at a time, cells spread over host threads the way the reference spreads them over MPI ranks:
//   assemble_system                 :362-576   -> assemble()
//   setup_basis_dofs_curl / _div    :225-359   -> closed-form boundary data (last bullet)
//   condense per basis              :1343-1371 -> rhs_j = f_I - A_IB g_B on the interior unknowns
//   solve_iterative(n_basis)        :637-847   -> mode 1, "reference-shaped": for EACH of the 18 right-hand sides
//        ILU(0) of block (0,0) (SparseILU, :672-677), InverseMatrix = GMRES(30) preconditioned by that ILU to
//        1e-6 ||src|| (include/linear_algebra/inverse_matrix.tpp:25-34), Schur complement S = B10 A^-1 B01 - B11
//        (schur_complement.tpp:68-76), CG on S to 1e-6 ||schur_rhs|| (:725-740) preconditioned by 14 CG steps on the
//        ILU-approximated Schur complement (approximate_schur_complement.tpp:38-41, approximate_inverse.tpp:27-34),
//        then sigma = A^-1 (f0 - B01 u) (:790-800)
//   solve_direct                    :579-634   -> mode 0, "exact": banded LDL^T of the symmetric form, layer/plane order
//   assemble_global_element_matrix  :850-948   -> element_matrix()
// Restrictions (it is a baseline, not a second oracle): pairing Ned_RT; coefficient = the harness' rough random
// field or the sine family of eqn_coeff_A.cc / the canonical B of eqn_coeff_B.cc:87-88; right-hand
// side = the polynomial family of the shipped .prm files, scale*((2x-1)(y^2-y)(z^2-z), cyclic).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <thread>
#include <vector>

namespace {

struct Params {
  int L, n;
  uint64_t seed; double sigma;              // random field (seed != 0)
  double a_scale[3], a_alpha[3]; int a_freq[3]; int rotate;
  double b_scale, b_alpha; int b_freq;
  double rhs_scale;
};

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
void field_normals(uint64_t seed, uint64_t gidx, double out[4]) {
  const uint64_t base = (gidx * 4ull) ^ (seed * 0xD1342543DE82EF95ull);
  double f[4];
  for (int c = 0; c < 4; ++c) f[c] = ((double)(splitmix64(base + c) >> 11) + 1.0) * (1.0 / 9007199254740992.0);
  const double r0 = std::sqrt(-2.0 * std::log(f[0])), r1 = std::sqrt(-2.0 * std::log(f[2]));
  const double t0 = 2.0 * M_PI * f[1], t1 = 2.0 * M_PI * f[3];
  out[0] = r0 * std::cos(t0); out[1] = r0 * std::sin(t0); out[2] = r1 * std::cos(t1); out[3] = r1 * std::sin(t1);
}
void rotation(bool rotate, double R[9]) {      // eqn_coeff_A.cc:25-42
  if (!rotate) { const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; std::memcpy(R, I, sizeof(I)); return; }
  const double a = M_PI / 3, b = M_PI / 6, g = M_PI / 4;
  R[0] = std::cos(a) * std::cos(g) - std::sin(a) * std::cos(b) * std::sin(g);
  R[1] = -std::cos(a) * std::sin(g) - std::sin(a) * std::cos(b) * std::cos(g);
  R[2] = std::sin(a) * std::sin(b);
  R[3] = std::sin(a) * std::cos(g) + std::cos(a) * std::cos(b) * std::sin(g);
  R[4] = -std::sin(a) * std::sin(g) + std::cos(a) * std::cos(b) * std::cos(g);
  R[5] = -std::cos(a) * std::sin(b);
  R[6] = std::sin(b) * std::sin(g); R[7] = std::sin(b) * std::cos(g); R[8] = std::cos(b);
}

// ---- reference-cell shape functions ---------------------------------------------------------------
inline double w(int bit, double s) { return bit ? s : 1.0 - s; }
inline double dw(int bit) { return bit ? 1.0 : -1.0; }
struct Line { int d, a0, b0, a1, b1; };
const Line kLines[12] = {{1, 0, 0, 2, 0}, {1, 0, 1, 2, 0}, {0, 1, 0, 2, 0}, {0, 1, 1, 2, 0}, {1, 0, 0, 2, 1}, {1, 0, 1, 2, 1},
                         {0, 1, 0, 2, 1}, {0, 1, 1, 2, 1}, {2, 0, 0, 1, 0}, {2, 0, 1, 1, 0}, {2, 0, 0, 1, 1}, {2, 0, 1, 1, 1}};
void ned_ref(const double xi[3], double val[12][3], double curl[12][3]) {
  for (int l = 0; l < 12; ++l) {
    const Line &q = kLines[l];
    const double f = w(q.b0, xi[q.a0]) * w(q.b1, xi[q.a1]);
    double g[3] = {0, 0, 0};
    g[q.a0] = dw(q.b0) * w(q.b1, xi[q.a1]);
    g[q.a1] = w(q.b0, xi[q.a0]) * dw(q.b1);
    for (int c = 0; c < 3; ++c) val[l][c] = 0;
    val[l][q.d] = f;
    double e[3] = {0, 0, 0}; e[q.d] = 1.0;
    curl[l][0] = g[1] * e[2] - g[2] * e[1]; curl[l][1] = g[2] * e[0] - g[0] * e[2]; curl[l][2] = g[0] * e[1] - g[1] * e[0];
  }
}
void rt_ref(const double xi[3], double val[6][3], double div[6]) {
  for (int f = 0; f < 6; ++f) { const int d = f / 2, s = f & 1; val[f][0] = val[f][1] = val[f][2] = 0; val[f][d] = w(s, xi[d]); div[f] = dw(s); }
}

// ---- fine grid: numbering of oracle/msfec_oracle.py:FineGrid ----------------------------------------------------------
struct Grid {
  int n, n1, nEx, nE, nFx, nF, nC;
  std::vector<int> cE, cF;                 // [nC][12], [nC][6]
  std::vector<double> epos, fpos;          // [nE][3], [nF][3]
  std::vector<int> edir, fdir;
  std::vector<char> ebnd, fbnd;
  explicit Grid(int n_) : n(n_), n1(n_ + 1) {
    nEx = n * n1 * n1; nE = 3 * nEx; nFx = n1 * n * n; nF = 3 * nFx; nC = n * n * n;
    auto EX = [&](int a, int b, int c) { return a + n * (b + n1 * c); };
    auto EY = [&](int a, int b, int c) { return nEx + a + n1 * (b + n * c); };
    auto EZ = [&](int a, int b, int c) { return 2 * nEx + a + n1 * (b + n1 * c); };
    auto FX = [&](int a, int b, int c) { return a + n1 * (b + n * c); };
    auto FY = [&](int a, int b, int c) { return nFx + a + n * (b + n1 * c); };
    auto FZ = [&](int a, int b, int c) { return 2 * nFx + a + n * (b + n * c); };
    cE.resize((size_t)nC * 12); cF.resize((size_t)nC * 6);
    for (int k = 0, T = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i, ++T) {
      const int e[12] = {EY(i, j, k), EY(i + 1, j, k), EX(i, j, k), EX(i, j + 1, k), EY(i, j, k + 1), EY(i + 1, j, k + 1), EX(i, j, k + 1),
                         EX(i, j + 1, k + 1), EZ(i, j, k), EZ(i + 1, j, k), EZ(i, j + 1, k), EZ(i + 1, j + 1, k)};
      const int f[6] = {FX(i, j, k), FX(i + 1, j, k), FY(i, j, k), FY(i, j + 1, k), FZ(i, j, k), FZ(i, j, k + 1)};
      std::copy(e, e + 12, &cE[(size_t)T * 12]); std::copy(f, f + 6, &cF[(size_t)T * 6]);
    }
    epos.assign((size_t)nE * 3, 0); fpos.assign((size_t)nF * 3, 0); edir.assign(nE, 0); fdir.assign(nF, 0);
    auto fill = [&](std::vector<double> &pos, std::vector<int> &dir, int off, int d, int nx, int ny, int nz, double ox, double oy, double oz) {
      int id = off;
      for (int c = 0; c < nz; ++c) for (int b = 0; b < ny; ++b) for (int a = 0; a < nx; ++a, ++id) { pos[3 * id] = a + ox; pos[3 * id + 1] = b + oy; pos[3 * id + 2] = c + oz; dir[id] = d; }
    };
    fill(epos, edir, 0, 0, n, n1, n1, 0.5, 0, 0); fill(epos, edir, nEx, 1, n1, n, n1, 0, 0.5, 0); fill(epos, edir, 2 * nEx, 2, n1, n1, n, 0, 0, 0.5);
    fill(fpos, fdir, 0, 0, n1, n, n, 0, 0.5, 0.5); fill(fpos, fdir, nFx, 1, n, n1, n, 0.5, 0, 0.5); fill(fpos, fdir, 2 * nFx, 2, n, n, n1, 0.5, 0.5, 0);
    ebnd.assign(nE, 0); fbnd.assign(nF, 0);
    for (int e = 0; e < nE; ++e) for (int c = 0; c < 3; ++c) if (c != edir[e] && (epos[3 * e + c] == 0 || epos[3 * e + c] == n)) ebnd[e] = 1;
    for (int f = 0; f < nF; ++f) { const double v = fpos[3 * f + fdir[f]]; fbnd[f] = (v == 0 || v == n); }
  }
};

struct Csr {
  int nr = 0, nc = 0;
  std::vector<int> ptr, col;
  std::vector<double> val;
  void mult(const double *x, double *y) const { for (int r = 0; r < nr; ++r) { double s = 0; for (int e = ptr[r]; e < ptr[r + 1]; ++e) s += val[e] * x[col[e]]; y[r] = s; } }
  void mult_t_add(const double *x, double *y, double sign) const { for (int r = 0; r < nr; ++r) { const double xr = sign * x[r]; for (int e = ptr[r]; e < ptr[r + 1]; ++e) y[col[e]] += val[e] * xr; } }
};
Csr from_maps(const std::vector<std::map<int, double>> &rows, int nc) {
  Csr A; A.nr = (int)rows.size(); A.nc = nc; A.ptr.assign(A.nr + 1, 0);
  for (int r = 0; r < A.nr; ++r) { for (auto &kv : rows[r]) { A.col.push_back(kv.first); A.val.push_back(kv.second); } A.ptr[r + 1] = (int)A.col.size(); }
  return A;
}

struct CellSystem {
  Csr A00, K, A11;                 // interior blocks: edges x edges, faces x edges, faces x faces
  std::vector<double> f0, f1;      // [k][n0], [k][n1] condensed right-hand sides
  // full (unconstrained) blocks and boundary data for the Gram product
  Csr A00f, Kf, A11f;
  std::vector<double> G0, G1, grhs;  // [18][nE], [18][nF], [nF]
  std::vector<int> i0, i1;
};

void assemble(const Params &P, const Grid &g, const double *corners, long long gid, CellSystem &cs) {
  const int n = g.n;
  const double x0[3] = {corners[0], corners[1], corners[2]}, H = corners[21] - corners[0], h = H / n, JxW = h * h * h / 8.0;
  const double G2[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
  double nedv[8][12][3], nedc[8][12][3], rtv[8][6][3], rtd[8][6];
  for (int q = 0; q < 8; ++q) {
    const double xi[3] = {G2[q & 1], G2[(q >> 1) & 1], G2[q >> 2]};
    ned_ref(xi, nedv[q], nedc[q]); rt_ref(xi, rtv[q], rtd[q]);
    for (int i = 0; i < 12; ++i) for (int c = 0; c < 3; ++c) { nedv[q][i][c] /= h; nedc[q][i][c] /= h * h; }
    for (int i = 0; i < 6; ++i) { for (int c = 0; c < 3; ++c) rtv[q][i][c] /= h * h; rtd[q][i] /= h * h * h; }
  }
  double R[9];
  rotation(P.rotate != 0, R);
  std::vector<std::map<int, double>> a00(g.nE), kk(g.nF), a11(g.nF);
  cs.grhs.assign(g.nF, 0.0);
  std::vector<double> F1((size_t)18 * g.nF, 0.0);
  for (int T = 0; T < g.nC; ++T) {
    const int ci = T % n, cj = (T / n) % n, ck = T / (n * n);
    double l00[12][12] = {}, l10[6][12] = {}, l11[6][6] = {}, lr[6] = {}, lf[12][6] = {};
    double xi4[4];
    if (P.seed) field_normals(P.seed, (uint64_t)gid * (uint64_t)g.nC + (uint64_t)T, xi4);
    for (int q = 0; q < 8; ++q) {
      const double px = x0[0] + h * (ci + G2[q & 1]), py = x0[1] + h * (cj + G2[(q >> 1) & 1]), pz = x0[2] + h * (ck + G2[q >> 2]);
      double d[3], B;
      if (P.seed) { for (int c = 0; c < 3; ++c) d[c] = std::exp(P.sigma * xi4[c]); B = std::exp(P.sigma * xi4[3]); }
      else {
        const double p[3] = {px, py, pz};
        for (int c = 0; c < 3; ++c) d[c] = P.a_scale[c] * (1.0 - P.a_alpha[c] * std::sin(2.0 * M_PI * P.a_freq[c] * p[c]));
        B = P.b_scale * (1.0 - P.b_alpha * std::sin(2.0 * M_PI * P.b_freq * px));
      }
      double Ai[3][3];
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Ai[a][b] = R[a * 3] * R[b * 3] / d[0] + R[a * 3 + 1] * R[b * 3 + 1] / d[1] + R[a * 3 + 2] * R[b * 3 + 2] / d[2];
      const double f[3] = {P.rhs_scale * (2 * px - 1) * (py * py - py) * (pz * pz - pz), P.rhs_scale * (2 * py - 1) * (px * px - px) * (pz * pz - pz),
                           P.rhs_scale * (2 * pz - 1) * (px * px - px) * (py * py - py)};
      // coarse Nedelec curls at the physical point (basis-specific volume rhs, ned_rt_basis.cc:503-511)
      const double xic[3] = {(px - x0[0]) / H, (py - x0[1]) / H, (pz - x0[2]) / H};
      double cv[12][3], cc[12][3];
      ned_ref(xic, cv, cc);
      for (int i = 0; i < 12; ++i) {
        double t[3];
        for (int a = 0; a < 3; ++a) t[a] = Ai[a][0] * nedv[q][i][0] + Ai[a][1] * nedv[q][i][1] + Ai[a][2] * nedv[q][i][2];
        for (int j = 0; j < 12; ++j) l00[i][j] += (t[0] * nedv[q][j][0] + t[1] * nedv[q][j][1] + t[2] * nedv[q][j][2]) * JxW;
      }
      for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 12; ++j) l10[i][j] += (rtv[q][i][0] * nedc[q][j][0] + rtv[q][i][1] * nedc[q][j][1] + rtv[q][i][2] * nedc[q][j][2]) * JxW;
        for (int j = 0; j < 6; ++j) l11[i][j] += rtd[q][i] * B * rtd[q][j] * JxW;
        lr[i] += (rtv[q][i][0] * f[0] + rtv[q][i][1] * f[1] + rtv[q][i][2] * f[2]) * JxW;
        for (int m = 0; m < 12; ++m) lf[m][i] += (rtv[q][i][0] * cc[m][0] + rtv[q][i][1] * cc[m][1] + rtv[q][i][2] * cc[m][2]) / (H * H) * JxW;
      }
    }
    const int *e = &g.cE[(size_t)T * 12], *f = &g.cF[(size_t)T * 6];
    for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) a00[e[i]][e[j]] += l00[i][j];
    for (int i = 0; i < 6; ++i) {
      for (int j = 0; j < 12; ++j) kk[f[i]][e[j]] += l10[i][j];
      for (int j = 0; j < 6; ++j) a11[f[i]][f[j]] += l11[i][j];
      cs.grhs[f[i]] += lr[i];
      for (int m = 0; m < 12; ++m) F1[(size_t)m * g.nF + f[i]] += lf[m][i];
    }
  }
  cs.A00f = from_maps(a00, g.nE); cs.Kf = from_maps(kk, g.nE); cs.A11f = from_maps(a11, g.nF);
  // boundary data: g_e = h (Ned_i^coarse(m) . t), g_F = h^2 (RT_j^coarse(c) . n)
  cs.G0.assign((size_t)18 * g.nE, 0.0); cs.G1.assign((size_t)18 * g.nF, 0.0);
  for (int e = 0; e < g.nE; ++e) if (g.ebnd[e]) {
    const double xi[3] = {g.epos[3 * e] / n, g.epos[3 * e + 1] / n, g.epos[3 * e + 2] / n};
    double v[12][3], c[12][3];
    ned_ref(xi, v, c);
    for (int m = 0; m < 12; ++m) cs.G0[(size_t)m * g.nE + e] = h * v[m][g.edir[e]] / H;
  }
  for (int f = 0; f < g.nF; ++f) if (g.fbnd[f]) {
    const double xi[3] = {g.fpos[3 * f] / n, g.fpos[3 * f + 1] / n, g.fpos[3 * f + 2] / n};
    double v[6][3], dv[6];
    rt_ref(xi, v, dv);
    for (int m = 0; m < 6; ++m) cs.G1[(size_t)(12 + m) * g.nF + f] = h * h * v[m][g.fdir[f]] / (H * H);
  }
  // interior index sets and blocks
  std::vector<int> m0(g.nE, -1), m1(g.nF, -1);
  cs.i0.clear(); cs.i1.clear();
  for (int e = 0; e < g.nE; ++e) if (!g.ebnd[e]) { m0[e] = (int)cs.i0.size(); cs.i0.push_back(e); }
  for (int f = 0; f < g.nF; ++f) if (!g.fbnd[f]) { m1[f] = (int)cs.i1.size(); cs.i1.push_back(f); }
  auto restrict = [&](const Csr &A, const std::vector<int> &rows, const std::vector<int> &cmap, int nc) {
    Csr B; B.nr = (int)rows.size(); B.nc = nc; B.ptr.assign(B.nr + 1, 0);
    for (int r = 0; r < B.nr; ++r) {
      for (int e = A.ptr[rows[r]]; e < A.ptr[rows[r] + 1]; ++e) if (cmap[A.col[e]] >= 0) { B.col.push_back(cmap[A.col[e]]); B.val.push_back(A.val[e]); }
      B.ptr[r + 1] = (int)B.col.size();
    }
    return B;
  };
  cs.A00 = restrict(cs.A00f, cs.i0, m0, (int)cs.i0.size());
  cs.K = restrict(cs.Kf, cs.i1, m0, (int)cs.i0.size());
  cs.A11 = restrict(cs.A11f, cs.i1, m1, (int)cs.i1.size());
  // condensed right-hand sides: row 0: A00 s - K^T u = 0, row 1: K s + A11 u = F1
  const int n0 = (int)cs.i0.size(), n1 = (int)cs.i1.size();
  cs.f0.assign((size_t)18 * n0, 0.0); cs.f1.assign((size_t)18 * n1, 0.0);
  std::vector<double> t0(g.nE), t1(g.nF), t2(g.nF);
  for (int m = 0; m < 18; ++m) {
    const double *g0 = &cs.G0[(size_t)m * g.nE], *g1 = &cs.G1[(size_t)m * g.nF];
    cs.A00f.mult(g0, t0.data());
    for (auto &v : t0) v = -v;
    cs.Kf.mult_t_add(g1, t0.data(), 1.0);
    cs.Kf.mult(g0, t1.data()); cs.A11f.mult(g1, t2.data());
    for (int i = 0; i < n0; ++i) cs.f0[(size_t)m * n0 + i] = t0[cs.i0[i]];
    for (int i = 0; i < n1; ++i) cs.f1[(size_t)m * n1 + i] = (m < 12 ? F1[(size_t)m * g.nF + cs.i1[i]] : 0.0) - t1[cs.i1[i]] - t2[cs.i1[i]];
  }
}

// ---- reference-shaped solver pieces ----------------------------------------------------------------------------------
struct Ilu0 {                      // SparseILU with default AdditionalData: ILU(0) on the pattern of A
  Csr LU; std::vector<int> diag;
  void init(const Csr &A) {
    LU = A; diag.assign(A.nr, -1);
    for (int r = 0; r < A.nr; ++r) for (int e = A.ptr[r]; e < A.ptr[r + 1]; ++e) if (A.col[e] == r) diag[r] = e;
    std::vector<int> pos(A.nc, -1);
    for (int i = 0; i < A.nr; ++i) {
      for (int e = LU.ptr[i]; e < LU.ptr[i + 1]; ++e) pos[LU.col[e]] = e;
      for (int e = LU.ptr[i]; e < LU.ptr[i + 1] && LU.col[e] < i; ++e) {
        const int k = LU.col[e];
        const double l = LU.val[e] / LU.val[diag[k]];
        LU.val[e] = l;
        for (int f = diag[k] + 1; f < LU.ptr[k + 1]; ++f) { const int p = pos[LU.col[f]]; if (p >= 0) LU.val[p] -= l * LU.val[f]; }
      }
      for (int e = LU.ptr[i]; e < LU.ptr[i + 1]; ++e) pos[LU.col[e]] = -1;
    }
  }
  void apply(const double *b, double *x) const {
    const int n = LU.nr;
    for (int i = 0; i < n; ++i) { double s = b[i]; for (int e = LU.ptr[i]; e < diag[i]; ++e) s -= LU.val[e] * x[LU.col[e]]; x[i] = s; }
    for (int i = n - 1; i >= 0; --i) { double s = x[i]; for (int e = diag[i] + 1; e < LU.ptr[i + 1]; ++e) s -= LU.val[e] * x[LU.col[e]]; x[i] = s / LU.val[diag[i]]; }
  }
};

double nrm2(const std::vector<double> &v) { double s = 0; for (double x : v) s += x * x; return std::sqrt(s); }

// left-preconditioned restarted GMRES(30): dst = A^-1 src to tol * ||src|| (InverseMatrix::vmult)
int gmres(const Csr &A, const Ilu0 &M, const std::vector<double> &src, std::vector<double> &dst, double tol_abs, int max_it) {
  const int n = A.nr, m = 30;
  dst.assign(n, 0.0);
  std::vector<std::vector<double>> V(m + 1, std::vector<double>(n));
  std::vector<double> Hh((size_t)(m + 1) * m), cs(m), sn(m), gvec(m + 1), tmp(n), y(m);
  int its = 0;
  while (its < max_it) {
    A.mult(dst.data(), tmp.data());
    for (int i = 0; i < n; ++i) tmp[i] = src[i] - tmp[i];
    M.apply(tmp.data(), V[0].data());
    double beta = nrm2(V[0]);
    if (beta <= tol_abs) return its;
    for (double &v : V[0]) v /= beta;
    std::fill(gvec.begin(), gvec.end(), 0.0); gvec[0] = beta;
    int j = 0;
    for (; j < m && its < max_it; ++j, ++its) {
      A.mult(V[j].data(), tmp.data());
      M.apply(tmp.data(), V[j + 1].data());
      for (int i = 0; i <= j; ++i) {
        double hij = 0; for (int q = 0; q < n; ++q) hij += V[i][q] * V[j + 1][q];
        Hh[(size_t)i * m + j] = hij;
        for (int q = 0; q < n; ++q) V[j + 1][q] -= hij * V[i][q];
      }
      const double hn = nrm2(V[j + 1]);
      Hh[(size_t)(j + 1) * m + j] = hn;
      if (hn > 0) for (double &v : V[j + 1]) v /= hn;
      for (int i = 0; i < j; ++i) {
        const double a = Hh[(size_t)i * m + j], b = Hh[(size_t)(i + 1) * m + j];
        Hh[(size_t)i * m + j] = cs[i] * a + sn[i] * b; Hh[(size_t)(i + 1) * m + j] = -sn[i] * a + cs[i] * b;
      }
      const double a = Hh[(size_t)j * m + j], b = Hh[(size_t)(j + 1) * m + j], r = std::hypot(a, b);
      cs[j] = a / r; sn[j] = b / r;
      Hh[(size_t)j * m + j] = r; Hh[(size_t)(j + 1) * m + j] = 0;
      gvec[j + 1] = -sn[j] * gvec[j]; gvec[j] = cs[j] * gvec[j];
      if (std::fabs(gvec[j + 1]) <= tol_abs) { ++j; ++its; break; }
    }
    for (int i = j - 1; i >= 0; --i) { double s = gvec[i]; for (int q = i + 1; q < j; ++q) s -= Hh[(size_t)i * m + q] * y[q]; y[i] = s / Hh[(size_t)i * m + i]; }
    for (int i = 0; i < j; ++i) for (int q = 0; q < n; ++q) dst[q] += y[i] * V[i][q];
    if (std::fabs(gvec[j]) <= tol_abs) return its;
  }
  return its;
}

template <class Op, class Prec>
int cg(const Op &A, const Prec &M, const std::vector<double> &b, std::vector<double> &x, double tol_abs, int max_it) {
  const size_t n = b.size();
  x.assign(n, 0.0);
  std::vector<double> r = b, z(n), p(n), Ap(n);
  if (nrm2(r) <= tol_abs) return 0;
  M(r, z); p = z;
  double rz = 0; for (size_t i = 0; i < n; ++i) rz += r[i] * z[i];
  for (int it = 1; it <= max_it; ++it) {
    A(p, Ap);
    double pAp = 0; for (size_t i = 0; i < n; ++i) pAp += p[i] * Ap[i];
    const double alpha = rz / pAp;
    for (size_t i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
    if (nrm2(r) <= tol_abs) return it;
    M(r, z);
    double rz2 = 0; for (size_t i = 0; i < n; ++i) rz2 += r[i] * z[i];
    const double beta = rz2 / rz; rz = rz2;
    for (size_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
  }
  return max_it;
}

// solve_iterative for ONE right-hand side; iteration counts are added to its[0] (outer CG) and its[1] (inner GMRES)
void solve_reference_shaped(const CellSystem &cs, const double *f0, const double *f1, double *sig, double *u, long its[2]) {
  const int n0 = cs.A00.nr, n1 = cs.A11.nr;
  Ilu0 ilu_inner, ilu_approx;                 // the reference builds the ILU twice per rhs
  ilu_inner.init(cs.A00); ilu_approx.init(cs.A00);
  auto inv00 = [&](const std::vector<double> &src, std::vector<double> &dst) { its[1] += gmres(cs.A00, ilu_inner, src, dst, 1e-6 * nrm2(src), std::max(n0, 1000)); };
  std::vector<double> t0(n0), t1(n0), t2(n1);
  // positive form of the reference's Schur complement: S = A11 + K A00^-1 K^T  (theirs: B10 A^-1 B01 - B11 = -S)
  auto S = [&](const std::vector<double> &v, std::vector<double> &y) {
    std::fill(t0.begin(), t0.end(), 0.0); cs.K.mult_t_add(v.data(), t0.data(), 1.0);
    inv00(t0, t1);
    y.resize(n1); cs.K.mult(t1.data(), y.data());
    cs.A11.mult(v.data(), t2.data());
    for (int i = 0; i < n1; ++i) y[i] += t2[i];
  };
  std::vector<double> a0(n0), a1(n0), a2(n1);
  auto S_approx = [&](const std::vector<double> &v, std::vector<double> &y) {
    std::fill(a0.begin(), a0.end(), 0.0); cs.K.mult_t_add(v.data(), a0.data(), 1.0);
    ilu_approx.apply(a0.data(), a1.data());
    y.resize(n1); cs.K.mult(a1.data(), y.data());
    cs.A11.mult(v.data(), a2.data());
    for (int i = 0; i < n1; ++i) y[i] += a2[i];
  };
  auto identity = [](const std::vector<double> &r, std::vector<double> &z) { z = r; };
  auto precond = [&](const std::vector<double> &r, std::vector<double> &z) { cg(S_approx, identity, r, z, 1e-6 * nrm2(r), 14); };
  // schur_rhs = f1 - K A00^-1 f0
  std::vector<double> F0(f0, f0 + n0), tmp, rhs(n1), uu, ss;
  inv00(F0, tmp);
  cs.K.mult(tmp.data(), rhs.data());
  for (int i = 0; i < n1; ++i) rhs[i] = f1[i] - rhs[i];
  its[0] += cg(S, precond, rhs, uu, 1e-6 * nrm2(rhs), n0 + n1);
  // sigma = A00^-1 (f0 + K^T u)
  std::vector<double> g0(F0);
  cs.K.mult_t_add(uu.data(), g0.data(), 1.0);
  inv00(g0, ss);
  std::copy(ss.begin(), ss.end(), sig); std::copy(uu.begin(), uu.end(), u);
}

// exact solve of all 18 right-hand sides: banded LDL^T (no pivoting) of [A00 -K^T; -K -A11] in layer / plane order with the
// sigma-type unknowns first inside a block (every leading principal minor is a sub-box problem)
void solve_exact(const Grid &g, const CellSystem &cs, std::vector<double> &X0, std::vector<double> &X1) {
  const int n0 = cs.A00.nr, n1 = cs.A11.nr, N = n0 + n1;
  std::vector<std::pair<long, int>> key(N);
  for (int i = 0; i < n0; ++i) { const double z = g.epos[3 * cs.i0[i] + 2]; const bool pl = z == std::floor(z); key[i] = {(long)(pl ? 2 * (int)z - 1 : 2 * (int)z) * 4 + 0, i}; }
  for (int i = 0; i < n1; ++i) { const double z = g.fpos[3 * cs.i1[i] + 2]; const bool pl = z == std::floor(z); key[n0 + i] = {(long)(pl ? 2 * (int)z - 1 : 2 * (int)z) * 4 + 1, n0 + i}; }
  std::sort(key.begin(), key.end());
  std::vector<int> perm(N), inv(N);
  for (int p = 0; p < N; ++p) { perm[p] = key[p].second; inv[key[p].second] = p; }
  // half bandwidth
  int bw = 0;
  auto upd = [&](int r, int c) { bw = std::max(bw, std::abs(inv[r] - inv[c])); };
  for (int r = 0; r < n0; ++r) for (int e = cs.A00.ptr[r]; e < cs.A00.ptr[r + 1]; ++e) upd(r, cs.A00.col[e]);
  for (int r = 0; r < n1; ++r) { for (int e = cs.K.ptr[r]; e < cs.K.ptr[r + 1]; ++e) upd(n0 + r, cs.K.col[e]); for (int e = cs.A11.ptr[r]; e < cs.A11.ptr[r + 1]; ++e) upd(n0 + r, n0 + cs.A11.col[e]); }
  const int ld = bw + 1;
  std::vector<double> B((size_t)N * ld, 0.0);            // B[p][p - q] = A(p, q), q <= p
  auto put = [&](int r, int c, double v) { const int p = inv[r], q = inv[c]; if (q <= p) B[(size_t)p * ld + (p - q)] += v; };
  for (int r = 0; r < n0; ++r) for (int e = cs.A00.ptr[r]; e < cs.A00.ptr[r + 1]; ++e) put(r, cs.A00.col[e], cs.A00.val[e]);
  for (int r = 0; r < n1; ++r) {
    for (int e = cs.K.ptr[r]; e < cs.K.ptr[r + 1]; ++e) { put(n0 + r, cs.K.col[e], -cs.K.val[e]); put(cs.K.col[e], n0 + r, -cs.K.val[e]); }
    for (int e = cs.A11.ptr[r]; e < cs.A11.ptr[r + 1]; ++e) put(n0 + r, n0 + cs.A11.col[e], -cs.A11.val[e]);
  }
  // LDL^T in place: row p holds L(p, q) for q in [p - bw, p), D(p) at offset 0
  std::vector<double> tmp(ld);
  for (int p = 0; p < N; ++p) {
    const int q0 = std::max(0, p - bw);
    for (int q = q0; q < p; ++q) {
      double s = B[(size_t)p * ld + (p - q)];
      const int t0 = std::max(q0, std::max(0, q - bw));
      for (int t = t0; t < q; ++t) s -= tmp[p - t] * B[(size_t)q * ld + (q - t)];
      tmp[p - q] = s;                                      // = L(p, q) D(q)
      B[(size_t)p * ld + (p - q)] = s / B[(size_t)q * ld];
    }
    double d = B[(size_t)p * ld];
    for (int q = q0; q < p; ++q) d -= tmp[p - q] * B[(size_t)p * ld + (p - q)];
    B[(size_t)p * ld] = d;
  }
  X0.assign((size_t)18 * n0, 0.0); X1.assign((size_t)18 * n1, 0.0);
  std::vector<double> y(N);
  for (int m = 0; m < 18; ++m) {
    for (int p = 0; p < N; ++p) { const int r = perm[p]; y[p] = r < n0 ? cs.f0[(size_t)m * n0 + r] : -cs.f1[(size_t)m * n1 + (r - n0)]; }
    for (int p = 0; p < N; ++p) { double s = y[p]; for (int q = std::max(0, p - bw); q < p; ++q) s -= B[(size_t)p * ld + (p - q)] * y[q]; y[p] = s; }
    for (int p = 0; p < N; ++p) y[p] /= B[(size_t)p * ld];
    for (int p = N - 1; p >= 0; --p) { const double yp = y[p]; for (int q = std::max(0, p - bw); q < p; ++q) y[q] -= B[(size_t)p * ld + (p - q)] * yp; }
    for (int p = 0; p < N; ++p) { const int r = perm[p]; if (r < n0) X0[(size_t)m * n0 + r] = y[p]; else X1[(size_t)m * n1 + (r - n0)] = y[p]; }
  }
}

void element_matrix(const Grid &g, const CellSystem &cs, const std::vector<double> &X0, const std::vector<double> &X1, double *M, double *r) {
  const int n0 = (int)cs.i0.size(), n1 = (int)cs.i1.size();
  // full basis vectors: sigma part of the 12 curl bases, u part of the 6 div bases (ned_rt_basis.cc:850-948)
  std::vector<double> Sg((size_t)12 * g.nE), U((size_t)6 * g.nF);
  for (int m = 0; m < 12; ++m) { std::copy(&cs.G0[(size_t)m * g.nE], &cs.G0[(size_t)(m + 1) * g.nE], &Sg[(size_t)m * g.nE]); for (int i = 0; i < n0; ++i) Sg[(size_t)m * g.nE + cs.i0[i]] = X0[(size_t)m * n0 + i]; }
  for (int m = 0; m < 6; ++m) { std::copy(&cs.G1[(size_t)(12 + m) * g.nF], &cs.G1[(size_t)(13 + m) * g.nF], &U[(size_t)m * g.nF]); for (int i = 0; i < n1; ++i) U[(size_t)m * g.nF + cs.i1[i]] = X1[(size_t)(12 + m) * n1 + i]; }
  std::vector<double> AS((size_t)12 * g.nE), KS((size_t)12 * g.nF), AU((size_t)6 * g.nF);
  for (int m = 0; m < 12; ++m) { cs.A00f.mult(&Sg[(size_t)m * g.nE], &AS[(size_t)m * g.nE]); cs.Kf.mult(&Sg[(size_t)m * g.nE], &KS[(size_t)m * g.nF]); }
  for (int m = 0; m < 6; ++m) cs.A11f.mult(&U[(size_t)m * g.nF], &AU[(size_t)m * g.nF]);
  auto dot = [](const double *a, const double *b, int n) { double s = 0; for (int i = 0; i < n; ++i) s += a[i] * b[i]; return s; };
  std::fill(M, M + 18 * 18, 0.0); std::fill(r, r + 18, 0.0);
  for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) M[i * 18 + j] = dot(&Sg[(size_t)i * g.nE], &AS[(size_t)j * g.nE], g.nE);
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 12; ++j) { const double v = dot(&U[(size_t)i * g.nF], &KS[(size_t)j * g.nF], g.nF); M[(12 + i) * 18 + j] = v; M[j * 18 + 12 + i] = -v; }
  for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) M[(12 + i) * 18 + 12 + j] = dot(&U[(size_t)i * g.nF], &AU[(size_t)j * g.nF], g.nF); r[12 + i] = dot(&U[(size_t)i * g.nF], cs.grhs.data(), g.nF); }
}

}  // namespace

extern "C" {

// mode 0: exact banded LDL^T; mode 1: reference-shaped Schur-complement CG + GMRES(ILU(0)) at 1e-6, one rhs at a time.
// prm: {a_scale[3], a_alpha[3], a_freq[3], rotate, b_scale, b_alpha, b_freq, rhs_scale} (15 doubles).
// stats (optional): [n_cells][2] outer CG / inner GMRES iterations summed over the 18 right-hand sides.  Returns 0.
int msfec_cpu_ned_rt(int L, int n_cells, const double *corners, const int64_t *ids, uint64_t seed, double sigma, const double *prm,
                     int mode, int n_threads, double *M, double *r, double *stats) {
  if (L < 1 || L > 5 || n_cells <= 0 || !corners || !M || !r || !prm) return 1;
  Params P{};
  P.L = L; P.n = 1 << L; P.seed = seed; P.sigma = sigma;
  for (int d = 0; d < 3; ++d) { P.a_scale[d] = prm[d]; P.a_alpha[d] = prm[3 + d]; P.a_freq[d] = (int)prm[6 + d]; }
  P.rotate = (int)prm[9]; P.b_scale = prm[10]; P.b_alpha = prm[11]; P.b_freq = (int)prm[12]; P.rhs_scale = prm[13];
  const Grid g(P.n);
  std::atomic<int> next{0};
  auto work = [&]() {
    for (;;) {
      const int c = next.fetch_add(1);
      if (c >= n_cells) break;
      CellSystem cs;
      assemble(P, g, corners + (size_t)c * 24, ids ? (long long)ids[c] : c, cs);
      const int n0 = (int)cs.i0.size(), n1 = (int)cs.i1.size();
      std::vector<double> X0, X1;
      long its[2] = {0, 0};
      if (mode == 0) solve_exact(g, cs, X0, X1);
      else {
        X0.assign((size_t)18 * n0, 0.0); X1.assign((size_t)18 * n1, 0.0);
        for (int m = 0; m < 18; ++m) solve_reference_shaped(cs, &cs.f0[(size_t)m * n0], &cs.f1[(size_t)m * n1], &X0[(size_t)m * n0], &X1[(size_t)m * n1], its);
      }
      element_matrix(g, cs, X0, X1, M + (size_t)c * 324, r + (size_t)c * 18);
      if (stats) { stats[2 * c] = (double)its[0]; stats[2 * c + 1] = (double)its[1]; }
    }
  };
  const int nt = std::max(1, n_threads);
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto &t : pool) t.join();
  return 0;
}

}  // extern "C"
