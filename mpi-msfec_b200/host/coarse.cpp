#include "coarse.h"

#include <algorithm>
#include <cmath>
#include <functional>
#include <numeric>
#include <sstream>
#include <stdexcept>

#include "../../include/msfec.h"
#include "../../include/msfec_coarse.h"

namespace msfec {

namespace {

enum { KV = 0, KE = 1, KF = 2, KC = 3 };

struct Grid {
  int m;
  long long nV() const { return (long long)(m + 1) * (m + 1) * (m + 1); }
  long long nE() const { return 3LL * m * (m + 1) * (m + 1); }
  long long nF() const { return 3LL * (m + 1) * m * m; }
  long long nC() const { return (long long)m * m * m; }
  long long count(int kind) const { return kind == KV ? nV() : kind == KE ? nE() : kind == KF ? nF() : kind == KC ? nC() : 0; }
  int V(int i, int j, int k) const { return i + (m + 1) * (j + (m + 1) * k); }
  int E(int d, int i, int j, int k) const {                 // edge parallel to axis d, lower end at vertex (i, j, k)
    const int ex = m * (m + 1) * (m + 1);
    if (d == 0) return i + m * (j + (m + 1) * k);
    if (d == 1) return ex + i + (m + 1) * (j + m * k);
    return 2 * ex + i + (m + 1) * (j + (m + 1) * k);
  }
  int F(int d, int i, int j, int k) const {                 // face normal to axis d, lower corner at vertex (i, j, k)
    const int fx = (m + 1) * m * m;
    if (d == 0) return i + (m + 1) * (j + m * k);
    if (d == 1) return fx + i + m * (j + (m + 1) * k);
    return 2 * fx + i + m * (j + m * k);
  }
  int C(int i, int j, int k) const { return i + m * (j + m * k); }
  static int ldofs(int kind) { return kind == KV ? 8 : kind == KE ? 12 : kind == KF ? 6 : kind == KC ? 1 : 0; }
  // local -> global entity indices of cell (ci, cj, ck), deal.II local order
  void local(int kind, int ci, int cj, int ck, int32_t *out) const {
    if (kind == KV) {
      for (int v = 0; v < 8; ++v) out[v] = V(ci + (v & 1), cj + ((v >> 1) & 1), ck + (v >> 2));
    } else if (kind == KE) {
      for (int t = 0; t < 2; ++t) {                          // bottom (z = ck) and top (z = ck + 1) faces
        out[4 * t + 0] = E(1, ci, cj, ck + t);               // x = 0, parallel to y
        out[4 * t + 1] = E(1, ci + 1, cj, ck + t);           // x = 1, parallel to y
        out[4 * t + 2] = E(0, ci, cj, ck + t);               // y = 0, parallel to x
        out[4 * t + 3] = E(0, ci, cj + 1, ck + t);           // y = 1, parallel to x
      }
      out[8] = E(2, ci, cj, ck); out[9] = E(2, ci + 1, cj, ck); out[10] = E(2, ci, cj + 1, ck); out[11] = E(2, ci + 1, cj + 1, ck);
    } else if (kind == KF) {
      out[0] = F(0, ci, cj, ck); out[1] = F(0, ci + 1, cj, ck);
      out[2] = F(1, ci, cj, ck); out[3] = F(1, ci, cj + 1, ck);
      out[4] = F(2, ci, cj, ck); out[5] = F(2, ci, cj, ck + 1);
    } else if (kind == KC) {
      out[0] = C(ci, cj, ck);
    }
  }
};

double dot(const std::vector<double> &a, const std::vector<double> &b) {
  double s = 0;
#pragma omp parallel for reduction(+ : s) if (a.size() > 20000)
  for (long i = 0; i < (long)a.size(); ++i) s += a[i] * b[i];
  return s;
}

struct Csr {
  int n_rows = 0, n_cols = 0;
  std::vector<int32_t> ptr, col;
  std::vector<double> val;
  void mult(const std::vector<double> &x, std::vector<double> &y) const {
    y.assign(n_rows, 0.0);
#pragma omp parallel for if (n_rows > 5000)
    for (int r = 0; r < n_rows; ++r) {
      double s = 0;
      for (int e = ptr[r]; e < ptr[r + 1]; ++e) s += val[e] * x[col[e]];
      y[r] = s;
    }
  }
};

// Jacobi-preconditioned CG; returns iterations (negative: not converged)
int cg(const std::function<void(const std::vector<double> &, std::vector<double> &)> &A, const std::vector<double> &dinv,
       const std::vector<double> &b, std::vector<double> &x, double rtol, int max_it) {
  const size_t n = b.size();
  x.assign(n, 0.0);
  std::vector<double> r = b, z(n), p(n), Ap(n);
  const double bn = std::sqrt(dot(b, b));
  if (bn == 0) return 0;
  for (size_t i = 0; i < n; ++i) z[i] = dinv.empty() ? r[i] : dinv[i] * r[i];
  p = z;
  double rz = dot(r, z);
  for (int it = 1; it <= max_it; ++it) {
    A(p, Ap);
    const double alpha = rz / dot(p, Ap);
    for (size_t i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
    if (std::sqrt(dot(r, r)) <= rtol * bn) return it;
    for (size_t i = 0; i < n; ++i) z[i] = dinv.empty() ? r[i] : dinv[i] * r[i];
    const double rz_new = dot(r, z);
    const double beta = rz_new / rz;
    rz = rz_new;
    for (size_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
  }
  return -max_it;
}

}  // namespace

CoarseProblem::CoarseProblem(int pairing, int g) : pairing_(pairing), g_(g), m_(1 << g) {
  k_ = msfec_k(pairing);
  if (k_ < 0 || g < 0 || g > 8) throw std::invalid_argument("CoarseProblem: bad pairing / refinement level");
  switch (pairing) {
    case MSFEC_Q: kind0_ = KV; kind1_ = -1; break;
    case MSFEC_Q_NED: kind0_ = KV; kind1_ = KE; break;
    case MSFEC_NED_RT: kind0_ = KE; kind1_ = KF; break;
    default: kind0_ = KF; kind1_ = KC; break;
  }
  const Grid G{m_};
  n0_ = (int)G.count(kind0_); n1_ = (int)G.count(kind1_);
  essential_.assign(n0_ + n1_, 0);
  if (pairing == MSFEC_Q || pairing == MSFEC_Q_NED) {
    for (int k = 0; k <= m_; ++k) for (int j = 0; j <= m_; ++j) for (int i = 0; i <= m_; ++i)
      if (i == 0 || i == m_ || j == 0 || j == m_ || k == 0 || k == m_) essential_[G.V(i, j, k)] = 1;
  }
  if (pairing == MSFEC_Q_NED) {
    auto bnd = [&](int c) { return c == 0 || c == m_; };
    for (int k = 0; k <= m_; ++k) for (int j = 0; j <= m_; ++j) for (int i = 0; i <= m_; ++i) {
      if (i < m_ && (bnd(j) || bnd(k))) essential_[n0_ + G.E(0, i, j, k)] = 1;
      if (j < m_ && (bnd(i) || bnd(k))) essential_[n0_ + G.E(1, i, j, k)] = 1;
      if (k < m_ && (bnd(i) || bnd(j))) essential_[n0_ + G.E(2, i, j, k)] = 1;
    }
  }
  b_.assign(n0_ + n1_, 0.0);
  x_.assign(n0_ + n1_, 0.0);
}

void CoarseProblem::ijk_of(long long cell, int &i, int &j, int &k) const {
  i = j = k = 0;
  for (int b = 0; b < g_; ++b) {                           // p4est z-order: x is the fastest bit
    i |= (int)((cell >> (3 * b + 0)) & 1LL) << b;
    j |= (int)((cell >> (3 * b + 1)) & 1LL) << b;
    k |= (int)((cell >> (3 * b + 2)) & 1LL) << b;
  }
}

void CoarseProblem::cell_dofs(long long cell, int32_t *dofs) const {
  int i, j, k;
  ijk_of(cell, i, j, k);
  const Grid G{m_};
  G.local(kind0_, i, j, k, dofs);
  const int l0 = Grid::ldofs(kind0_);
  if (kind1_ >= 0) {
    G.local(kind1_, i, j, k, dofs + l0);
    for (int q = 0; q < Grid::ldofs(kind1_); ++q) dofs[l0 + q] += n0_;
  }
}

void CoarseProblem::add_cell(long long cell, const double *M, const double *r) {
  if (finalized_) throw std::logic_error("CoarseProblem: add_cell after solve");
  int32_t d[24];
  cell_dofs(cell, d);
  for (int a = 0; a < k_; ++a) {
    b_[d[a]] += r[a];
    for (int c = 0; c < k_; ++c) { ti_.push_back(d[a]); tj_.push_back(d[c]); tv_.push_back(M[a * k_ + c]); }
  }
}

void CoarseProblem::finalize() {
  const int N = n0_ + n1_;
  std::vector<int32_t> cnt(N + 1, 0);
  for (int32_t r : ti_) cnt[r + 1]++;
  for (int r = 0; r < N; ++r) cnt[r + 1] += cnt[r];
  std::vector<int32_t> order(ti_.size()), fill(cnt.begin(), cnt.end() - 1);
  for (size_t e = 0; e < ti_.size(); ++e) order[fill[ti_[e]]++] = (int32_t)e;
  ptr_.assign(N + 1, 0); col_.clear(); val_.clear();
  for (int r = 0; r < N; ++r) {
    std::sort(order.begin() + cnt[r], order.begin() + cnt[r + 1], [&](int32_t a, int32_t b) { return tj_[a] != tj_[b] ? tj_[a] < tj_[b] : a < b; });
    for (int e = cnt[r]; e < cnt[r + 1]; ++e) {
      const int32_t t = order[e];
      if (!col_.empty() && (int)col_.size() > ptr_[r] && col_.back() == tj_[t]) val_.back() += tv_[t];   // fixed summation order
      else { col_.push_back(tj_[t]); val_.push_back(tv_[t]); }
    }
    ptr_[r + 1] = (int32_t)col_.size();
  }
  ti_.clear(); tj_.clear(); tv_.clear(); ti_.shrink_to_fit(); tj_.shrink_to_fit(); tv_.shrink_to_fit();
  finalized_ = true;
}

std::string CoarseProblem::solve() {
  if (!finalized_) finalize();
  const int N = n0_ + n1_;
  std::vector<int32_t> fmap(N, -1), free;
  for (int d = 0; d < N; ++d) if (!essential_[d]) { fmap[d] = (int32_t)free.size(); free.push_back(d); }
  const int nf = (int)free.size();
  int nf0 = 0;
  for (int d : free) nf0 += d < n0_;
  std::vector<double> bf(nf);
  for (int i = 0; i < nf; ++i) bf[i] = b_[free[i]];     // homogeneous essential data: nothing moves to the rhs
  std::vector<double> xf(nf, 0.0);
  std::ostringstream info;
  if (nf <= dense_limit_) {
    // dense LU with partial pivoting
    std::vector<double> A((size_t)nf * nf, 0.0);
    for (int i = 0; i < nf; ++i) {
      const int r = free[i];
      for (int e = ptr_[r]; e < ptr_[r + 1]; ++e) if (fmap[col_[e]] >= 0) A[(size_t)i * nf + fmap[col_[e]]] = val_[e];
    }
    xf = bf;
    for (int p = 0; p < nf; ++p) {
      int piv = p;
      double best = std::fabs(A[(size_t)p * nf + p]);
      for (int r = p + 1; r < nf; ++r) if (std::fabs(A[(size_t)r * nf + p]) > best) { best = std::fabs(A[(size_t)r * nf + p]); piv = r; }
      if (!(best > 0)) throw std::runtime_error("coarse system is singular");
      if (piv != p) { for (int c = 0; c < nf; ++c) std::swap(A[(size_t)p * nf + c], A[(size_t)piv * nf + c]); std::swap(xf[p], xf[piv]); }
      const double inv = 1.0 / A[(size_t)p * nf + p];
#pragma omp parallel for if (nf - p > 256)
      for (int r = p + 1; r < nf; ++r) {
        const double l = A[(size_t)r * nf + p] * inv;
        if (l == 0.0) continue;
        double *ar = &A[(size_t)r * nf];
        const double *ap = &A[(size_t)p * nf];
        for (int c = p + 1; c < nf; ++c) ar[c] -= l * ap[c];
        xf[r] -= l * xf[p];
      }
    }
    for (int p = nf - 1; p >= 0; --p) {
      double s = xf[p];
      for (int c = p + 1; c < nf; ++c) s -= A[(size_t)p * nf + c] * xf[c];
      xf[p] = s / A[(size_t)p * nf + p];
    }
    info << "dense LU with partial pivoting, " << nf << " unknowns";
  } else {
    // blocks of the free system
    Csr B[2][2];
    const int nb[2] = {nf0, nf - nf0};
    for (int a = 0; a < 2; ++a) for (int c = 0; c < 2; ++c) { B[a][c].n_rows = nb[a]; B[a][c].n_cols = nb[c]; B[a][c].ptr.assign(nb[a] + 1, 0); }
    for (int i = 0; i < nf; ++i) {
      const int r = free[i], a = i < nf0 ? 0 : 1, li = a ? i - nf0 : i;
      for (int e = ptr_[r]; e < ptr_[r + 1]; ++e) {
        const int fj = fmap[col_[e]];
        if (fj < 0) continue;
        const int c = fj < nf0 ? 0 : 1;
        B[a][c].col.push_back(c ? fj - nf0 : fj); B[a][c].val.push_back(val_[e]);
      }
      for (int c = 0; c < 2; ++c) B[a][c].ptr[li + 1] = (int32_t)B[a][c].col.size();
    }
    if (device_ >= 0) {
      // the same nested iteration on the rank's GPU (csrc/coarse_dev.cu); no silent fallback: a failure is an error
      auto view = [](const Csr &M) { return msfec_coarse_csr{M.n_rows, M.n_cols, M.ptr.data(), M.col.data(), M.val.data()}; };
      const msfec_coarse_csr v00 = view(B[0][0]), v01 = view(B[0][1]), v10 = view(B[1][0]), v11 = view(B[1][1]);
      msfec_coarse_stats cs{};
      const bool two = nb[1] > 0;
      if (msfec_coarse_solve_device(device_, &v00, two ? &v01 : nullptr, two ? &v10 : nullptr, two ? &v11 : nullptr, bf.data(),
                                    two ? bf.data() + nb[0] : nullptr, 1e-13, 1e-11, xf.data(), two ? xf.data() + nb[0] : nullptr, &cs))
        throw std::runtime_error(std::string("coarse solve on the device: ") + msfec_coarse_last_error());
      if (two) info << "Schur-complement CG on device " << device_ << ", " << cs.outer_iterations << " outer iterations, " << cs.inner_iterations << " inner CG iterations";
      else info << "CG (Jacobi) on device " << device_ << " on the SPD coarse system, " << cs.inner_iterations << " iterations";
      info << ", " << cs.kernel_launches << " kernel launches, " << cs.ms_device << " ms";
    } else {
    std::vector<double> d0(nb[0]);
    for (int r = 0; r < nb[0]; ++r) {
      double d = 0;
      for (int e = B[0][0].ptr[r]; e < B[0][0].ptr[r + 1]; ++e) if (B[0][0].col[e] == r) d = B[0][0].val[e];
      d0[r] = d != 0 ? 1.0 / d : 1.0;
    }
    std::vector<double> b0(bf.begin(), bf.begin() + nb[0]), b1(bf.begin() + nb[0], bf.end());
    long inner_total = 0;
    auto inv00 = [&](const std::vector<double> &rhs, std::vector<double> &out) {
      const int it = cg([&](const std::vector<double> &v, std::vector<double> &y) { B[0][0].mult(v, y); }, d0, rhs, out, 1e-13, 20 * nb[0] + 100);
      if (it < 0) throw std::runtime_error("coarse solve: inner CG on block (0,0) did not converge");
      inner_total += it;
    };
    if (nb[1] == 0) {
      inv00(b0, xf);
      info << "CG (Jacobi) on the SPD coarse system, " << inner_total << " iterations";
    } else {
      // S u = b1 - A10 A00^-1 b0,   S = A11 - A10 A00^-1 A01  (SPD because A01 = -A10^T)
      std::vector<double> t0, t1, rhsS(nb[1]), u, sig;
      inv00(b0, t0);
      B[1][0].mult(t0, t1);
      for (int i = 0; i < nb[1]; ++i) rhsS[i] = b1[i] - t1[i];
      std::vector<double> ds(nb[1], 0.0);     // Jacobi for S: diag(A11) + diag(A10 diag(A00)^-1 A10^T)
      for (int r = 0; r < nb[1]; ++r) {
        double d = 0;
        for (int e = B[1][1].ptr[r]; e < B[1][1].ptr[r + 1]; ++e) if (B[1][1].col[e] == r) d += B[1][1].val[e];
        for (int e = B[1][0].ptr[r]; e < B[1][0].ptr[r + 1]; ++e) d += B[1][0].val[e] * B[1][0].val[e] * d0[B[1][0].col[e]];
        ds[r] = d > 0 ? 1.0 / d : 1.0;
      }
      std::vector<double> w0, w1, w2, w3;
      const int outer = cg([&](const std::vector<double> &v, std::vector<double> &y) {
        B[0][1].mult(v, w0); inv00(w0, w1); B[1][0].mult(w1, w2); B[1][1].mult(v, w3);
        y.resize(v.size());
        for (size_t i = 0; i < v.size(); ++i) y[i] = w3[i] - w2[i];
      }, ds, rhsS, u, 1e-11, 20 * nb[1] + 100);
      if (outer < 0) throw std::runtime_error("coarse solve: Schur-complement CG did not converge");
      B[0][1].mult(u, w0);
      for (int i = 0; i < nb[0]; ++i) w0[i] = b0[i] - w0[i];
      inv00(w0, sig);
      std::copy(sig.begin(), sig.end(), xf.begin());
      std::copy(u.begin(), u.end(), xf.begin() + nb[0]);
      info << "Schur-complement CG, " << outer << " outer iterations, " << inner_total << " inner CG iterations";
    }
    }
  }
  std::fill(x_.begin(), x_.end(), 0.0);
  for (int i = 0; i < nf; ++i) x_[free[i]] = xf[i];
  // true residual over the free unknowns
  double rn = 0, bn = 0;
  for (int i = 0; i < nf; ++i) {
    const int r = free[i];
    double s = b_[r];
    for (int e = ptr_[r]; e < ptr_[r + 1]; ++e) s -= val_[e] * x_[col_[e]];
    rn += s * s; bn += b_[r] * b_[r];
  }
  residual_ = bn > 0 ? std::sqrt(rn / bn) : std::sqrt(rn);
  info << ", relative residual " << residual_;
  return info.str();
}

void CoarseProblem::cell_weights(long long cell, double *w) const {
  int32_t d[24];
  cell_dofs(cell, d);
  for (int a = 0; a < k_; ++a) w[a] = x_[d[a]];
}

}  // namespace msfec
