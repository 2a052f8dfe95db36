#pragma once
#include <deal.II/base/shim_common.h>
#include <deal.II/lac/vector.h>
namespace dealii {
template <typename Number>
class FullMatrix {
 public:
  FullMatrix() = default;
  FullMatrix(std::size_t m, std::size_t n) : m_(m), n_(n), a_(m * n, Number(0)) {}
  Number &operator()(std::size_t i, std::size_t j) { return a_[i * n_ + j]; }
  const Number &operator()(std::size_t i, std::size_t j) const { return a_[i * n_ + j]; }
  FullMatrix &operator=(Number s) { for (auto &x : a_) x = s; return *this; }
  // w = A v,  w = A^T v
  void vmult(Vector<Number> &w, const Vector<Number> &v) const {
    for (std::size_t i = 0; i < m_; ++i) { Number s = 0; for (std::size_t j = 0; j < n_; ++j) s += (*this)(i, j) * v(j); w(i) = s; }
  }
  void Tvmult(Vector<Number> &w, const Vector<Number> &v) const {
    for (std::size_t j = 0; j < n_; ++j) { Number s = 0; for (std::size_t i = 0; i < m_; ++i) s += (*this)(i, j) * v(i); w(j) = s; }
  }
  Number determinant() const {
    const FullMatrix &A = *this;
    if (m_ == 1 && n_ == 1) return A(0, 0);
    if (m_ == 2 && n_ == 2) return A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0);
    if (m_ == 3 && n_ == 3)
      return A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
             A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
    throw std::runtime_error("FullMatrix::determinant: only up to 3 x 3");
  }
  std::size_t m() const { return m_; }
  std::size_t n() const { return n_; }
  // *this = M^-1 (Gauss-Jordan with partial pivoting)
  void invert(const FullMatrix &M) {
    const std::size_t n = M.m_;
    if (M.n_ != n || m_ != n || n_ != n) throw std::runtime_error("FullMatrix::invert: shape");
    std::vector<long double> w(n * 2 * n, 0.0L);
    for (std::size_t i = 0; i < n; ++i) {
      for (std::size_t j = 0; j < n; ++j) w[i * 2 * n + j] = M(i, j);
      w[i * 2 * n + n + i] = 1.0L;
    }
    for (std::size_t c = 0; c < n; ++c) {
      std::size_t p = c;
      for (std::size_t r = c + 1; r < n; ++r)
        if (std::fabs((double)w[r * 2 * n + c]) > std::fabs((double)w[p * 2 * n + c])) p = r;
      if (w[p * 2 * n + c] == 0.0L) throw std::runtime_error("FullMatrix::invert: singular");
      if (p != c)
        for (std::size_t j = 0; j < 2 * n; ++j) std::swap(w[p * 2 * n + j], w[c * 2 * n + j]);
      const long double d = w[c * 2 * n + c];
      for (std::size_t j = 0; j < 2 * n; ++j) w[c * 2 * n + j] /= d;
      for (std::size_t r = 0; r < n; ++r) {
        if (r == c) continue;
        const long double f = w[r * 2 * n + c];
        if (f == 0.0L) continue;
        for (std::size_t j = 0; j < 2 * n; ++j) w[r * 2 * n + j] -= f * w[c * 2 * n + j];
      }
    }
    for (std::size_t i = 0; i < n; ++i)
      for (std::size_t j = 0; j < n; ++j) (*this)(i, j) = (Number)w[i * 2 * n + n + j];
  }
 private:
  std::size_t m_ = 0, n_ = 0;
  std::vector<Number> a_;
};
}  // namespace dealii
