#pragma once
#include <deal.II/base/shim_common.h>
namespace dealii {
template <typename Number>
class FullMatrix {
 public:
  FullMatrix() = default;
  FullMatrix(std::size_t m, std::size_t n) : m_(m), n_(n), a_(m * n, Number(0)) {}
  Number &operator()(std::size_t i, std::size_t j) { return a_[i * n_ + j]; }
  const Number &operator()(std::size_t i, std::size_t j) const { return a_[i * n_ + j]; }
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
