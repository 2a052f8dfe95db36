#pragma once
#include <deal.II/base/shim_common.h>
namespace dealii {
// Tensor<rank, dim>: value-initialised to zero, operator[] peels one rank (as in deal.II)
template <int rank, int dim>
class Tensor {
 public:
  Tensor() = default;
  Tensor<rank - 1, dim> &operator[](unsigned int i) { return v_[i]; }
  const Tensor<rank - 1, dim> &operator[](unsigned int i) const { return v_[i]; }
  void clear() { for (auto &t : v_) t.clear(); }
 private:
  std::array<Tensor<rank - 1, dim>, dim> v_{};
};
template <int dim>
class Tensor<1, dim> {
 public:
  Tensor() = default;
  double &operator[](unsigned int i) { return v_[i]; }
  const double &operator[](unsigned int i) const { return v_[i]; }
  void clear() { v_.fill(0.0); }
 private:
  std::array<double, dim> v_{};
};
// contraction of the last index of a with the first of b (rank 2 x rank 2)
template <int dim>
Tensor<2, dim> operator*(const Tensor<2, dim> &a, const Tensor<2, dim> &b) {
  Tensor<2, dim> c;
  for (int i = 0; i < dim; ++i)
    for (int j = 0; j < dim; ++j) {
      double s = 0.0;
      for (int k = 0; k < dim; ++k) s += a[i][k] * b[k][j];
      c[i][j] = s;
    }
  return c;
}
template <int dim>
Tensor<2, dim> transpose(const Tensor<2, dim> &a) {
  Tensor<2, dim> c;
  for (int i = 0; i < dim; ++i)
    for (int j = 0; j < dim; ++j) c[i][j] = a[j][i];
  return c;
}
}  // namespace dealii
