#pragma once
#include <deal.II/base/tensor.h>
namespace dealii {
template <int dim>
class Point : public Tensor<1, dim> {
 public:
  Point() = default;
  Point(double x, double y) { static_assert(dim == 2, "Point(x, y)"); (*this)[0] = x; (*this)[1] = y; }
  Point(double x, double y, double z) { static_assert(dim == 3, "Point(x, y, z)"); (*this)[0] = x; (*this)[1] = y; (*this)[2] = z; }
  double operator()(unsigned int i) const { return (*this)[i]; }
  double &operator()(unsigned int i) { return (*this)[i]; }
  Point &operator+=(const Point &o) { for (int d = 0; d < dim; ++d) (*this)[d] += o[d]; return *this; }
};
template <int dim>
Point<dim> operator+(const Point<dim> &a, const Point<dim> &b) { Point<dim> r = a; r += b; return r; }
template <int dim>
Point<dim> operator*(const Point<dim> &p, double s) {
  Point<dim> r;
  for (int d = 0; d < dim; ++d) r(d) = s * p(d);
  return r;
}
template <int dim>
Point<dim> operator*(double s, const Point<dim> &p) {
  Point<dim> r;
  for (int d = 0; d < dim; ++d) r(d) = s * p(d);
  return r;
}
}  // namespace dealii
