#pragma once
#include <deal.II/base/tensor.h>
namespace dealii {
template <int dim>
class Point : public Tensor<1, dim> {
 public:
  Point() = default;
  double operator()(unsigned int i) const { return (*this)[i]; }
  double &operator()(unsigned int i) { return (*this)[i]; }
};
}  // namespace dealii
