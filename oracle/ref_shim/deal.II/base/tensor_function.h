#pragma once
#include <deal.II/base/point.h>
namespace dealii {
template <int rank, int dim>
class TensorFunction {
 public:
  TensorFunction() = default;
  virtual ~TensorFunction() = default;
  virtual Tensor<rank, dim> value(const Point<dim> &) const { return Tensor<rank, dim>(); }
  virtual void value_list(const std::vector<Point<dim>> &points, std::vector<Tensor<rank, dim>> &values) const {
    for (std::size_t p = 0; p < points.size(); ++p) values[p] = value(points[p]);
  }
};
}  // namespace dealii
