#pragma once
#include <deal.II/base/point.h>
#include <deal.II/lac/vector.h>
namespace dealii {
template <int dim>
class Function {
 public:
  explicit Function(unsigned int n_components = 1) : n_components(n_components) {}
  virtual ~Function() = default;
  virtual double value(const Point<dim> &, const unsigned int = 0) const { return 0.0; }
  virtual void value_list(const std::vector<Point<dim>> &points, std::vector<double> &values,
                          const unsigned int component = 0) const {
    for (std::size_t p = 0; p < points.size(); ++p) values[p] = value(points[p], component);
  }
  virtual void vector_value(const Point<dim> &, Vector<double> &) const {}
  virtual void vector_value_list(const std::vector<Point<dim>> &points, std::vector<Vector<double>> &values) const {
    for (std::size_t p = 0; p < points.size(); ++p) vector_value(points[p], values[p]);
  }
  const unsigned int n_components;
};
}  // namespace dealii
