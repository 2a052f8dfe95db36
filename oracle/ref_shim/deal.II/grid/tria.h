#pragma once
#include <deal.II/base/point.h>
namespace dealii {
// Only what BasisQ1 / BasisQ1Grad ask of a cell: vertex(i), i in deal.II's lexicographic vertex order
// (x fastest: vertex i has reference coordinates (i & 1, (i >> 1) & 1, i >> 2)).
template <int dim>
struct ShimCell {
  std::array<Point<dim>, (1u << dim)> vertices;
  const Point<dim> &vertex(unsigned int i) const { return vertices[i]; }
};
template <int dim>
struct ShimCellIterator {
  const ShimCell<dim> *cell = nullptr;
  const ShimCell<dim> *operator->() const { return cell; }
};
template <int dim>
class Triangulation {
 public:
  using active_cell_iterator = ShimCellIterator<dim>;
};
}  // namespace dealii
