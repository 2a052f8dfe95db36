#pragma once
#include <deal.II/base/point.h>
namespace dealii {
template <int dim>
struct GeometryInfo {
  // vertex i of the unit cell: coordinates (i & 1, (i >> 1) & 1, i >> 2), x fastest
  static Point<dim> unit_cell_vertex(unsigned int i) {
    Point<dim> p;
    for (int d = 0; d < dim; ++d) p(d) = (i >> d) & 1u;
    return p;
  }
};
}  // namespace dealii
