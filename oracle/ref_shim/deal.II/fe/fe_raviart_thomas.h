#pragma once
#include <deal.II/base/point.h>
namespace dealii {
// Lowest-order Raviart-Thomas element on the unit cube: one shape function per face, normal component 1
// on its own face and 0 on the opposite one, n in the +coordinate direction on BOTH faces of a pair; faces x0 x1 y0 y1 z0 z1.
// NOTE: stand-in, not reference code (see fe_nedelec.h): the compiled reference contributes mapping and Piola transform.
template <int dim>
class FE_RaviartThomas {
 public:
  explicit FE_RaviartThomas(unsigned int degree) { if (degree != 0) throw std::runtime_error("FE_RaviartThomas stand-in: lowest order only"); }
  double shape_value_component(unsigned int i, const Point<dim> &p, unsigned int component) const {
    if (dim != 3 || i >= 6) throw std::runtime_error("FE_RaviartThomas stand-in: 3D, 6 faces");
    const unsigned int axis = i / 2, side = i % 2;
    if (component != axis) return 0.0;
    return side ? p(axis) : 1.0 - p(axis);
  }
};
}  // namespace dealii
