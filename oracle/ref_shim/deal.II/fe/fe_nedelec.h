#pragma once
#include <deal.II/base/point.h>
namespace dealii {
// Lowest-order Nedelec element on the unit cube, restated from deal.II's documented conventions: one shape
// function per line, tangential component 1 on its own line and 0 on the others, t in the +coordinate direction.  Lines:
// bottom face z = 0 {0: x=0 ||y, 1: x=1 ||y, 2: y=0 ||x, 3: y=1 ||x}, top face 4..7 alike, vertical 8..11 at (0,0),(1,0),(0,1),(1,1).
// NOTE: this stand-in is NOT reference code; the reference classes compiled against it contribute the mapping and the
// covariant transform (basis_nedelec.tpp:43-54), which is what tests/test_reference_compiled.py checks with it.
template <int dim>
class FE_Nedelec {
 public:
  explicit FE_Nedelec(unsigned int degree) { if (degree != 0) throw std::runtime_error("FE_Nedelec stand-in: lowest order only"); }
  double shape_value_component(unsigned int i, const Point<dim> &p, unsigned int component) const {
    if (dim != 3 || i >= 12) throw std::runtime_error("FE_Nedelec stand-in: 3D, 12 lines");
    static const int dir[12] = {1, 1, 0, 0, 1, 1, 0, 0, 2, 2, 2, 2};
    // transverse position of the line: bits for the two other axes in increasing axis order
    static const int tb[12][2] = {{0, 0}, {1, 0}, {0, 0}, {1, 0}, {0, 1}, {1, 1}, {0, 1}, {1, 1}, {0, 0}, {1, 0}, {0, 1}, {1, 1}};
    if ((int)component != dir[i]) return 0.0;
    double v = 1.0;
    int t = 0;
    for (int a = 0; a < 3; ++a) {
      if (a == dir[i]) continue;
      v *= tb[i][t] ? p(a) : 1.0 - p(a);
      ++t;
    }
    return v;
  }
};
}  // namespace dealii
