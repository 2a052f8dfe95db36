// C entry points into the REFERENCE's own closed-form data classes, compiled unmodified from
// /root/reference (oracle/Makefile, target _ref/libmsfec_ref.so) against the stand-in headers in
// oracle/ref_shim/deal.II/.  TEST INFRASTRUCTURE: the checker the Python oracle's coefficient sampling and
// coarse Q1 shape functions are pinned against (tests/test_reference_compiled.py,
// tests/golden/make_reference_compiled_golden.py).  Never loaded by the product path.
//   Diffusion_A / DiffusionInverse_A : source/equation_data/eqn_coeff_A.cc:5-242
//   ReactionRate                     : source/equation_data/eqn_coeff_R.cc:7-14
//   BasisQ1<3> / BasisQ1Grad<3>      : include/functions/basis_q1.tpp:43-109, basis_q1_grad.tpp:43-107
//   MyMappingQ1<3>                   : include/functions/my_mapping_q1.tpp:268-582 (real <-> unit cell, Jacobians)
//   BasisNedelec<3>                  : include/functions/basis_nedelec.tpp:40-86  (covariant transform J^-T phi)
//   BasisRaviartThomas<3>            : include/functions/basis_raviart_thomas.tpp:38-88 (Piola transform J phi / det J)
//     -- for these two the reference shape functions on the UNIT cell come from deal.II's FE_Nedelec / FE_RaviartThomas,
//     which are stand-ins here (ref_shim/deal.II/fe): what the compiled reference contributes is mapping + transform.
#include <equation_data/eqn_coeff_A.h>
#include <equation_data/eqn_coeff_R.h>
#include <functions/basis_q1.h>
#include <functions/basis_q1_grad.h>
#include <functions/basis_nedelec.h>
#include <functions/basis_raviart_thomas.h>
#include <functions/my_mapping_q1.h>

#include <cstring>
#include <memory>

using namespace dealii;

static std::string g_err;
static std::vector<Point<3>> to_points(int n, const double *xyz) {
  std::vector<Point<3>> p(n);
  for (int i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) p[i](d) = xyz[3 * i + d];
  return p;
}
static ShimCell<3> to_cell(const double *vertices) {
  ShimCell<3> c;
  for (int v = 0; v < 8; ++v)
    for (int d = 0; d < 3; ++d) c.vertices[v](d) = vertices[3 * v + d];
  return c;
}

extern "C" {
const char *msfec_ref_last_error() { return g_err.c_str(); }

// out[n][3][3]; mode 0: Diffusion_A::value_list, 1: DiffusionInverse_A::value_list, 2 / 3: the same through value()
int msfec_ref_diffusion_a(const char *prm, int mode, int n, const double *xyz, double *out) {
  try {
    const std::vector<Point<3>> pts = to_points(n, xyz);
    std::vector<Tensor<2, 3>> val(n);
    std::unique_ptr<TensorFunction<2, 3>> f;
    if (mode & 1) f.reset(new EquationData::DiffusionInverse_A(prm));
    else f.reset(new EquationData::Diffusion_A(prm));
    if (mode & 2) for (int i = 0; i < n; ++i) val[i] = f->value(pts[i]);
    else f->value_list(pts, val);
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) out[9 * i + 3 * a + b] = val[i][a][b];
    return 0;
  } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

int msfec_ref_reaction_rate(int n, const double *xyz, double *out) {
  try {
    std::vector<double> val(n, -1.0);
    EquationData::ReactionRate().value_list(to_points(n, xyz), val);
    std::memcpy(out, val.data(), sizeof(double) * n);
    return 0;
  } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

// vertices[8][3] in deal.II vertex order; out[n]
int msfec_ref_basis_q1(const double *vertices, int index, int n, const double *xyz, double *out) {
  try {
    const ShimCell<3> cell = to_cell(vertices);
    ShapeFun::BasisQ1<3> basis(Triangulation<3>::active_cell_iterator{&cell});
    basis.set_index(index);
    std::vector<double> val(n);
    basis.value_list(to_points(n, xyz), val);
    std::memcpy(out, val.data(), sizeof(double) * n);
    return 0;
  } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

// out[n][3]
int msfec_ref_basis_q1_grad(const double *vertices, int index, int n, const double *xyz, double *out) {
  try {
    const ShimCell<3> cell = to_cell(vertices);
    ShapeFun::BasisQ1Grad<3> basis(Triangulation<3>::active_cell_iterator{&cell});
    basis.set_index(index);
    std::vector<Tensor<1, 3>> val(n);
    basis.tensor_value_list(to_points(n, xyz), val);
    for (int i = 0; i < n; ++i)
      for (int d = 0; d < 3; ++d) out[3 * i + d] = val[i][d];
    return 0;
  } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

// MyMappingQ1<3> of the cell: unit-cell coordinates ref[n][3] and the Jacobian of real -> unit, inv_jac[n][3][3]
int msfec_ref_mapping(const double *vertices, int n, const double *xyz, double *ref, double *inv_jac) {
  try {
    const ShimCell<3> cell = to_cell(vertices);
    ShapeFun::MyMappingQ1<3> mapping(Triangulation<3>::active_cell_iterator{&cell});
    const std::vector<Point<3>> pts = to_points(n, xyz);
    std::vector<Point<3>> unit(n);
    mapping.map_real_to_unit_cell(pts, unit);
    std::vector<FullMatrix<double>> jac(n, FullMatrix<double>(3, 3));
    mapping.jacobian_map_real_to_unit_cell(pts, jac);
    for (int i = 0; i < n; ++i) {
      for (int d = 0; d < 3; ++d) ref[3 * i + d] = unit[i](d);
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) inv_jac[9 * i + 3 * a + b] = jac[i](a, b);
    }
    return 0;
  } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

// which: 0 BasisNedelec<3> (12 functions), 1 BasisRaviartThomas<3> (6); out[n][3]; list != 0: vector_value_list
int msfec_ref_basis_vector(const double *vertices, int which, int index, int list, int n, const double *xyz, double *out) {
  try {
    const ShimCell<3> cell = to_cell(vertices);
    const Triangulation<3>::active_cell_iterator it{&cell};
    const std::vector<Point<3>> pts = to_points(n, xyz);
    std::vector<Vector<double>> val(n, Vector<double>(3));
    auto eval = [&](auto &basis) {
      basis.set_index(index);
      if (list) basis.vector_value_list(pts, val);
      else for (int i = 0; i < n; ++i) basis.vector_value(pts[i], val[i]);
    };
    if (which == 0) { ShapeFun::BasisNedelec<3> b(it, 0); eval(b); }
    else { ShapeFun::BasisRaviartThomas<3> b(it, 0); eval(b); }
    for (int i = 0; i < n; ++i)
      for (int d = 0; d < 3; ++d) out[3 * i + d] = val[i](d);
    return 0;
  } catch (const std::exception &e) { g_err = e.what(); return 1; }
}
}
