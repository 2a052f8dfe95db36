// C entry points into the REFERENCE's own closed-form data classes, compiled unmodified from
// /root/reference (oracle/Makefile, target _ref/libmsfec_ref.so) against the stand-in headers in
// oracle/ref_shim/deal.II/.  TEST INFRASTRUCTURE: the checker the Python oracle's coefficient sampling and
// coarse Q1 shape functions are pinned against (tests/test_reference_compiled.py,
// tests/golden/make_reference_compiled_golden.py).  Never loaded by the product path.
//   Diffusion_A / DiffusionInverse_A : source/equation_data/eqn_coeff_A.cc:5-242
//   ReactionRate                     : source/equation_data/eqn_coeff_R.cc:7-14
//   BasisQ1<3> / BasisQ1Grad<3>      : include/functions/basis_q1.tpp:43-109, basis_q1_grad.tpp:43-107
#include <equation_data/eqn_coeff_A.h>
#include <equation_data/eqn_coeff_R.h>
#include <functions/basis_q1.h>
#include <functions/basis_q1_grad.h>

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
}
