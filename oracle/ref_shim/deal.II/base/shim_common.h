// Minimal stand-in for the handful of deal.II types the reference's closed-form data classes touch
// (equation_data/eqn_coeff_A, eqn_coeff_R, functions/basis_q1, basis_q1_grad).  TEST INFRASTRUCTURE ONLY:
// it exists so that oracle/Makefile can compile those reference sources, unmodified and where they lie
// under /root/reference, into oracle/_ref/libmsfec_ref.so -- the checker the Python oracle is pinned
// against (tests/test_reference_compiled.py).  deal.II itself is not installed in this image.  Nothing
// here is deal.II code; each class implements only the documented behaviour of the members used.
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

#define Assert(cond, exc)                                                              \
  do {                                                                                 \
    if (!(cond)) throw std::runtime_error(std::string("Assert failed: ") + #cond);    \
  } while (0)
#define ExcDimensionMismatch(a, b) 0

namespace dealii {
namespace numbers {
static constexpr double PI = 3.14159265358979323846264338327950288;
}
}  // namespace dealii
