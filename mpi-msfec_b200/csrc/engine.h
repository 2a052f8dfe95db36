// Device engine of the basis build (declarations shared by engine.cu and capi.cpp).
#pragma once
#include <stdexcept>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/msfec.h"
#include "expr.h"
#include "mfplan.h"
#include "topology.h"

namespace msfec {

constexpr int kLanes = 32;   // coarse cells interleaved per group: lane == cell

struct CoefParams {           // passed by value to the sampling kernel
  double rot[9];
  double a_scale[3], a_alpha[3];
  int a_freq[3];
  int tensor_inverse, scalar_inverse;
  int use_random;
  unsigned long long seed;
  double sigma;
  int b_prog_off, b_prog_len;
  int rhs_prog_off[3], rhs_prog_len[3];
  int rhs_ncomp;
  int n, nC;
};

struct ProblemSpec {          // owned copy of msfec_problem with strings resolved
  msfec_problem p;
  std::string b_expr, rhs_expr, rhs_consts;
  std::vector<ExprInstr> programs;   // concatenated: B, rhs components
  CoefParams coef;
};

class Engine;   // defined in engine.cu

// no usable sm_100 device (-> MSFEC_ENODEVICE at the C ABI; there is no CPU fallback)
struct NoDeviceError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// Factory; throws NoDeviceError, std::invalid_argument, std::runtime_error (CUDA).  device >= 0.
Engine *engine_create(int device, const ProblemSpec &spec, const Topology &topo, const DirectPlan &plan, const MfPlan &mf);
void engine_destroy(Engine *e);
int engine_build(Engine *e, int n_cells, const double *corners, const int64_t *cell_ids, double *elem_matrix,
                 double *elem_rhs, bool device_ptrs, msfec_stats *stats, std::string &err);
int engine_set_weights(Engine *e, int n_cells, const double *weights, std::string &err);
int engine_get_fine_solution(Engine *e, int cell, double *b0, double *b1, std::string &err);
int engine_solution_norms(Engine *e, int n_cells, double *norms, std::string &err);
int engine_get_basis(Engine *e, int cell, int basis, double *b0, double *b1, std::string &err);
int engine_cell_values(Engine *e, int cell, double *values, size_t *count, std::string &err);

}  // namespace msfec
