// Global multiscale driver with the reference's structure (include/Ned_RT/ned_rt_global.h; source/Ned_RT/
// ned_rt_global.cc and the Q / Q_Ned / RT_DQ siblings): one class, the same member functions in the same order
//   make_grid -> initialize_and_compute_basis -> setup_system_matrix -> setup_constraints -> assemble_system ->
//   solve_iterative -> send_global_weights_to_cell -> output_results
// driven by run() (ned_rt_global.cc:704-771).  deal.II / p4est / Trilinos are replaced by: the structured coarse grid
// in z-order (contiguous chunks per rank = what p4est ownership yields on the uniformly refined cube), the batched B200
// basis build behind the reference's per-cell interface (basis.h), a host coarse solver (coarse.h) and NCCL for the two
// cross-rank steps (include/msfec_comm.h).
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/msfec_comm.h"
#include "basis.h"
#include "coarse.h"

namespace msfec {

template <int PAIRING>
class Multiscale {
 public:
  using Basis = BasisT<PAIRING>;
  // shared_comm: an NCCL communicator owned by the caller (the driver runs the fine-grid comparator and the multiscale
  // method on one communicator); nullptr: create one when world > 1.
  // parameters.standard: this object is the reference's XStd (source/Ned_RT/ned_rt_ref.cc and siblings): the same pipeline on
  // the mesh refined `Standard method parameters / Mesh / refinements` times with 0 local refinements -- the basis build
  // returns the standard lowest-order element matrices of every cell, the coarse solve IS the fine-grid solve.
  Multiscale(const ParametersMs &parameters, const std::string &parameter_filename, int rank, int world, int device,
             const char *name, msfec_comm *shared_comm = nullptr);
  ~Multiscale();
  void run();

  // results of the last run (tests, logging)
  const std::array<double, 4> &solution_norms() const { return norms_; }   // ||b0||_L2, |b0|_semi, ||b1||_L2, |b1|_semi (global)

 private:
  void make_grid();
  void initialize_and_compute_basis();
  void setup_system_matrix();
  void setup_constraints();
  void assemble_system();
  void solve_iterative();
  void send_global_weights_to_cell();
  std::vector<std::string> collect_filenames_on_mpi_process();
  void output_results();
  void compute_norms();

  const ParametersMs &parameters;
  std::string parameter_filename;
  int rank_, world_, device_;
  std::string name_;
  msfec_comm *comm_ = nullptr;
  bool owns_comm_ = false;
  long long n_global_cells_ = 0, lo_ = 0, hi_ = 0;
  CellId first_cell_;
  std::shared_ptr<BasisBatch> batch_;
  std::map<CellId, Basis> cell_basis_map;
  std::unique_ptr<CoarseProblem> coarse_;
  std::vector<double> all_elements_;          // [n_global_cells][k*k + k], gathered from all ranks
  std::array<double, 4> norms_{};
  double t_basis_ = 0, t_assemble_ = 0, t_solve_ = 0, t_output_ = 0;
};

}  // namespace msfec
