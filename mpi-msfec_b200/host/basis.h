// Reference-shaped host classes on top of the batched C ABI (include/msfec.h).
//
// The reference constructs one XBasis object per locally owned coarse cell, stores them by value
// in a std::map<CellId, XBasis> and calls run() on each in a serial loop
// (source/Ned_RT/ned_rt_global.cc:61-98).  These classes keep that object model and the method
// names of include/Ned_RT/ned_rt_basis.h:99-153 (and the Q / Q_Ned / RT_DQ siblings) but defer
// the work to a BasisBatch shared by all cells of a rank: the first run() of any cell builds the
// bases of ALL registered cells in one msfec_build_basis call, later run() calls return at once.
#pragma once
#include <array>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/msfec.h"

namespace msfec {

struct ParametersMs {          // subset of NedRT::ParametersMs the basis path reads
  msfec_problem problem{};
  std::string filename_output = "Ms", dirname_output = "data-output";
  bool verbose = false, verbose_basis = false, prevent_output = false, write_first_basis = false;
  int n_refine_global = 2, n_refine_local = 1;
  // Reads the reference's .prm verbatim (ned_rt_parameters.cc:131-257).  Throws std::runtime_error.
  ParametersMs(const std::string &prm_filename, int pairing);
  ~ParametersMs();
  ParametersMs(const ParametersMs &) = delete;
  ParametersMs &operator=(const ParametersMs &) = delete;
};

// All locally owned coarse cells of one rank (one msfec_ctx on one GPU).
class BasisBatch {
 public:
  BasisBatch(const ParametersMs &prm, int device);
  ~BasisBatch();
  int add_cell(const std::array<std::array<double, 3>, 8> &corners, long long global_id);   // returns local index
  void build();                                   // idempotent
  bool built() const { return built_; }
  int k() const { return k_; }
  const double *matrix(int cell) const { return &M_[(size_t)cell * k_ * k_]; }
  const double *rhs(int cell) const { return &r_[(size_t)cell * k_]; }
  void set_weights(int cell, const std::vector<double> &w);
  void fine_solution(int cell, std::vector<double> &b0, std::vector<double> &b1);
  // Sum over this rank's cells of the squared fine-grid norms of the solution set by set_weights: L2 and natural
  // semi-norm (H1 / H(curl) / H(div)) of block 0, then of block 1 (msfec_solution_norms).  Ranks add these four
  // numbers (MPI_Allreduce / ncclAllReduce, sum, FP64) and take square roots.
  std::array<double, 4> solution_norms_squared();
  void basis_function(int cell, int index, std::vector<double> &b0, std::vector<double> &b1);
  // VTU (ParaView) file of fine-grid data of one coarse cell: what output_global_solution_in_cell /
  // output_basis write through deal.II DataOut (ned_rt_basis.cc:951-1031, 1092-1147).  Vector fields are
  // written as cell data evaluated at the fine-cell centres, nodal fields as point data.
  void write_vtu(int cell, const std::string &path, const std::vector<double> &b0, const std::vector<double> &b1);
  const msfec_stats &stats() const { return stats_; }
  int n_cells() const { return (int)ids_.size(); }

 private:
  msfec_ctx *ctx_ = nullptr;
  int k_ = 0, pairing_ = 0, L_ = 0;
  bool built_ = false, weights_dirty_ = false;
  std::vector<double> corners_, M_, r_, w_;
  std::vector<int64_t> ids_;
  msfec_stats stats_{};
  std::vector<double> pos_[2];
  std::vector<int32_t> axis_[2];
  void load_layout();
};

// Per-cell facade with the reference's interface.  Copyable before run(), like the reference
// (is_copyable, ned_rt_basis.cc:177).
template <int PAIRING>
class BasisT {
 public:
  BasisT(std::shared_ptr<BasisBatch> batch, const std::array<std::array<double, 3>, 8> &corners, long long cell_id,
         long long first_cell, unsigned local_subdomain)
      : batch_(std::move(batch)), first_cell_(first_cell), cell_id_(cell_id), subdomain_(local_subdomain) {
    local_ = batch_->add_cell(corners, cell_id);
  }
  void run() { batch_->build(); }                                              // ned_rt_basis.h:119
  // k x k row-major, sigma-type coarse DoFs first                             // ned_rt_basis.h:131
  const double *get_global_element_matrix() const { require(); return batch_->matrix(local_); }
  const double *get_global_element_rhs() const { require(); return batch_->rhs(local_); }     // :137
  void set_global_weights(const std::vector<double> &w) { require(); batch_->set_weights(local_, w); }   // :152
  void get_global_solution(std::vector<double> &block0, std::vector<double> &block1) { batch_->fine_solution(local_, block0, block1); }
  std::string get_filename_global() const {                                   // :143, naming of :1254-1259
    return "fine_solution.cell-" + std::to_string(cell_id_) + ".vtu";
  }
  int n_coarse_dofs() const { return batch_->k(); }

 private:
  void require() const { if (!batch_->built()) throw std::logic_error("basis not built: call run() first"); }
  std::shared_ptr<BasisBatch> batch_;
  int local_ = -1;
  long long first_cell_, cell_id_;
  unsigned subdomain_;
};

using QBasis = BasisT<MSFEC_Q>;
using QNedBasis = BasisT<MSFEC_Q_NED>;
using NedRTBasis = BasisT<MSFEC_NED_RT>;
using RTDQBasis = BasisT<MSFEC_RT_DQ>;

// p4est z-order enumeration of the uniformly refined unit cube and the contiguous chunk owned by
// `rank` (parallel::distributed::Triangulation + is_locally_owned, ned_rt_global.cc:12-15,61-63).
void morton_cell(int global_refinements, long long index, std::array<std::array<double, 3>, 8> &corners);
void owned_range(long long n_cells, int rank, int world, long long &lo, long long &hi);

// Shared body of the four executables (source/main_ned_rt.cxx:15-117): "-p file.prm".
int driver_main(int argc, char **argv, int pairing, const char *name);

}  // namespace msfec
