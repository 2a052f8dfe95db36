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
  bool compute_solution = true;   // "Control flow / compute solution" of the section read
  bool standard = false;          // true: these are the reference's ParametersStd (ned_rt_parameters.cc:8-125)
  // Reads the reference's .prm verbatim (ned_rt_parameters.cc:131-257).  Throws std::runtime_error.
  // standard = true reads `subsection Standard method parameters` instead (the fine-grid comparator XStd): its
  // `Mesh / refinements` become the global refinements, there are 0 local refinements (every cell of that mesh is its own
  // fine grid: the basis build returns the standard element matrices), its output names replace the multiscale ones.
  ParametersMs(const std::string &prm_filename, int pairing, bool standard = false);
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
  int pairing() const { return pairing_; }
  double seconds_per_cell() const { return built_ && !ids_.empty() ? stats_.ms_total * 1e-3 / ids_.size() : 0.0; }

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

// Minimal stand-ins for the deal.II return types of the reference's getters (FullMatrix<double>, Vector<double>): views
// into the batch's host buffers, valid as long as the batch lives.
struct FullMatrix {
  int n_rows = 0, n_cols = 0;
  const double *data = nullptr;
  double operator()(int i, int j) const { return data[(size_t)i * n_cols + j]; }
  int m() const { return n_rows; }
  int n() const { return n_cols; }
};
struct Vector {
  int n = 0;
  const double *data = nullptr;
  double operator()(int i) const { return data[i]; }
  double operator[](int i) const { return data[i]; }
  int size() const { return n; }
};

// deal.II CellId of an active cell of the uniformly refined unit cube: "<coarse cell>_<levels>:<child indices>" with the
// child index of every level = one octal digit of the z-order index, most significant first (CellId::to_string).
struct CellId {
  long long index = 0;
  int levels = 0;
  CellId() = default;
  CellId(long long morton_index, int global_refinements) : index(morton_index), levels(global_refinements) {}
  std::string to_string() const;
  bool operator<(const CellId &o) const { return index < o.index; }
  bool operator==(const CellId &o) const { return index == o.index && levels == o.levels; }
};

// what the reference passes as Triangulation<3>::active_cell_iterator: the 8 vertices (deal.II order) and the id
struct CoarseCell {
  std::array<std::array<double, 3>, 8> vertices{};
  CellId id;
  CoarseCell() = default;
  CoarseCell(int global_refinements, long long morton_index);
};

// Per-cell class with the reference's interface (include/Ned_RT/ned_rt_basis.h:99-153 and the Q / Q_Ned / RT_DQ
// siblings).  Copyable before run(), like the reference (is_copyable, ned_rt_basis.cc:177): *Multiscale stores the
// objects by value in a std::map<CellId, XBasis> (ned_rt_global.cc:61-77).  The MPI communicator argument of the
// reference is replaced by the rank's BasisBatch (one msfec_ctx on one GPU): the first run() of any cell builds the
// bases of ALL cells registered with the batch in one msfec_build_basis call, later run() calls return at once.
template <int PAIRING>
class BasisT {
 public:
  BasisT(const ParametersMs &parameters_ms, const std::string &parameter_filename, const CoarseCell &global_cell,
         const CellId &first_cell, unsigned int local_subdomain, std::shared_ptr<BasisBatch> batch)
      : prm_(&parameters_ms), parameter_filename_(parameter_filename), batch_(std::move(batch)), cell_(global_cell),
        first_cell_(first_cell), subdomain_(local_subdomain), filename_global_(parameters_ms.filename_output) {
    local_ = batch_->add_cell(global_cell.vertices, global_cell.id.index);
  }
  BasisT(const BasisT &) = default;

  void run();                                                                        // ned_rt_basis.h:119
  void output_global_solution_in_cell();                                             // :125, ned_rt_basis.cc:1092-1147
  const FullMatrix &get_global_element_matrix() const { require(); return M_; }      // :131  k x k, sigma-type DoFs first
  const Vector &get_global_element_rhs() const { require(); return r_; }             // :137
  const std::string &get_filename_global() const { return filename_global_; }        // :143
  void set_global_weights(const std::vector<double> &global_weights) {               // :152, ned_rt_basis.cc:1156-1181
    require(); batch_->set_weights(local_, global_weights);
  }
  // not in the reference: the reconstructed fine-scale solution of this cell (what the .vtu is written from)
  void get_global_solution(std::vector<double> &block0, std::vector<double> &block1) { batch_->fine_solution(local_, block0, block1); }
  const CellId &id() const { return cell_.id; }
  static const char *basis_file_stem();                                              // "basis_ned-rt" ...
  static std::string basis_file_tag(int n_basis, int &index_in_family);              // ".curl" / ".div" / ".h1" / ""

 private:
  void require() const { if (!batch_->built()) throw std::logic_error("basis not built: call run() first"); }
  void output_basis();                                                               // ned_rt_basis.cc:951-1031
  const ParametersMs *prm_;
  std::string parameter_filename_;
  std::shared_ptr<BasisBatch> batch_;
  CoarseCell cell_;
  CellId first_cell_;
  unsigned int subdomain_;
  int local_ = -1;
  std::string filename_global_;
  FullMatrix M_;
  Vector r_;
};

using QBasis = BasisT<MSFEC_Q>;
using QNedBasis = BasisT<MSFEC_Q_NED>;
using NedRTBasis = BasisT<MSFEC_NED_RT>;
using RTDQBasis = BasisT<MSFEC_RT_DQ>;

std::string int_to_string(long long value, int digits);      // Utilities::int_to_string

// p4est z-order enumeration of the uniformly refined unit cube and the contiguous chunk owned by
// `rank` (parallel::distributed::Triangulation + is_locally_owned, ned_rt_global.cc:12-15,61-63).
void morton_cell(int global_refinements, long long index, std::array<std::array<double, 3>, 8> &corners);
void owned_range(long long n_cells, int rank, int world, long long &lo, long long &hi);

// Shared body of the four executables (source/main_ned_rt.cxx:15-117): "-p file.prm" -> XMultiscale(prm).run().
int driver_main(int argc, char **argv, int pairing, const char *name);

}  // namespace msfec
