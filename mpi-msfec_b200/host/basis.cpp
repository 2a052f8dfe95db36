#include "basis.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sys/stat.h>
#include <unistd.h>

namespace msfec {

namespace {
// "set <key> = value" directly inside "subsection Multiscale method parameters"
std::string ms_top_level_value(const std::string &file, const std::string &key, const std::string &dflt) {
  std::ifstream in(file);
  std::string line, value = dflt;
  std::vector<std::string> stack;
  auto trim = [](std::string s) {
    const size_t b = s.find_first_not_of(" \t\r"), e = s.find_last_not_of(" \t\r");
    return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
  };
  while (std::getline(in, line)) {
    line = trim(line.substr(0, line.find('#')));
    if (line.rfind("subsection", 0) == 0) stack.push_back(trim(line.substr(10)));
    else if (line == "end") { if (!stack.empty()) stack.pop_back(); }
    else if (line.rfind("set ", 0) == 0 && stack.size() == 1 && stack[0] == "Multiscale method parameters") {
      const size_t eq = line.find('=');
      if (eq != std::string::npos && trim(line.substr(4, eq - 4)) == key) value = trim(line.substr(eq + 1));
    }
  }
  return value;
}
// "set <key> = true|false" inside "Multiscale method parameters / Control flow"
bool ms_control_flag(const std::string &file, const std::string &key, bool dflt) {
  std::ifstream in(file);
  std::string line;
  std::vector<std::string> stack;
  bool value = dflt;
  auto trim = [](std::string s) {
    const size_t b = s.find_first_not_of(" \t\r"), e = s.find_last_not_of(" \t\r");
    return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
  };
  while (std::getline(in, line)) {
    line = trim(line.substr(0, line.find('#')));
    if (line.rfind("subsection", 0) == 0) stack.push_back(trim(line.substr(10)));
    else if (line == "end") { if (!stack.empty()) stack.pop_back(); }
    else if (line.rfind("set ", 0) == 0 && stack.size() == 2 && stack[0] == "Multiscale method parameters" && stack[1] == "Control flow") {
      const size_t eq = line.find('=');
      if (eq != std::string::npos && trim(line.substr(4, eq - 4)) == key) value = trim(line.substr(eq + 1)) == "true";
    }
  }
  return value;
}

// "set <key> = value" inside the nested subsections `path` (exact nesting)
std::string prm_value(const std::string &file, const std::vector<std::string> &path, const std::string &key, const std::string &dflt) {
  std::ifstream in(file);
  std::string line, value = dflt;
  std::vector<std::string> stack;
  auto trim = [](std::string s) {
    const size_t b = s.find_first_not_of(" \t\r"), e = s.find_last_not_of(" \t\r");
    return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
  };
  while (std::getline(in, line)) {
    line = trim(line.substr(0, line.find('#')));
    if (line.rfind("subsection", 0) == 0) stack.push_back(trim(line.substr(10)));
    else if (line == "end") { if (!stack.empty()) stack.pop_back(); }
    else if (line.rfind("set ", 0) == 0 && stack == path) {
      const size_t eq = line.find('=');
      if (eq != std::string::npos && trim(line.substr(4, eq - 4)) == key) value = trim(line.substr(eq + 1));
    }
  }
  return value;
}
}  // namespace

ParametersMs::ParametersMs(const std::string &prm_filename, int pairing, bool standard_method) : standard(standard_method) {
  if (msfec_problem_from_prm(prm_filename.c_str(), pairing, &problem))
    throw std::runtime_error(std::string("parameter file: ") + msfec_last_error(nullptr));
  if (standard) {
    // ParametersStd (ned_rt_parameters.cc:8-125): defaults n_refine = 3, compute solution = true
    const std::vector<std::string> top{"Standard method parameters"}, mesh{top[0], "Mesh"}, flow{top[0], "Control flow"};
    static const char *dflt_name[4] = {"Q_Std", "Q_NED_Std", "NED_RT_Std", "RT_DQ_Std"};
    const std::string r = prm_value(prm_filename, mesh, "refinements", "3");
    char *end = nullptr;
    const long refinements = std::strtol(r.c_str(), &end, 10);
    if (end == r.c_str() || *end != 0 || refinements < 1 || refinements > 8)
      throw std::runtime_error("parameter file: 'Standard method parameters/Mesh/refinements' must be an integer in [1, 8]");
    problem.n_refine_global = (int)refinements;
    problem.n_refine_local = 0;
    problem.verbose_basis = 0;
    compute_solution = prm_value(prm_filename, flow, "compute solution", "true") == "true";
    verbose = prm_value(prm_filename, flow, "verbose", "false") == "true";
    filename_output = prm_value(prm_filename, top, "filename output", dflt_name[pairing]);
    dirname_output = prm_value(prm_filename, top, "dirname output", dirname_output);
    write_first_basis = false;
  } else {
    filename_output = ms_top_level_value(prm_filename, "filename output", filename_output);   // ned_rt_parameters.cc:226-236
    dirname_output = ms_top_level_value(prm_filename, "dirname output", dirname_output);
    write_first_basis = ms_control_flag(prm_filename, "write first basis", false);
    compute_solution = ms_control_flag(prm_filename, "compute solution", true);
  }
  n_refine_global = problem.n_refine_global;
  n_refine_local = problem.n_refine_local;
  verbose_basis = problem.verbose_basis != 0;
}
ParametersMs::~ParametersMs() { msfec_problem_free(&problem); }

BasisBatch::BasisBatch(const ParametersMs &prm, int device) {
  pairing_ = prm.problem.pairing;
  L_ = prm.problem.n_refine_local;
  k_ = msfec_k(pairing_);
  if (msfec_create(device, &prm.problem, &ctx_)) throw std::runtime_error(msfec_last_error(nullptr));
}
BasisBatch::~BasisBatch() { msfec_destroy(ctx_); }

int BasisBatch::add_cell(const std::array<std::array<double, 3>, 8> &c, long long id) {
  if (built_) throw std::logic_error("BasisBatch: cannot add cells after the build");
  for (auto &v : c) for (double x : v) corners_.push_back(x);
  ids_.push_back(id);
  return (int)ids_.size() - 1;
}

void BasisBatch::build() {
  if (built_) return;
  const int n = (int)ids_.size();
  M_.resize((size_t)n * k_ * k_); r_.resize((size_t)n * k_); w_.assign((size_t)n * k_, 0.0);
  const int rc = msfec_build_basis(ctx_, n, corners_.data(), ids_.data(), M_.data(), r_.data(), &stats_);
  if (rc) throw std::runtime_error(std::string("msfec_build_basis: ") + msfec_last_error(ctx_));
  built_ = true;
}

void BasisBatch::set_weights(int cell, const std::vector<double> &w) {
  if ((int)w.size() != k_) throw std::invalid_argument("set_global_weights: wrong number of weights");
  std::memcpy(&w_[(size_t)cell * k_], w.data(), k_ * sizeof(double));
  weights_dirty_ = true;
}

void BasisBatch::fine_solution(int cell, std::vector<double> &b0, std::vector<double> &b1) {
  if (weights_dirty_) {
    if (msfec_set_weights(ctx_, (int)ids_.size(), w_.data())) throw std::runtime_error(msfec_last_error(ctx_));
    weights_dirty_ = false;
  }
  int n0 = 0, n1 = 0;
  msfec_n_fine_dofs(pairing_, L_, &n0, &n1);
  b0.resize(n0); b1.resize(n1);
  if (msfec_get_fine_solution(ctx_, cell, b0.data(), n1 ? b1.data() : nullptr)) throw std::runtime_error(msfec_last_error(ctx_));
}

std::array<double, 4> BasisBatch::solution_norms_squared() {
  if (weights_dirty_) {
    if (msfec_set_weights(ctx_, (int)ids_.size(), w_.data())) throw std::runtime_error(msfec_last_error(ctx_));
    weights_dirty_ = false;
  }
  std::vector<double> per_cell(ids_.size() * 4);
  if (msfec_solution_norms(ctx_, (int)ids_.size(), per_cell.data())) throw std::runtime_error(msfec_last_error(ctx_));
  std::array<double, 4> sum{0, 0, 0, 0};
  for (size_t c = 0; c < ids_.size(); ++c) for (int i = 0; i < 4; ++i) sum[i] += per_cell[4 * c + i];
  return sum;
}

void BasisBatch::basis_function(int cell, int index, std::vector<double> &b0, std::vector<double> &b1) {
  int n0 = 0, n1 = 0;
  msfec_n_fine_dofs(pairing_, L_, &n0, &n1);
  b0.resize(n0); b1.resize(n1);
  if (msfec_get_basis(ctx_, cell, index, b0.data(), n1 ? b1.data() : nullptr)) throw std::runtime_error(msfec_last_error(ctx_));
}

void BasisBatch::load_layout() {
  if (!pos_[0].empty()) return;
  int n[2] = {0, 0};
  msfec_n_fine_dofs(pairing_, L_, &n[0], &n[1]);
  for (int b = 0; b < 2; ++b) {
    if (!n[b]) continue;
    pos_[b].resize((size_t)n[b] * 3); axis_[b].resize(n[b]);
    if (msfec_fine_dof_layout(ctx_, b, pos_[b].data(), axis_[b].data(), nullptr)) throw std::runtime_error(msfec_last_error(nullptr));
  }
}

void BasisBatch::write_vtu(int cell, const std::string &path, const std::vector<double> &b0, const std::vector<double> &b1) {
  load_layout();
  const int n = 1 << L_, n1 = n + 1;
  const double *c = &corners_[(size_t)cell * 24];
  const double H = c[21] - c[0], h = H / n;
  // lookup: doubled integer position -> DoF index, per block
  std::map<long long, int> key[2];
  auto K = [&](double x, double y, double z) { return ((long long)std::llround(2 * x) * 4096 + std::llround(2 * y)) * 4096 + std::llround(2 * z); };
  for (int b = 0; b < 2; ++b)
    for (size_t i = 0; i < axis_[b].size(); ++i) key[b][K(pos_[b][3 * i], pos_[b][3 * i + 1], pos_[b][3 * i + 2])] = (int)i;
  auto edge_vec = [&](const std::vector<double> &v, int blk, int i, int j, int k, double out[3]) {   // Nedelec at the cell centre
    for (int d = 0; d < 3; ++d) {
      double s = 0;
      for (int a = 0; a < 2; ++a) for (int bb = 0; bb < 2; ++bb) {
        double p[3] = {(double)i, (double)j, (double)k};
        p[d] += 0.5; p[(d + 1) % 3] += a; p[(d + 2) % 3] += bb;
        s += v[key[blk].at(K(p[0], p[1], p[2]))];
      }
      out[d] = s / (4.0 * h);
    }
  };
  auto face_vec = [&](const std::vector<double> &v, int blk, int i, int j, int k, double out[3], double &div) {   // RT at the centre
    div = 0;
    for (int d = 0; d < 3; ++d) {
      double p[3] = {i + 0.5, j + 0.5, k + 0.5};
      p[d] -= 0.5; const double lo = v[key[blk].at(K(p[0], p[1], p[2]))];
      p[d] += 1.0; const double hi = v[key[blk].at(K(p[0], p[1], p[2]))];
      out[d] = 0.5 * (lo + hi) / (h * h);
      div += (hi - lo) / (h * h * h);
    }
  };
  std::ofstream f(path);
  if (!f) throw std::runtime_error("cannot write " + path);
  f.precision(12);
  const int np = n1 * n1 * n1, nc = n * n * n;
  f << "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n"
    << "<Piece NumberOfPoints=\"" << np << "\" NumberOfCells=\"" << nc << "\">\n<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n";
  for (int k = 0; k < n1; ++k) for (int j = 0; j < n1; ++j) for (int i = 0; i < n1; ++i) f << c[0] + i * h << ' ' << c[1] + j * h << ' ' << c[2] + k * h << '\n';
  f << "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int32\" Name=\"connectivity\" format=\"ascii\">\n";
  auto P = [&](int i, int j, int k) { return i + n1 * (j + n1 * k); };
  for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i)
    f << P(i, j, k) << ' ' << P(i + 1, j, k) << ' ' << P(i + 1, j + 1, k) << ' ' << P(i, j + 1, k) << ' ' << P(i, j, k + 1) << ' '
      << P(i + 1, j, k + 1) << ' ' << P(i + 1, j + 1, k + 1) << ' ' << P(i, j + 1, k + 1) << '\n';
  f << "</DataArray>\n<DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">\n";
  for (int e = 1; e <= nc; ++e) f << 8 * e << '\n';
  f << "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n";
  for (int e = 0; e < nc; ++e) f << "12\n";
  f << "</DataArray>\n</Cells>\n";
  const bool nodal0 = pairing_ == MSFEC_Q || pairing_ == MSFEC_Q_NED;
  if (nodal0) {
    f << "<PointData Scalars=\"" << (pairing_ == MSFEC_Q ? "u" : "sigma") << "\">\n<DataArray type=\"Float64\" Name=\""
      << (pairing_ == MSFEC_Q ? "u" : "sigma") << "\" format=\"ascii\">\n";
    for (int k = 0; k < n1; ++k) for (int j = 0; j < n1; ++j) for (int i = 0; i < n1; ++i) f << b0[key[0].at(K(i, j, k))] << '\n';
    f << "</DataArray>\n</PointData>\n";
  }
  f << "<CellData>\n";
  auto vec_array = [&](const char *name, auto &&fn) {
    f << "<DataArray type=\"Float64\" Name=\"" << name << "\" NumberOfComponents=\"3\" format=\"ascii\">\n";
    for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) { double o[3]; fn(i, j, k, o); f << o[0] << ' ' << o[1] << ' ' << o[2] << '\n'; }
    f << "</DataArray>\n";
  };
  double dv;
  if (pairing_ == MSFEC_Q_NED) vec_array("u", [&](int i, int j, int k, double *o) { edge_vec(b1, 1, i, j, k, o); });
  if (pairing_ == MSFEC_NED_RT) {
    vec_array("sigma", [&](int i, int j, int k, double *o) { edge_vec(b0, 0, i, j, k, o); });
    vec_array("u", [&](int i, int j, int k, double *o) { face_vec(b1, 1, i, j, k, o, dv); });
    f << "<DataArray type=\"Float64\" Name=\"div_u\" format=\"ascii\">\n";
    for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) { double o[3]; face_vec(b1, 1, i, j, k, o, dv); f << dv << '\n'; }
    f << "</DataArray>\n";
  }
  if (pairing_ == MSFEC_RT_DQ) {
    vec_array("sigma", [&](int i, int j, int k, double *o) { face_vec(b0, 0, i, j, k, o, dv); });
    f << "<DataArray type=\"Float64\" Name=\"u\" format=\"ascii\">\n";
    for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) f << b1[key[1].at(K(i + 0.5, j + 0.5, k + 0.5))] << '\n';
    f << "</DataArray>\n";
  }
  f << "</CellData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n";
}

void morton_cell(int g, long long index, std::array<std::array<double, 3>, 8> &corners) {
  long long ijk[3] = {0, 0, 0};
  for (int b = 0; b < g; ++b)
    for (int d = 0; d < 3; ++d) ijk[d] |= ((index >> (3 * b + d)) & 1LL) << b;
  const double H = 1.0 / (double)(1LL << g);
  for (int v = 0; v < 8; ++v) {
    corners[v][0] = (ijk[0] + (v & 1)) * H;
    corners[v][1] = (ijk[1] + ((v >> 1) & 1)) * H;
    corners[v][2] = (ijk[2] + (v >> 2)) * H;
  }
}

void owned_range(long long n, int rank, int world, long long &lo, long long &hi) {
  lo = (rank * n) / world;
  hi = ((rank + 1) * n) / world;
}

std::string int_to_string(long long value, int digits) {
  std::string v = std::to_string(value);
  while ((int)v.size() < digits) v = "0" + v;
  return v;
}

std::string CellId::to_string() const {
  std::string s = "0_" + std::to_string(levels) + ":";
  for (int l = levels - 1; l >= 0; --l) s += char('0' + ((index >> (3 * l)) & 7));
  return s;
}

CoarseCell::CoarseCell(int global_refinements, long long morton_index) : id(morton_index, global_refinements) {
  morton_cell(global_refinements, morton_index, vertices);
}

// ---- per-cell classes ------------------------------------------------------------------------------------------
template <int PAIRING> const char *BasisT<PAIRING>::basis_file_stem() {
  return PAIRING == MSFEC_Q ? "basis_q" : PAIRING == MSFEC_Q_NED ? "basis_q-ned" : PAIRING == MSFEC_NED_RT ? "basis_ned-rt" : "basis_rt-dq";
}

// family tag and index inside the family (q_basis.cc:463-468, q_ned_basis.cc:991-1008, ned_rt_basis.cc:1001-1020,
// rt_dq_basis.cc:952-958)
template <int PAIRING> std::string BasisT<PAIRING>::basis_file_tag(int n_basis, int &idx) {
  idx = n_basis;
  if (PAIRING == MSFEC_Q) return "";
  if (PAIRING == MSFEC_Q_NED) { if (n_basis < 8) return ".h1"; idx = n_basis - 8; return ".curl"; }
  if (PAIRING == MSFEC_NED_RT) { if (n_basis < 12) return ".curl"; idx = n_basis - 12; return ".div"; }
  return ".div";
}

template <int PAIRING> void BasisT<PAIRING>::run() {
  char host[256] = "unknown";
  ::gethostname(host, sizeof(host) - 1);
  const bool talk = prm_->verbose_basis;
  // the reference prints this line for every cell (ned_rt_basis.cc:1284-1297, 1421-1427); here under "verbose basis",
  // the time being this cell's share of the batched device build
  if (talk)
    std::cout << "\tSolving for basis in cell   " << cell_.id.to_string() << "   [machine: " << host << " | rank: " << subdomain_
              << "]   .....";
  batch_->build();
  M_ = FullMatrix{batch_->k(), batch_->k(), batch_->matrix(local_)};
  r_ = Vector{batch_->k(), batch_->rhs(local_)};
  // set_filename_global (ned_rt_basis.cc:1254-1259)
  filename_global_ = prm_->filename_output + "." + int_to_string(subdomain_, 5) + ".cell-" + cell_.id.to_string() + ".vtu";
  // set_output_flag + output_basis (ned_rt_basis.cc:1404-1419): only the first cell writes its basis functions
  if (prm_->write_first_basis && cell_.id == first_cell_) {
    ::mkdir(prm_->dirname_output.c_str(), 0755);
    output_basis();
  }
  if (talk) std::cout << "done in   " << batch_->seconds_per_cell() << "   seconds." << std::endl;
}

template <int PAIRING> void BasisT<PAIRING>::output_basis() {
  std::vector<double> b0, b1;
  const int kb = msfec_k(PAIRING) - (PAIRING == MSFEC_RT_DQ ? 1 : 0);
  for (int n_basis = 0; n_basis < kb; ++n_basis) {
    int idx = 0;
    const std::string tag = basis_file_tag(n_basis, idx);
    const std::string filename = std::string(basis_file_stem()) + tag + "." + int_to_string(subdomain_, 5) + ".cell-" +
                                 cell_.id.to_string() + ".index-" + int_to_string(idx, 2) + ".vtu";
    batch_->basis_function(local_, n_basis, b0, b1);
    batch_->write_vtu(local_, prm_->dirname_output + "/" + filename, b0, b1);
  }
}

template <int PAIRING> void BasisT<PAIRING>::output_global_solution_in_cell() {
  std::vector<double> b0, b1;
  batch_->fine_solution(local_, b0, b1);
  batch_->write_vtu(local_, prm_->dirname_output + "/" + filename_global_, b0, b1);
}

template class BasisT<MSFEC_Q>;
template class BasisT<MSFEC_Q_NED>;
template class BasisT<MSFEC_NED_RT>;
template class BasisT<MSFEC_RT_DQ>;

}  // namespace msfec
