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
}  // namespace

ParametersMs::ParametersMs(const std::string &prm_filename, int pairing) {
  if (msfec_problem_from_prm(prm_filename.c_str(), pairing, &problem))
    throw std::runtime_error(std::string("parameter file: ") + msfec_last_error(nullptr));
  filename_output = ms_top_level_value(prm_filename, "filename output", filename_output);   // ned_rt_parameters.cc:226-236
  dirname_output = ms_top_level_value(prm_filename, "dirname output", dirname_output);
  write_first_basis = ms_control_flag(prm_filename, "write first basis", false);
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

namespace {
int env_int(const char *a, const char *b, const char *c, int dflt) {
  for (const char *n : {a, b, c})
    if (n) if (const char *v = std::getenv(n)) return std::atoi(v);
  return dflt;
}
}  // namespace

// Mirrors source/main_ned_rt.cxx:15-117: parse "-p <prm>", run, catch-all.  The fine-grid `*Std` solve and the
// global coarse solve are outside the hot path; this driver performs the basis build of the
// rank's cells, reports the "basis initialization and computation" time and writes the coarse element matrices.
int driver_main(int argc, char **argv, int pairing, const char *name) {
  try {
    std::string prm_file;
    for (int i = 1; i < argc; ++i) {
      const std::string a = argv[i];
      if (a == "-p" && i + 1 < argc) prm_file = argv[++i];
      else if (a == "-h" || a == "--help") { std::cout << "usage: MsFEC_" << name << " -p parameter_file.prm\n"; return 0; }
      else { std::cerr << "Unknown command line option: " << a << "\nusage: MsFEC_" << name << " -p parameter_file.prm\n"; return 1; }
    }
    if (prm_file.empty()) { std::cerr << "usage: MsFEC_" << name << " -p parameter_file.prm\n"; return 1; }
    const int rank = env_int("OMPI_COMM_WORLD_RANK", "PMI_RANK", "RANK", 0);
    const int world = env_int("OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "WORLD_SIZE", 1);
    const int device = env_int("OMPI_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", "LOCAL_RANK", 0);
    ParametersMs prm(prm_file, pairing);
    const long long n_cells = 1LL << (3 * prm.n_refine_global);
    long long lo, hi;
    owned_range(n_cells, rank, world, lo, hi);
    if (rank == 0)
      std::cout << "MsFEC_" << name << ": " << n_cells << " coarse cells (global refinements " << prm.n_refine_global
                << "), " << prm.n_refine_local << " local refinements, " << world << " rank(s)\n";
    auto batch = std::make_shared<BasisBatch>(prm, device);
    std::vector<int> locals;
    std::array<std::array<double, 3>, 8> c;
    for (long long id = lo; id < hi; ++id) { morton_cell(prm.n_refine_global, id, c); locals.push_back(batch->add_cell(c, id)); }
    const auto t0 = std::chrono::steady_clock::now();
    batch->build();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const msfec_stats &st = batch->stats();
    std::cout << "[rank " << rank << "] " << name << " basis initialization and computation: " << (hi - lo) << " cells in " << sec
              << " s (device " << st.ms_total << " ms; assemble " << st.ms_assemble << ", lift " << st.ms_lift << ", solve "
              << st.ms_solve << ", coarse matrices " << st.ms_gram << "; solver " << (st.solver ? "direct" : "MINRES")
              << ", max its " << st.iterations_max << ", " << st.kernel_launches << " kernel launches)\n";
    ::mkdir(prm.dirname_output.c_str(), 0755);
    // "write first basis" (ned_rt_basis.cc:1404-1419): the basis functions of the first coarse cell as VTU
    if (prm.write_first_basis && rank == 0 && !locals.empty()) {
      std::vector<double> b0, b1;
      const int kb = msfec_k(pairing) - (pairing == MSFEC_RT_DQ ? 1 : 0);
      for (int i = 0; i < kb; ++i) {
        batch->basis_function(locals[0], i, b0, b1);
        batch->write_vtu(locals[0], prm.dirname_output + "/basis_" + name + ".cell-" + std::to_string(lo) + ".index-" + std::to_string(i) + ".vtu", b0, b1);
      }
      std::cout << "[rank 0] wrote " << kb << " basis VTU files of coarse cell " << lo << " to " << prm.dirname_output << "\n";
    }
    const std::string out = prm.dirname_output + "/" + std::string(name) + "_element_matrices.rank" + std::to_string(rank) + ".bin";
    std::ofstream f(out, std::ios::binary);
    const int k = batch->k();
    const long long hdr[4] = {hi - lo, k, lo, pairing};
    f.write((const char *)hdr, sizeof(hdr));
    double checksum = 0;
    for (int i : locals) {
      f.write((const char *)batch->matrix(i), sizeof(double) * k * k);
      f.write((const char *)batch->rhs(i), sizeof(double) * k);
      for (int q = 0; q < k * k; ++q) checksum += batch->matrix(i)[q];
    }
    std::printf("[rank %d] wrote %s ; sum of all matrix entries = %.15e\n", rank, out.c_str(), checksum);
    return 0;
  } catch (std::exception &exc) {
    std::cerr << "\n----------------------------------------------------\nException on processing:\n" << exc.what()
              << "\nAborting!\n----------------------------------------------------\n";
    return 1;
  } catch (...) {
    std::cerr << "\nUnknown exception!\nAborting!\n";
    return 1;
  }
}

}  // namespace msfec
