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
}  // namespace

ParametersMs::ParametersMs(const std::string &prm_filename, int pairing) {
  if (msfec_problem_from_prm(prm_filename.c_str(), pairing, &problem))
    throw std::runtime_error(std::string("parameter file: ") + msfec_last_error(nullptr));
  filename_output = ms_top_level_value(prm_filename, "filename output", filename_output);   // ned_rt_parameters.cc:226-236
  dirname_output = ms_top_level_value(prm_filename, "dirname output", dirname_output);
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
