// Compiles and exercises the reference-shaped per-cell classes (basis.h) the way *Multiscale uses them
// (reference source/Ned_RT/ned_rt_global.cc:61-98, 245-252, 483-485): construct one XBasis per owned cell, COPY it into
// a std::map<CellId, XBasis> before run(), run() each, read get_global_element_matrix()/get_global_element_rhs(),
// set_global_weights(), and compare bit for bit with one batched msfec_build_basis call over the same cells.
// usage: test_basis_map <pairing 0..3> <file.prm>      (needs a B200; run by tests/test_gpu_host_driver.py)
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>

#include "basis.h"

using namespace msfec;

template <int PAIRING>
int check(const std::string &prm_file) {
  ParametersMs prm(prm_file, PAIRING);
  const int g = prm.n_refine_global;
  const long long n = 1LL << (3 * g);
  auto batch = std::make_shared<BasisBatch>(prm, 0);
  std::map<CellId, BasisT<PAIRING>> cell_basis_map;
  const CellId first_cell(0, g);
  for (long long id = 0; id < n; ++id) {
    const CoarseCell cell(g, id);
    BasisT<PAIRING> current_cell_problem(prm, prm_file, cell, first_cell, 0u, batch);
    BasisT<PAIRING> copy(current_cell_problem);                   // copyable before run()
    cell_basis_map.emplace(cell.id, copy);
  }
  bool threw = false;
  try { cell_basis_map.begin()->second.get_global_element_matrix(); } catch (const std::logic_error &) { threw = true; }
  if (!threw) { std::cerr << "getter before run() did not throw\n"; return 1; }
  for (auto &kv : cell_basis_map) kv.second.run();
  // the same cells through the C ABI in one call
  const int k = msfec_k(PAIRING);
  std::vector<double> corners((size_t)n * 24), M((size_t)n * k * k), r((size_t)n * k);
  std::vector<int64_t> ids(n);
  for (long long id = 0; id < n; ++id) {
    const CoarseCell cell(g, id);
    for (int v = 0; v < 8; ++v) for (int d = 0; d < 3; ++d) corners[(size_t)id * 24 + 3 * v + d] = cell.vertices[v][d];
    ids[id] = id;
  }
  msfec_ctx *ctx = nullptr;
  if (msfec_create(0, &prm.problem, &ctx)) { std::cerr << msfec_last_error(nullptr) << "\n"; return 1; }
  if (msfec_build_basis(ctx, (int)n, corners.data(), ids.data(), M.data(), r.data(), nullptr)) { std::cerr << msfec_last_error(ctx) << "\n"; return 1; }
  long long id = 0;
  for (auto &kv : cell_basis_map) {
    const FullMatrix &Mc = kv.second.get_global_element_matrix();
    const Vector &rc = kv.second.get_global_element_rhs();
    if (Mc.m() != k || Mc.n() != k || rc.size() != k) { std::cerr << "wrong sizes\n"; return 1; }
    if (std::memcmp(Mc.data, &M[(size_t)id * k * k], sizeof(double) * k * k) || std::memcmp(rc.data, &r[(size_t)id * k], sizeof(double) * k)) {
      std::cerr << "cell " << kv.first.to_string() << ": per-cell interface differs from the batched C ABI\n";
      return 1;
    }
    if (Mc(1, 2) != M[(size_t)id * k * k + k + 2]) { std::cerr << "operator() indexing\n"; return 1; }
    ++id;
  }
  // set_global_weights + reconstruction: weights e_0 reproduce basis function 0
  std::vector<double> w(k, 0.0), b0, b1, f0, f1;
  w[0] = 1.0;
  for (auto &kv : cell_basis_map) kv.second.set_global_weights(w);
  auto &last = cell_basis_map.rbegin()->second;
  last.get_global_solution(f0, f1);
  if (msfec_get_basis(ctx, (int)n - 1, 0, (b0.resize(f0.size()), b0.data()), f1.empty() ? nullptr : (b1.resize(f1.size()), b1.data()))) { std::cerr << msfec_last_error(ctx) << "\n"; return 1; }
  for (size_t i = 0; i < f0.size(); ++i) if (f0[i] != b0[i]) { std::cerr << "reconstruction differs from basis function 0\n"; return 1; }
  if (last.get_filename_global().find(".cell-" + last.id().to_string() + ".vtu") == std::string::npos) { std::cerr << "filename_global: " << last.get_filename_global() << "\n"; return 1; }
  msfec_destroy(ctx);
  std::printf("test_basis_map pairing %d: %lld cells, k = %d, per-cell interface == batched C ABI (bitwise); filename %s\n", PAIRING, n, k,
              last.get_filename_global().c_str());
  return 0;
}

int main(int argc, char **argv) {
  if (argc != 3) { std::cerr << "usage: test_basis_map <pairing 0..3> <file.prm>\n"; return 2; }
  try {
    switch (std::atoi(argv[1])) {
      case 0: return check<MSFEC_Q>(argv[2]);
      case 1: return check<MSFEC_Q_NED>(argv[2]);
      case 2: return check<MSFEC_NED_RT>(argv[2]);
      default: return check<MSFEC_RT_DQ>(argv[2]);
    }
  } catch (std::exception &e) { std::cerr << "test_basis_map: " << e.what() << "\n"; return 1; }
}
