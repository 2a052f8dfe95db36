// CPU-only check of the host coarse solver (coarse.h): reads coarse element matrices from a file, assembles and solves
// the global coarse system, writes the per-cell weights.  Driven by tests/test_host_driver.py, which compares with the
// harness' scipy solve (tests/coarse_solve.py).
//   input  (binary): int64 {pairing, global refinements, n_cells, dense_limit}, then per cell: int64 id, double M[k*k], double r[k]
//   output (binary): double weights[n_cells][k] in input order
// optional third argument: CUDA device for the iterative solve (include/msfec_coarse.h); default: host
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "../../include/msfec.h"
#include "coarse.h"

int main(int argc, char **argv) {
  if (argc != 3 && argc != 4) { std::cerr << "usage: coarse_test <in.bin> <out.bin> [device]\n"; return 2; }
  try {
    std::ifstream in(argv[1], std::ios::binary);
    long long hdr[4];
    in.read((char *)hdr, sizeof(hdr));
    msfec::CoarseProblem cp((int)hdr[0], (int)hdr[1]);
    cp.set_dense_limit((int)hdr[3]);
    if (argc == 4) cp.set_device(std::atoi(argv[3]));
    const int k = cp.k();
    std::vector<long long> ids(hdr[2]);
    std::vector<double> M((size_t)k * k), r(k);
    for (long long c = 0; c < hdr[2]; ++c) {
      in.read((char *)&ids[c], sizeof(long long));
      in.read((char *)M.data(), sizeof(double) * k * k);
      in.read((char *)r.data(), sizeof(double) * k);
      if (!in) throw std::runtime_error("short input file");
      cp.add_cell(ids[c], M.data(), r.data());
    }
    std::cout << cp.solve() << std::endl;
    std::ofstream out(argv[2], std::ios::binary);
    std::vector<double> w(k);
    for (long long c = 0; c < hdr[2]; ++c) { cp.cell_weights(ids[c], w.data()); out.write((const char *)w.data(), sizeof(double) * k); }
    return 0;
  } catch (std::exception &e) { std::cerr << "coarse_test: " << e.what() << "\n"; return 1; }
}
