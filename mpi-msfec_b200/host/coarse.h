// Global coarse problem of the multiscale method on the host: what *Multiscale::setup_system_matrix /
// assemble_system / solve_iterative / send_global_weights_to_cell do with deal.II + Trilinos in the reference
// (source/Ned_RT/ned_rt_global.cc:100-488; q_global.cc:89-363; q_ned_global.cc:101-470; rt_dq_global.cc:96-470).
//
// The coarse mesh is the unit cube refined `global refinements` times (hyper_cube + refine_global,
// ned_rt_global.cc:36-46): m = 2^g cells per direction, lowest-order elements, so DoFs are the vertices / edges /
// faces / cells of a structured grid.  Local numbering of a coarse cell follows deal.II: vertex
// v = a + 2b + 4c; lines 0-3 bottom {x=0 ||y, x=1 ||y, y=0 ||x, y=1 ||x}, 4-7 top, 8-11 vertical; faces x0 x1 y0 y1
// z0 z1 -- the order of the k x k element matrices the basis build returns.
//
// Essential data (all homogeneous in the shipped runs): Q -- boundary vertices (q_global.cc:127-139); Q_Ned -- boundary
// vertices and boundary edges (q_ned_global.cc:175-207); Ned_RT / RT_DQ -- none (natural conditions,
// ned_rt_global.cc:162-190, rt_dq_global.cc:154-182).
//
// Solver: Trilinos is not available here.  Dense LU with partial pivoting up to a few thousand unknowns (exact: the parity
// tests), beyond that the reference's own scheme -- Schur-complement CG with an inner CG on block (0,0)
// (ned_rt_global.cc:330-461), Q: plain CG -- with tolerances tightened from the reference's 1e-6 so that the final
// solution is reproducible to 1e-8.  The iteration runs on the rank's GPU when a device was set (set_device; the
// drivers do: include/msfec_coarse.h, csrc/coarse_dev.cu -- C5's 205 920 unknowns: about a second instead of a minute on
// one host core) and on the host otherwise (coarse_test, CPU tests).
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace msfec {

class CoarseProblem {
 public:
  CoarseProblem(int pairing, int global_refinements);
  int k() const { return k_; }
  long long n_cells() const { return (long long)m_ * m_ * m_; }
  int n_dofs() const { return n0_ + n1_; }
  int n_block0() const { return n0_; }
  int n_block1() const { return n1_; }
  // global coarse DoF indices of the cell with Morton (z-order) index `cell`, block 0 first
  void cell_dofs(long long cell, int32_t *dofs) const;
  bool is_essential(int dof) const { return essential_[dof] != 0; }
  // add one cell's element matrix (k x k row-major) and rhs (k)
  void add_cell(long long cell, const double *M, const double *r);
  // solve; returns a short description of what ran (solver, iterations, residual)
  std::string solve();
  const std::vector<double> &solution() const { return x_; }
  // the k weights of one cell (send_global_weights_to_cell, ned_rt_global.cc:464-488)
  void cell_weights(long long cell, double *w) const;
  double residual() const { return residual_; }   // ||b - A x||_2 / ||b||_2 over the free unknowns
  void set_dense_limit(int n) { dense_limit_ = n; }
  void set_device(int device) { device_ = device; }   // >= 0: the iterative solve runs on that GPU

 private:
  void ijk_of(long long cell, int &i, int &j, int &k) const;
  void finalize();
  int pairing_, g_, m_, k_;
  int kind0_, kind1_;            // entity kind of each block: 0 vertex, 1 edge, 2 face, 3 cell, -1 none
  int n0_ = 0, n1_ = 0;
  std::vector<uint8_t> essential_;
  // triplets, then CSR
  std::vector<int32_t> ti_, tj_;
  std::vector<double> tv_;
  std::vector<int32_t> ptr_, col_;
  std::vector<double> val_, b_, x_;
  bool finalized_ = false;
  double residual_ = 0;
  int dense_limit_ = 6000;
  int device_ = -1;
};

}  // namespace msfec
