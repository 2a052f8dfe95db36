// Shared fine-grid topology, sparsity patterns and assembly tables.
//
// Every coarse cell of the reference is refined identically
// (GridGenerator::general_cell + refine_global, reference
// source/Ned_RT/ned_rt_basis.cc:167-179) and gets the same FESystem, DoF enumeration,
// sparsity pattern and boundary-constraint structure (:181-359).  The reference redoes
// that setup for every cell; here it is done ONCE per (pairing, n) on the host and all
// coarse cells share it.  Only matrix *values* (the "slots") differ between cells.
//
// Numbering: each block (sigma-type = block 0, u-type = block 1) numbers its interior
// DoFs first, then its boundary (essential-BC) DoFs.  Fine DoF numbering never leaves
// the reference's basis object, so this choice is free.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace msfec {

enum EntityKind { ENT_V = 0, ENT_E = 1, ENT_F = 2, ENT_C = 3 };

struct BlockTopo {
  int kind = 0;
  int ldofs = 0;                 // local DoFs per fine cell: 8 / 12 / 6 / 1
  int n_total = 0, n_int = 0;
  std::vector<int32_t> cell_dofs;   // [nC][ldofs] -> index in interior-first numbering
  std::vector<double> pos;          // [n_total][3], units of h (edge midpoints, face centres)
  std::vector<int32_t> axis;        // [n_total] direction of edge / face normal, -1 otherwise
  std::vector<uint8_t> bnd;         // [n_total]
};

// value = scale * sum_{contrib of slot} sum_{e in pair list} w_e * coef[fine cell][idx_e]
struct AsmTable {
  int n_slots = 0;
  int coef_stride = 0;                  // coefficient channels per fine cell (8*7 or 8*ncomp)
  std::vector<int32_t> contrib_ptr;     // [n_slots+1]
  std::vector<int32_t> contrib_cell;    // fine cell
  std::vector<int32_t> contrib_pair;    // index into pair_ptr
  std::vector<int32_t> pair_ptr;        // [n_pairs+1]
  std::vector<int32_t> pair_idx;        // coefficient channel (q*ncomp + comp)
  std::vector<double> pair_w;
  int h_exponent = 0;                   // scale = h^h_exponent / 8
};

// CSR operator whose entries reference per-cell slots (cell part) or cell-independent
// values (shared part, to be multiplied by kscale = h^k_h_exponent).
struct RefOperator {
  int n_rows = 0, n_cols = 0;
  std::vector<int32_t> cptr, ccol, cref;   // cref = (slot << 1) | negate
  std::vector<int32_t> sptr, scol;
  std::vector<double> sval;
};

struct Topology {
  int pairing = 0, n = 0, nC = 0;
  int k_solve = 0, k_gram = 0, k0 = 0;     // k0 = number of sigma-type coarse functions
  bool two_blocks = false;
  BlockTopo blk[2];
  int NI = 0, NB = 0, NF = 0;              // interior / boundary / full sizes over both blocks
  int n_slots0 = 0, n_slots1 = 0;          // A00 slots, A11 slots (slot space = [A00 | A11])
  AsmTable asm00, asm11, asm_rhs;          // asm_rhs: one "slot" per u-type DoF (all DoFs for Q)
  int coef_tensor_is_inverse = 0;          // tensor channel holds A^{-1} (else A)
  int coef_scalar_is_inverse = 0;          // scalar channel holds 1/B (else B)
  int rhs_ncomp = 1;
  int k_h_exponent = 0;                    // K ~ h^p
  int f1_H_exponent = 0;                   // F1 (unit-H table) scales with H^p
  RefOperator sys;                         // interior saddle system, symmetric form [A00 -K^T; -K -A11]
  RefOperator lift;                        // rows = interior, cols = boundary DoFs [B0 | B1]
  RefOperator full;                        // rows/cols = all DoFs [all0 | all1], form [A00 -K^T; K A11]
  RefOperator kint;                        // K restricted to interior rows (blk1) x interior cols (blk0)
  std::vector<int32_t> diag_slot0, diag_slot1;   // slot of (r,r) for interior rows of each block (-1: none)
  std::vector<double> G;                   // [k_solve][NB] essential boundary data (H-independent)
  std::vector<double> F1;                  // [k_solve][NI] volume rhs of the symmetric-form system, unit H
  int rhs_block = 0;                       // block the global rhs lives on
};

// ---------------------------------------------------------------------------------------
// Plan of the batched direct solver ("use direct solver basis = true", reference
// solve_direct, ned_rt_basis.cc:579-634).  The interior saddle system is symmetric in the
// form [A00 -K^T; -K -A11].  Interior DoFs are grouped into BLOCKS along z, eliminated in order:
//   "split" (default): layer 0, plane 1, layer 1, plane 2, ...  (DoFs inside a fine-cell layer /
//                      DoFs lying on the plane between two layers)
//   "slab":            layer k together with the plane above it
// sigma-type DoFs first inside a block, padded to a multiple of 32 with identity pivots.  Every
// leading set of blocks is a sub-box problem with essential conditions on the cut, so an LDL^T
// factorisation WITHOUT pivoting exists (positive pivots on sigma-type, negative on u-type DoFs;
// DESIGN.md).  A symbolic factorisation on the block graph gives, for every block column s, the
// list of later blocks it reaches (its "front": a layer reaches the plane above it; a plane reaches
// the next layer and the next plane).  Storage per cell ("band"): block column s is a column-major
// (ld_s x bs_s) panel whose rows are [block s | reached blocks ... | 32 right-hand-side rows]; the
// k right-hand sides ride along as extra rows so the forward substitution is part of the
// factorisation.
struct DirectPlan {
  static constexpr int kPanel = 32;     // panel width == padding granularity
  static constexpr int kRhsRows = 32;   // rhs rows per block column (k <= 20 used)
  int n_slabs = 0, NP = 0;              // number of blocks; NP = padded number of unknowns
  std::vector<int32_t> bs, slab_off, ld;      // per block (bs/ld padded), slab_off in padded numbering
  std::vector<int32_t> front_rows;            // ld - kRhsRows: DoF rows of the front of block column s
  std::vector<int64_t> col_off;               // offset (doubles) of block column s inside a cell's band
  int64_t band_doubles = 0;
  // front structure in 32-row chunks: chunk t of block column s covers front rows [32t, 32t+32)
  std::vector<int32_t> chunk_off;             // [n_slabs+1]
  std::vector<int32_t> chunk_blk;             // block the chunk belongs to
  std::vector<int32_t> chunk_local;           // row offset of the chunk inside that block
  std::vector<int32_t> front_pos;             // [n_slabs*n_slabs] front row offset of block b in column c, -1 if absent
  std::vector<int32_t> perm;                  // interior row (stacked numbering) -> padded index
  std::vector<int32_t> inv_perm;              // padded index -> interior row or -1 (padding / pinned)
  std::vector<int32_t> cell_dest, cell_ref;   // fill list, per-cell slot entries (lower triangle)
  std::vector<int32_t> shared_dest;           // fill list, cell-independent entries (x kscale)
  std::vector<double> shared_val;
  std::vector<int32_t> const_dest;            // padding / pinned pivots
  std::vector<double> const_val;
  std::vector<int32_t> rhs_dest;              // [NI] band offset of rhs row 0 of this DoF's column (-1: pinned)
  int pinned_row = -1;                        // RT_DQ: u DoF fixed to 0 (constant null space)
  double update_flops = 0;                    // FP64 flops of all trailing updates per cell (lower triangle)
};

// Gram matrix of a fine-grid norm over the DoFs of ONE block (interior-first numbering), CSR, identical for
// every coarse cell up to the factor h^h_exponent (values include the 1/8 quadrature weight of the unit cube).
// The reference computes no error norm (integrate_difference never appears); these define the harness norms of
// the north star: L2 and the natural semi-norm of each field with unit coefficients, Gauss 2x2x2 on every fine cell.
struct NormOperator {
  int n = 0, h_exponent = 0;
  std::vector<int32_t> ptr, col;
  std::vector<double> val;
};
// [2 b + 0]: L2 Gram of block b;  [2 b + 1]: |.|_H1 (vertices), |.|_H(curl) (edges), |.|_H(div) (faces), empty for
// cell-wise constants.  Entries of a missing block are empty (n = 0).
std::vector<NormOperator> build_norm_operators(const Topology &t);

// Builds everything above.  pairing: enum msfec_pairing; n = 2^L.
Topology build_topology(int pairing, int n);
// ordering: 0 = split layers/planes (default), 1 = slabs, 2 = geometric nested dissection (chosen automatically from
// n = 16 on when its symbolic update flop count is lower; MSFEC_DIRECT_ORDERING = split | slab | nd overrides)
DirectPlan build_direct_plan(const Topology &t, int ordering = 0);

// Quadrature abscissae of QGauss<3>(2) on the unit cube, x fastest.
void gauss_points(double qp[8][3]);

}  // namespace msfec
