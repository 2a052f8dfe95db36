// Plan of the batched MULTIFRONTAL direct solver (the fast path of "use direct solver basis", reference
// solve_direct, source/Ned_RT/ned_rt_basis.cc:579-634, and -- by automatic selection -- of the iterative branch
// solve_iterative :637-847 whose tolerance the exact factorisation satisfies).
//
// The interior saddle system of one coarse cell ([A00 -K^T; -K -A11], Topology::sys) is ordered by geometric
// nested dissection down to boxes of 2 x 2 x 2 fine cells.  Every box interior / separator piece is a SUPERNODE;
// supernode f owns s_f unknowns (sigma-type first) and reaches u_f unknowns of its ancestors.  Its FRONT is the dense
// symmetric matrix on [own | reached | rhs rows]; the k right-hand sides ride along as extra rows, so the forward
// substitution is part of the factorisation (same trick as the band solver, DirectPlan).  Classic multifrontal
// recurrence, one front per CTA, fronts of one tree level per launch:
//     panel  P = [own columns of the front]            (m x s8, shared memory)
//            := original matrix entries + rhs + first n_own columns of the children's contribution blocks
//     factor P:  L11 D L11^T = P11,  X = P21 L11^-T,  L21 = X D^-1
//     C (u8 + kr) x u8  :=  sum_children C_child (extend-add through index maps)  -  X L21^T     -> global memory
// and top-down   x_own = L11^-T (z_own - L21^T x_reached).
// No pivoting: every leading set of unknowns is a union of sub-box problems with essential conditions on the
// cuts, pivots are positive on sigma-type and negative on u-type unknowns; the CPU test-suite
// replays the plan in numpy (tests/emulate.py: emulate_multifrontal) for every pairing.
//
// At n = 8 (Ned_RT, C5) the exact elimination costs 34 MFLOP per cell instead of the 420 MFLOP the 32-padded
// layer/plane band executes, and every front fits one SM's shared memory, so the build becomes HBM-bound on the
// contribution blocks and the factor.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "topology.h"

namespace msfec {

// one row of MfPlan::fronts (all int32 so that the table can be exported / uploaded verbatim)
struct MfFront {
  int32_t s, s8;          // own unknowns, padded to a multiple of 8 (identity pivots)
  int32_t u, u8;          // reached unknowns of ancestors, padded to a multiple of 8
  int32_t m;              // rows of the panel: s8 + u8 + kr
  int32_t ldx;            // shared-memory row stride of the panel (= 4 or 12 mod 16: conflict-free DMMA fragment loads)
  int32_t level;          // leaves 0; a front is one level above its highest child
  int32_t parent;         // -1: root
  int32_t own_base;       // padded position (MfPlan numbering) of own column 0
  int32_t idx_off;        // MfPlan::front_idx[idx_off .. + s8 + u8): padded position of every front row (-1: padding)
  int32_t row_off;        // MfPlan::own_rows[row_off .. + s8): interior row of every own column (-1: padding / pinned)
  int32_t pe_lo, pe_hi;   // per-cell matrix entries   panel[pe_dest] = +-vals[pe_ref >> 1]
  int32_t ps_lo, ps_hi;   // cell-independent entries  panel[ps_dest] = ps_val * kscale
  int32_t pc_lo, pc_hi;   // constants (padding pivots 1, pinned pivot -1)
  int32_t ch_lo, ch_hi;   // MfPlan::children[ch_lo .. ch_hi)
  int32_t l_off;          // per-cell factor storage: the front's record [panel m x ldx | 1/d | d | pivot-tile factors], doubles
  int32_t c_off;          // per-cell contribution storage: column-major (u8 + kr) x u8, doubles
  int32_t n_rows_real;    // s + u (diagnostics)
  int32_t ldc_max;        // largest column length (u8 + kr) of the children's contribution blocks (0: leaf)
};
constexpr int kMfFrontFields = sizeof(MfFront) / sizeof(int32_t);

struct MfChild {
  int32_t front;          // child front
  int32_t n_own;          // the first n_own columns of the child's C fall into the parent's own columns
  int32_t cmap_off;       // MfPlan::cmap[cmap_off .. + u8_c + kr): child row -> parent front row (-1: padding)
  int32_t pinv_off;       // MfPlan::pinv[pinv_off .. + m_parent): parent front row -> child row (-1: absent)
};

struct MfPlan {
  bool feasible = false;
  std::string why;                 // reason when not feasible
  int kr = 0;                      // rhs rows per front: k_solve rounded up to a multiple of 8
  int NP = 0;                      // padded number of unknowns (sum of s8, rounded up to a multiple of 32)
  int n_levels = 0;
  std::vector<MfFront> fronts;     // elimination order (children before parents)
  std::vector<int32_t> level_off, level_fronts;   // fronts grouped by level: level_fronts[level_off[l] .. level_off[l+1])
  std::vector<MfChild> children;
  std::vector<int32_t> front_idx, own_rows, cmap, pinv;
  std::vector<int32_t> pe_dest, pe_ref, ps_dest, pc_dest;
  std::vector<double> ps_val, pc_val;
  std::vector<int32_t> perm, inv_perm;   // interior row -> padded position; padded position -> interior row (-1)
  int64_t l_doubles = 0, c_doubles = 0;  // per-cell storage
  int pinned_row = -1;                   // RT_DQ: u DoF fixed to 0 (constant null space)
  double flops = 0;                      // FP64 flops per cell (factor + solves + contribution updates, padded sizes)
  double bytes_fwd = 0, bytes_bwd = 0;   // algorithmic HBM bytes per cell of k_mf_forward / k_mf_backward
  std::vector<int32_t> smem_fwd;         // per level: dynamic shared memory of the forward kernel (bytes)
  std::vector<int32_t> smem_bwd;         // per level: ... of the backward kernel
  std::vector<int32_t> smem_fwd_st;      // per level: ... of the forward kernel with staged children (0: level has leaves only)
  std::vector<int32_t> rt_max;           // per level: most row tiles (u8 + kr) / 8 of a contribution block
};

// smem_budget: bytes of dynamic shared memory one CTA may use (227 KB on sm_100).  min_cells: edge (fine cells) of the
// boxes that are not dissected further.
MfPlan build_mf_plan(const Topology &t, int smem_budget = 232448 - 1024, int min_cells = 2);

// shared-memory layout of the kernels (bytes), used by both the plan (feasibility) and the launches
int mf_ldx(int s8);
size_t mf_smem_forward(const MfFront &f, int n_children, int kr);
size_t mf_smem_forward_staged(const MfFront &f, int n_children);   // children's columns streamed through a ring (mf.cuh)
constexpr int kMfStageBufs = 2;   // ring depth of the staged forward kernel
size_t mf_smem_backward(const MfFront &f, int kr);

}  // namespace msfec
