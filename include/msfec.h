/*
 * msfec.h -- C ABI of the B200-native MsFEC multiscale basis build.
 *
 * The reference (konsim83/MPI-MSFEC) has no FFI layer; the seam it offers is the C++
 * interface of its four per-coarse-cell classes QBasis / QNedBasis / NedRTBasis /
 * RTDQBasis (reference: include/Ned_RT/ned_rt_basis.h:99-153 and siblings), driven by
 * *Multiscale::initialize_and_compute_basis (source/Ned_RT/ned_rt_global.cc:49-99).
 * This header is the batched, exception-free C equivalent of that seam: one context
 * per GPU/rank, all locally owned coarse cells in one call.  A thin C++ adaptor with
 * the reference's per-cell method names sits on top (mpi-msfec_b200/host/basis.h).
 *
 * Conventions: every function returns 0 on success and a non-zero MSFEC_E* code on
 * failure; msfec_last_error() then describes it.  The caller owns all host buffers.
 * No C++ exception crosses this boundary.  Calls on one context are serialised by the
 * caller; different contexts are independent.  There is NO CPU fallback: compute
 * entry points fail with MSFEC_ENODEVICE when no sm_100 device is usable.
 */
#ifndef MSFEC_H_
#define MSFEC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSFEC_ABI_VERSION 2

enum msfec_pairing {
  MSFEC_Q      = 0, /* FE_Q(1)                      k = 8   (q_basis.cc:16)        */
  MSFEC_Q_NED  = 1, /* FE_Q(1) + FE_Nedelec(0)      k = 8+12 (q_ned_basis.cc:18)   */
  MSFEC_NED_RT = 2, /* FE_Nedelec(0)+FE_RT(0)       k = 12+6 (ned_rt_basis.cc:18)  */
  MSFEC_RT_DQ  = 3  /* FE_RT(0) + FE_DGQ(0)         k = 6+1  (rt_dq_basis.cc:18)   */
};

enum msfec_error {
  MSFEC_OK = 0,
  MSFEC_EINVAL = 1,    /* bad argument / malformed problem description            */
  MSFEC_ENODEVICE = 2, /* no usable CUDA device (there is no CPU fallback)        */
  MSFEC_ECUDA = 3,     /* CUDA runtime error                                      */
  MSFEC_ENOMEM = 4,
  MSFEC_ENOTCONVERGED = 5, /* Krylov solve hit the iteration cap (results still written) */
  MSFEC_ESTATE = 6,    /* call order violated (e.g. set_weights before build)     */
  MSFEC_EPARSE = 7     /* .prm or expression syntax error                         */
};

/* Coefficient / control description.  Mirrors what the reference reads from the .prm
 * in ParametersMs / ParametersBasis (source/Ned_RT/ned_rt_parameters.cc:160-270),
 * Diffusion_A_Data (source/equation_data/eqn_coeff_A.cc:58-140), Diffusion_B_Data
 * (eqn_coeff_B.cc:22-70) and RightHandSideParsed (eqn_rhs.cc:58-88). */
typedef struct msfec_problem {
  int32_t pairing;               /* enum msfec_pairing                               */
  int32_t n_refine_local;        /* "local refinements": n = 2^L fine cells per axis, L in [0, 6].  L = 0: the cell is its own
                                    fine grid, nothing is solved and the coarse matrices are the STANDARD lowest-order
                                    element matrices (the fine-grid comparator of the host driver, ned_rt_ref.cc)        */
  int32_t n_refine_global;       /* "global refinements" (host driver only)          */
  int32_t use_direct_solver_basis; /* "use direct solver basis"                      */
  int32_t verbose_basis;         /* "verbose basis"                                  */
  int32_t a_rotate;              /* Diffusion A / rotate                             */
  int32_t a_freq[3];             /* Diffusion A / frequency {x,y,z}                  */
  int32_t b_freq;                /* Diffusion B / frequency                          */
  double a_scale[3];             /* Diffusion A / scale {x,y,z}                      */
  double a_alpha[3];             /* Diffusion A / alpha {x,y,z}                      */
  double b_scale, b_alpha;       /* Diffusion B / scale, alpha                       */
  const char *b_expression;      /* Diffusion B / Function expression                */
  const char *rhs_expression;    /* Right-hand side / Function expression (';' sep.) */
  const char *rhs_constants;     /* Right-hand side / Function constants "a=1,b=2"   */
  /* Harness-defined rough random field.  seed == 0
   * selects the reference's sine family above. */
  uint64_t random_field_seed;
  double random_field_sigma;
  /* Krylov controls (the reference hard-codes 1e-6 / 1e-12; parity against the exact
   * discrete solution needs tighter values).  <= 0 selects defaults. */
  double krylov_rtol;
  int32_t krylov_max_iter;
  int32_t cells_per_batch;       /* coarse cells resident on the device at once; <= 0: derived from free memory */
  /* Solver of the local problems (enum msfec_solver).  MSFEC_SOLVER_AUTO follows the .prm: "use direct solver
   * basis = true" asks for the exact solve; "= false" (what every shipped .prm sets) asks for the reference's
   * iterative tolerance (1e-6), which the exact factorisation satisfies -- so AUTO picks the fastest exact solver that
   * fits (multifrontal up to 3 local refinements, banded block LDL^T beyond) and falls back to batched MINRES only when
   * no factorisation plan exists for the problem size. */
  int32_t solver;
  int32_t reserved0;
} msfec_problem;

enum msfec_solver {
  MSFEC_SOLVER_AUTO = 0,
  MSFEC_SOLVER_MINRES = 1,       /* batched preconditioned MINRES (solve_iterative, ned_rt_basis.cc:637-847)        */
  MSFEC_SOLVER_BAND = 2,         /* batched block LDL^T on a layer/plane or dissected band (solve_direct, :579-634)  */
  MSFEC_SOLVER_MULTIFRONTAL = 3  /* batched multifrontal LDL^T, fronts in shared memory (solve_direct, :579-634)    */
};

typedef struct msfec_stats {
  int32_t n_cells;
  int32_t k;                     /* coarse DoFs per cell (8, 20, 18, 7)              */
  int32_t n_fine_dofs;           /* fine DoFs per cell, both blocks                  */
  int32_t n_fine_dofs_interior;
  int32_t iterations_max;        /* max Krylov iterations over all cells/rhs         */
  int32_t not_converged;         /* number of (cell, rhs) columns over the cap       */
  int32_t kernel_launches;       /* kernels launched by the last build               */
  int32_t reserved;
  double iterations_mean;
  double residual_max;           /* MINRES: max final relative preconditioned residual;
                                  * direct: max over cells of the TRUE relative residual
                                  * of a fixed generic combination of the k rhs        */
  double ms_assemble, ms_lift, ms_solve, ms_gram, ms_total;   /* CUDA-event times    */
  double krylov_matrix_bytes;    /* algorithmic bytes of the Krylov kernel*/
  double krylov_ms_spmm;         /* device time inside the SpMM kernel               */
  int64_t krylov_spmm_launches;
  /* direct path ("use direct solver basis = true") */
  int64_t direct_update_launches;/* k_direct_update_s launches of the build (one per chunk of 32-column panels)     */
  double direct_flops;           /* FP64 tensor-core flops of the build: trailing updates (lower triangle) + solves */
  double direct_flops_timed;     /* flops of the k_direct_update_s launches that were bracketed by events           */
  double direct_ms_update;       /* summed device time of those launches                                            */
  int32_t solver;                /* 0 = batched MINRES, 1 = batched block LDL^T (band), 2 = batched multifrontal LDL^T,
                                    3 = none (0 local refinements: no interior unknowns)                              */
  int32_t direct_timed_launches; /* number of event-bracketed k_direct_update_s launches                             */
  /* multifrontal path */
  double mf_flops;               /* FP64 flops of the front kernels of the build (padded front sizes)                */
  double mf_bytes_fwd;           /* algorithmic HBM bytes of k_mf_forward: slot values + rhs read, every factor panel
                                  * written once, every contribution block written once and read once    */
  double mf_bytes_bwd;           /* ... of k_mf_backward: factor panels read once, solution gathered / written        */
  double mf_ms_fwd, mf_ms_bwd;   /* summed device time of the k_mf_forward / k_mf_backward launches (CUDA events)     */
  int64_t mf_launches;           /* k_mf_forward + k_mf_backward launches of the build                               */
} msfec_stats;

typedef struct msfec_ctx msfec_ctx;

/* Fill *p with the reference's declared defaults (the "declare_entry" defaults of the
 * three ParameterHandler blocks cited above). */
void msfec_problem_defaults(msfec_problem *p, int pairing);

/* Parse a reference .prm file verbatim (replaces ParametersMs::ParametersMs,
 * ned_rt_parameters.cc:131-158, plus the per-coefficient re-parses).  Strings are
 * owned by an internal arena released by msfec_problem_free(). */
int msfec_problem_from_prm(const char *prm_path, int pairing, msfec_problem *p);
void msfec_problem_free(msfec_problem *p);

/* Host-only: shape queries (no device needed). */
int msfec_k(int pairing);                         /* 8, 20, 18, 7                  */
int msfec_n_fine_dofs(int pairing, int n_refine_local, int *n_block0, int *n_block1);

/* Create a context on CUDA device `device`.  Builds the shared fine-grid topology,
 * sparsity patterns and assembly tables on the host and uploads them.  Replaces the
 * per-cell setup_grid / setup_system_matrix / setup_basis_dofs_* of the reference
 * (ned_rt_basis.cc:167-359), which are identical for every coarse cell.
 * device < 0 creates a host-only context (tables only; for tests and introspection). */
int msfec_create(int device, const msfec_problem *p, msfec_ctx **out);
void msfec_destroy(msfec_ctx *ctx);
const char *msfec_last_error(const msfec_ctx *ctx); /* ctx may be NULL: global slot */

/* THE hot path: replaces `for (basis : cell_basis_map) basis.run()`
 * (ned_rt_global.cc:91-98 -> NedRTBasis::run, ned_rt_basis.cc:1280-1428; same for
 * Q / Q_Ned / RT_DQ) followed by get_global_element_matrix()/get_global_element_rhs()
 * (ned_rt_basis.h:131-137).
 *   corners     [n_cells][8][3]  coarse-cell vertices, deal.II vertex order (host)
 *   cell_ids    [n_cells] global coarse-cell index (keys the random field), or NULL
 *   elem_matrix [n_cells][k][k]  row-major, sigma-type coarse DoFs first (host, out)
 *   elem_rhs    [n_cells][k]     (host, out)
 *   stats       optional                                                            */
int msfec_build_basis(msfec_ctx *ctx, int n_cells, const double *corners,
                      const int64_t *cell_ids, double *elem_matrix, double *elem_rhs,
                      msfec_stats *stats);

/* Same, with corners / cell_ids / outputs already resident in device memory of the
 * context's GPU (no host<->device copies inside). */
int msfec_build_basis_device(msfec_ctx *ctx, int n_cells, const double *d_corners,
                             const int64_t *d_cell_ids, double *d_elem_matrix,
                             double *d_elem_rhs, msfec_stats *stats);

/* Replaces NedRTBasis::set_global_weights (ned_rt_basis.cc:1156-1181): forms
 * u_fine = sum_i w_i b_i for every cell of the last build; kept on the device.
 *   weights [n_cells][k] (host) */
int msfec_set_weights(msfec_ctx *ctx, int n_cells, const double *weights);

/* Fetch the fine-scale solution of one cell of the last build after set_weights
 * (what output_global_solution_in_cell writes, ned_rt_basis.cc:1092-1147).
 *   block0 [n_block0], block1 [n_block1] in the library's fine numbering
 *   (see msfec_fine_dof_layout). Either pointer may be NULL. */
int msfec_get_fine_solution(msfec_ctx *ctx, int cell, double *block0, double *block1);

/* Squared fine-grid norms of the reconstructed solution of every cell of the last build
 * (after msfec_set_weights): norms[n_cells][4] = { ||b0||^2_L2, |b0|^2, ||b1||^2_L2, |b1|^2 }
 * with unit coefficients, |.| the natural semi-norm of the block's element (H1 for nodal,
 * H(curl) for edge, H(div) for face DoFs; 0 for cell-wise constants), Gauss 2x2x2 per fine
 * cell.  The reconstruction is linear in the weights, so the norm of the DIFFERENCE of two
 * multiscale solutions is obtained by passing the difference of their weights to
 * msfec_set_weights.  The caller sums over cells (and ranks: one FP64 all-reduce, the
 * error-norm reduction of SURVEY.md s.8(e)) and takes the square root.  The reference
 * computes no error norm (integrate_difference never appears); this is the harness-defined
 * norm of the north star.  norms is a host buffer. */
int msfec_solution_norms(msfec_ctx *ctx, int n_cells, double *norms);

/* Fetch basis function `basis` (0..k_solve-1) of one cell of the last build, full
 * length incl. boundary values (what basis_curl_v / basis_div_v hold). */
int msfec_get_basis(msfec_ctx *ctx, int cell, int basis, double *block0, double *block1);

/* Position (in units of the fine mesh width h, relative to the coarse-cell origin)
 * and axis of every fine DoF of a block: pos[n][3], axis[n] (-1 for vertices/cells).
 * Lets a caller map the library's numbering to its own.  Host-only. */
int msfec_fine_dof_layout(const msfec_ctx *ctx, int block, double *pos, int32_t *axis,
                          uint8_t *on_boundary);

/* Introspection of host-built tables (tests, debugging).  name is one of the strings
 * documented in DESIGN.md; *count receives the element count; out may be NULL to
 * query the size.  dtype: 0 = int32, 1 = float64. */
int msfec_debug_table(const msfec_ctx *ctx, const char *name, void *out, size_t *count,
                      int *dtype);
/* Copy per-cell assembled slot values of the last build batch back to the host
 * (values[n_slots] of one cell). Requires problem.cells_per_batch >= n_cells. */
int msfec_debug_cell_values(msfec_ctx *ctx, int cell, double *values, size_t *count);

int msfec_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MSFEC_H_ */
