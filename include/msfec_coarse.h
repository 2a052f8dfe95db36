/*
 * msfec_coarse.h -- C ABI of the device solve of the GLOBAL COARSE system (part of libmsfec_b200.so).
 *
 * Replaces, for the host driver of this repository, the reference's Trilinos solves of the coarse problem:
 *   Ned_RT / Q_Ned / RT_DQ: Schur-complement CG with an inner CG on block (0,0)
 *       (source/Ned_RT/ned_rt_global.cc:330-461, q_ned_global.cc:331-470, rt_dq_global.cc:330-470),
 *   Q: CG on the SPD system (source/Q/q_global.cc:327-363).
 * The coarse system is small next to the basis build (C5: 205 920 unknowns) but its nested iteration is tens of
 * thousands of sparse matrix-vector products; on the host cores that is a minute, on the otherwise idle B200 about a
 * second.  All vectors and the four CSR blocks stay in HBM; the CG scalars (alpha, beta and the dot products they come
 * from) stay on the device and are consumed by the next kernel, the host reads one number per iteration (the residual
 * norm of the stopping test).  Dot products are reduced in a fixed order (per-block partial sums, summed by the last
 * block to finish), so the result is reproducible from run to run.
 *
 * The caller passes the FREE part of the system (essential unknowns eliminated), split into the blocks
 *     [A00 A01] [x0]   [b0]
 *     [A10 A11] [x1] = [b1],     A00 SPD,  A01 = -A10^T,  A11 symmetric positive semi-definite
 * as CSR with 32-bit indices (host pointers).  n1 = 0: CG on A00 x0 = b0, the other blocks are ignored (may be NULL).
 * Preconditioners as in host/coarse.cpp: Jacobi on A00, diag(A11) + diag(A10 diag(A00)^-1 A10^T) on the Schur
 * complement.  Stopping tests: ||r|| <= rtol * ||b|| on the recurrence residual (inner: rtol_inner, outer: rtol_outer).
 *
 * Returns 0 on success, non-zero otherwise (msfec_coarse_last_error() describes the failure: no device, CUDA error,
 * iteration cap reached).  There is no CPU fallback inside this call.
 */
#ifndef MSFEC_COARSE_H_
#define MSFEC_COARSE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct msfec_coarse_csr {
  int32_t n_rows, n_cols;
  const int32_t *ptr; /* [n_rows + 1] */
  const int32_t *col; /* [ptr[n_rows]] */
  const double *val;  /* [ptr[n_rows]] */
} msfec_coarse_csr;

typedef struct msfec_coarse_stats {
  int32_t outer_iterations;   /* Schur-complement CG (0 when n1 = 0) */
  int64_t inner_iterations;   /* CG on block (0,0), summed over all inner solves */
  int64_t kernel_launches;
  double ms_device;           /* CUDA-event time of the whole solve */
} msfec_coarse_stats;

int msfec_coarse_solve_device(int device, const msfec_coarse_csr *A00, const msfec_coarse_csr *A01,
                              const msfec_coarse_csr *A10, const msfec_coarse_csr *A11, const double *b0,
                              const double *b1, double rtol_inner, double rtol_outer, double *x0, double *x1,
                              msfec_coarse_stats *stats);

const char *msfec_coarse_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MSFEC_COARSE_H_ */
