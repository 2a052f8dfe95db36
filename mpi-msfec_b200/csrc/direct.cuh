// Batched direct solver kernels: block-tridiagonal LDL^T without pivoting (see DirectPlan in
// topology.h), replacing NedRTBasis::solve_direct (reference ned_rt_basis.cc:579-634, UMFPACK
// re-factorised for each of the k right-hand sides) for all cells and all k rhs at once.
//
// Per cell the "band" holds, for every z-slab s, a column-major panel of ld_s rows
// [slab s | slab s+1 | 32 rhs rows] x bs_s columns.  Right-looking blocked factorisation with
// panel width 32:
//   k_direct_panel   factor the 32x32 diagonal block (one warp, shared memory), then
//                    row-parallel triangular solve of everything below it (incl. the rhs rows:
//                    forward substitution is fused into the factorisation)
//   k_direct_update  trailing update C -= L_r D L_c^T on FP64 tensor cores
//                    (mma.sync.m8n8k4.f64, 64x64 tiles, 32x32 per warp)
//   k_direct_backward  L^T x = z for all k rhs, one CTA per cell
// Included by engine.cu.
#pragma once

namespace msfec {
namespace {

constexpr int kDP = 32;   // panel width (DirectPlan::kPanel)
constexpr int kMaxWindow = 6;   // panels per delayed update window

// ---- fill ------------------------------------------------------------------------------
// per-cell slot entries.  grid (ceil(ne/8), groups), block (32, 8); cell = g*32+lane
__global__ void k_direct_fill_cell(int ne, const int *__restrict__ dest, const int *__restrict__ ref,
                                   const double *__restrict__ vals, int n_slots, int g0, int cell_lo, int cell_hi,
                                   double *__restrict__ band, size_t band_stride) {
  const int lane = threadIdx.x, g = g0 + blockIdx.y;
  const int e = blockIdx.x * blockDim.y + threadIdx.y;
  const int cell = g * kLanes + lane;
  if (e >= ne || cell < cell_lo || cell >= cell_hi) return;
  const int r = ref[e];
  double a = vals[((size_t)g * n_slots + (r >> 1)) * kLanes + lane];
  if (r & 1) a = -a;
  band[(size_t)(cell - cell_lo) * band_stride + dest[e]] = a;
}

// cell-independent entries.  grid (ceil(ne/256), cells), block 256
__global__ void k_direct_fill_shared(int ne, const int *__restrict__ dest, const double *__restrict__ val,
                                     double scale, double *__restrict__ band, size_t band_stride) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  band[(size_t)blockIdx.y * band_stride + dest[e]] = val[e] * scale;
}

// right-hand sides: rows of the lifted rhs b[g][row][k][32] -> rhs rows of the band.
// grid (ceil(NI/8), groups), block (32, 8)
__global__ void k_direct_fill_rhs(int NI, int k, const int *__restrict__ rhs_dest, const double *__restrict__ b,
                                  int g0, int cell_lo, int cell_hi, double *__restrict__ band, size_t band_stride) {
  const int lane = threadIdx.x, g = g0 + blockIdx.y;
  const int row = blockIdx.x * blockDim.y + threadIdx.y;
  const int cell = g * kLanes + lane;
  if (row >= NI || cell < cell_lo || cell >= cell_hi) return;
  const int d = rhs_dest[row];
  if (d < 0) return;
  double *o = band + (size_t)(cell - cell_lo) * band_stride + d;
  for (int j = 0; j < k; ++j) o[j] = b[(((size_t)g * NI + row) * k + j) * kLanes + lane];
}

// ---- panel factorisation ---------------------------------------------------------------
// LDL^T of a 32x32 block held one row per lane (registers + shuffles, no shared-memory
// round trips).  On return lane i holds row i of the unit-lower factor in a[0..i-1] and the
// pivot d_i in a[i]; the function returns d_lane.
__device__ __forceinline__ double ldl32_rows(double (&a)[kDP], int lane, int *bad) {
  double my_d = 0.0;
#pragma unroll
  for (int p = 0; p < kDP; ++p) {
    const double ap = a[p];                                  // unscaled column p, one entry per lane
    const double d = __shfl_sync(0xffffffffu, ap, p);
    const double dinv = 1.0 / d;
    const double l = ap * dinv;
    if (lane == p) {
      my_d = d;
      if (!(fabs(d) > 1e-300) || !isfinite(d)) atomicExch(bad, 1);
    }
#pragma unroll
    for (int j = p + 1; j < kDP; ++j) {
      const double ajp = __shfl_sync(0xffffffffu, ap, j);
      if (j <= lane) a[j] = fma(-l, ajp, a[j]);
    }
    if (lane > p) a[p] = l;
  }
  return my_d;
}

// Factor the 32x32 diagonal block at (j0, j0) of every cell: one warp per cell, publishes the unit-lower
// factor (column-major, pivots on the diagonal) to diagL and the pivots to dvec.
// grid (ceil(cells/4)), block 128
__global__ void __launch_bounds__(128)
k_direct_diag(const double *__restrict__ band, size_t band_stride, long long col_off, int ld, int j0, int pglob, int NP,
              int n_cells, double *__restrict__ diagL, double *__restrict__ dvec, int *__restrict__ bad) {
  const int cell = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (cell >= n_cells) return;
  const double *P = band + (size_t)cell * band_stride + col_off;
  double a[kDP];
#pragma unroll
  for (int p = 0; p < kDP; ++p) a[p] = (p <= lane) ? P[(size_t)(j0 + p) * ld + j0 + lane] : 0.0;
  const double di = ldl32_rows(a, lane, bad);
  double *dl = diagL + ((size_t)cell * NP + pglob) * kDP;
#pragma unroll
  for (int p = 0; p < kDP; ++p) dl[(size_t)p * kDP + lane] = a[p];
  dvec[(size_t)cell * NP + pglob + lane] = di;
}

// Row-parallel triangular solve of the virtual rows v in [j0+32, ld) (rest of slab s, slab s+1 and the rhs
// rows: forward substitution is fused into the factorisation) against the factored diagonal block.
// grid (ceil(nrows/128), cells), block 128
__global__ void __launch_bounds__(128)
k_direct_panel(double *__restrict__ band, size_t band_stride, long long col_off, int ld, int j0, int pglob, int NP,
               const double *__restrict__ diagL, const double *__restrict__ dvec, double *__restrict__ ybuf, int slot,
               int ldy) {
  __shared__ double Ld[kDP][kDP + 1];
  __shared__ double dinv[kDP];
  const int cell = blockIdx.y, tid = threadIdx.x;
  double *P = band + (size_t)cell * band_stride + col_off;
  const int v = j0 + kDP + blockIdx.x * 128 + tid;
  double y[kDP];
  if (v < ld) {
#pragma unroll
    for (int p = 0; p < kDP; ++p) y[p] = P[(size_t)(j0 + p) * ld + v];
  }
  const double *dl = diagL + ((size_t)cell * NP + pglob) * kDP;
  for (int idx = tid; idx < kDP * kDP; idx += 128) {
    const int i = idx & 31, p = idx >> 5;
    Ld[i][p] = dl[(size_t)p * kDP + i];
  }
  if (tid < kDP) dinv[tid] = 1.0 / dvec[(size_t)cell * NP + pglob + tid];
  __syncthreads();
  if (v >= ld) return;
#pragma unroll
  for (int p = 1; p < kDP; ++p) {
    double s = y[p];
#pragma unroll
    for (int q = 0; q < p; ++q) s = fma(-y[q], Ld[p][q], s);
    y[p] = s;
  }
  // L = y D^-1 goes back in place; the trailing updates use C -= L (y)^T, so -y is kept in the window scratch
  // (slot = position of this panel inside the current update window) to keep the MMA loop free of FP64 ALU work
  double *Y = ybuf + ((size_t)cell * kMaxWindow + slot) * kDP * ldy;
#pragma unroll
  for (int p = 0; p < kDP; ++p) {
    P[(size_t)(j0 + p) * ld + v] = y[p] * dinv[p];
    Y[(size_t)p * ldy + v] = -y[p];
  }
}

// ---- trailing update on FP64 tensor cores ----------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// stage a TROWS-row x 32-column block of one source panel (global: column-major, rows contiguous) into
// shared memory [p][i] with row stride TROWS+8; rows >= ld are zero-filled.  NT threads.
template <int TROWS, int NT>
__device__ __forceinline__ void stage_block(double *dst, const double *P, int ld, int n_rows, int jsrc, int row0, int tid) {
  constexpr int LDS = TROWS + 8, CPC = TROWS / 2;          // 16-byte chunks per column
#pragma unroll
  for (int t = 0; t < CPC * kDP / NT; ++t) {
    const int chunk = tid + t * NT;
    const int p = chunk / CPC, i = (chunk % CPC) * 2;
    double *d = dst + p * LDS + i;
    if (row0 + i < n_rows) cp_async16(d, P + (size_t)(jsrc + p) * ld + row0 + i);
    else { d[0] = 0.0; d[1] = 0.0; }
  }
}

// Device copy of the block structure (DirectPlan): all lookups below are warp-uniform.
struct DirectPlanDev {
  int n_slabs, NP;
  const int *bs, *slab_off, *ld, *front_rows;
  const long long *col_off;
  const int *chunk_off, *chunk_blk, *chunk_local, *front_pos;
};

template <int TM, int TN>
constexpr size_t update_smem_bytes(int max_src) {
  return ((size_t)max_src * kDP * (TN + 8) + 2 * (size_t)kDP * (TM + 8)) * sizeof(double);
}

// C(vr, vc) -= sum_{p in source panels} L(vr, p) d_p L(vc, p) for target columns vc in [vc_lo, vc_hi) and
// rows vr >= vc of block column s (virtual index space [slab s | slab s+1 | rhs rows]); targets in
// columns >= bs live in block column s+1.  The nq source panels (32 columns each, starting at jsrc) are
// applied in one pass so every C tile is read and written once per window (K = 32 nq).
// grid (column tiles, row splits Z, cells); block 128 = (TM/32) x (TN/32) warps of 32x32.  A CTA keeps
// the column operands of all source panels in shared memory and walks its row tiles with a cp.async
// double buffer; the C tile is loaded straight into the DMMA accumulators.
template <int TM, int TN>
__global__ void __launch_bounds__((TM / 32) * (TN / 32) * 32)
k_direct_update(double *__restrict__ band, size_t band_stride, DirectPlanDev D, int s, int jsrc, int nq, int vc_lo,
                int vc_hi, const double *__restrict__ ybuf, int ldy) {
  constexpr int LDR = TM + 8, LDC = TN + 8, WN = TN / 32, NT = (TM / 32) * (TN / 32) * 32;
  extern __shared__ __align__(16) double upd_smem[];
  double *Lc = upd_smem;                                   // [nq][32][LDC]
  double *Lr = Lc + (size_t)nq * kDP * LDC;                // [2][32][LDR]
  const int tj = blockIdx.x, cell = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cbase = vc_lo + tj * TN;
  const int ld = D.ld[s], front_rows = D.front_rows[s];
  const long long col_off = D.col_off[s];
  const int choff = D.chunk_off[s];
  double *cb = band + (size_t)cell * band_stride;
  const double *P = cb + col_off;
  // row tiles of this CTA: rbase = r_first + (z + Z*i) * TM, first tile contains row cbase
  const int r_first = cbase - (cbase - vc_lo) % TM;        // TM-aligned relative to vc_lo
  const int T = (ld - r_first + TM - 1) / TM;
  const int Z = gridDim.y;
  if ((int)blockIdx.y >= T) return;
  // column operands: -y = -(L D) of the nq source panels, written by k_direct_panel into the window scratch
  const double *Yc = ybuf + (size_t)cell * kMaxWindow * kDP * ldy;
  for (int q = 0; q < nq; ++q) stage_block<TN, NT>(Lc + (size_t)q * kDP * LDC, Yc + (size_t)q * kDP * ldy, ldy, ld, 0, cbase, tid);
  stage_block<TM, NT>(Lr, P, ld, ld, jsrc, r_first + blockIdx.y * TM, tid);
  cp_async_commit();
  const int wr = warp / WN, wc = warp % WN;
  const int fr = lane >> 2, fk = lane & 3;
  const int vc0 = cbase + wc * 32;
  const bool col_ok = vc0 < vc_hi && vc0 < front_rows;
  // destination of this warp's columns: block column s itself, or the block column of the reached block
  double *cdst = cb;
  int ldc = ld, cblk = s;
  if (col_ok) {
    cblk = D.chunk_blk[choff + (vc0 >> 5)];
    if (cblk == s) cdst = cb + col_off + (size_t)vc0 * ld;
    else { ldc = D.ld[cblk]; cdst = cb + D.col_off[cblk] + (size_t)D.chunk_local[choff + (vc0 >> 5)] * ldc; }
  }
  int buf = 0;
  for (int ti = blockIdx.y; ti < T; ti += Z) {
    const int rbase = r_first + ti * TM;
    const int vr0 = rbase + wr * 32;
    const bool active = col_ok && vr0 >= vc0 && vr0 < ld;
    int roff = vr0;
    if (active && cblk != s) {
      const int rb = D.chunk_blk[choff + (vr0 >> 5)];
      roff = rb < 0 ? D.front_rows[cblk] + (vr0 - front_rows)
                    : D.front_pos[cblk * D.n_slabs + rb] + D.chunk_local[choff + (vr0 >> 5)];
    }
    // C is fetched into its own registers now and consumed after the MMAs, so its latency hides
    // behind the nq source panels; the product L (-y)^T is accumulated from zero.
    double acc[4][4][2], cold[4][4][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    if (active) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            cold[mt][nt][h] = cdst[(size_t)(nt * 8 + fk * 2 + h) * ldc + roff + mt * 8 + fr];
    }
    for (int q = 0; q < nq; ++q) {
      // prefetch the next (row tile, source panel) operand block
      if (q + 1 < nq) stage_block<TM, NT>(Lr + (buf ^ 1) * kDP * LDR, P, ld, ld, jsrc + (q + 1) * kDP, rbase, tid);
      else if (ti + Z < T) stage_block<TM, NT>(Lr + (buf ^ 1) * kDP * LDR, P, ld, ld, jsrc, rbase + Z * TM, tid);
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      if (active) {
        const double *A = Lr + buf * kDP * LDR + wr * 32 + fr;
        const double *B = Lc + (size_t)q * kDP * LDC + wc * 32 + fr;
#pragma unroll
        for (int ks = 0; ks < kDP / 4; ++ks) {
          const int kk = ks * 4 + fk;
          double af[4], bf[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            af[t] = A[kk * LDR + t * 8];
            bf[t] = B[kk * LDC + t * 8];
          }
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
      }
      __syncthreads();
      buf ^= 1;
    }
    if (active) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            cdst[(size_t)(nt * 8 + fk * 2 + h) * ldc + roff + mt * 8 + fr] = cold[mt][nt][h] + acc[mt][nt][h];
    }
  }
  cp_async_wait<0>();
}

// ---- backward substitution -------------------------------------------------------------
// L^T x = z.  One CTA (128 threads) per cell; xT[cell][j][NP] is both output and the running
// solution read by later (lower-numbered) panels.  Per panel of 32 columns:
//  (1) T = Z - L(below)^T X as a (32 x K) x (K x 24) product on the FP64 tensor cores; the K rows
//      below the panel are split over the 4 warps, DMMA fragments are read straight from global
//      memory (each quad of lanes reads one full 32-byte sector), partial sums meet in shared memory;
//  (2) unit upper-triangular solve with the 32x32 diagonal block, one thread per rhs, column oriented.
__global__ void __launch_bounds__(128, 6)
k_direct_backward(const double *__restrict__ band, size_t band_stride, DirectPlanDev D, const double *__restrict__ diagL,
                  int k, double *xT) {
  __shared__ double tt[kDP][24 + 1];
  __shared__ double Ld[kDP][kDP + 1];
  const int cell = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fr = lane >> 2, fk = lane & 3;
  const int NP = D.NP;
  double *x = xT + (size_t)cell * k * NP;
  for (int s = D.n_slabs - 1; s >= 0; --s) {
    const int bs = D.bs[s], ld = D.ld[s], so = D.slab_off[s], choff = D.chunk_off[s];
    const int rows_dof = D.front_rows[s], rhs_row = rows_dof;
    const double *P = band + (size_t)cell * band_stride + D.col_off[s];
    for (int j0 = bs - kDP; j0 >= 0; j0 -= kDP) {
      // tt := z (rhs rows of the panel columns), padded rhs columns := 0
      for (int idx = tid; idx < kDP * 24; idx += 128) {
        const int c = idx / 24, j = idx % 24;
        tt[c][j] = j < k ? P[(size_t)(j0 + c) * ld + rhs_row + j] : 0.0;
      }
      const double *dl = diagL + ((size_t)cell * NP + so + j0) * kDP;
      for (int idx = tid; idx < kDP * kDP; idx += 128) {
        const int i = idx & 31, p = idx >> 5;
        Ld[i][p] = dl[(size_t)p * kDP + i];
      }
      __syncthreads();
      const int r_lo = j0 + kDP, nsteps = (rows_dof - r_lo) / 4;       // rows_dof, r_lo multiples of 32
      if (nsteps > 0) {
        double acc[4][3][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        // rows are multiples of 32, so nsteps is a multiple of 8: each of the 4 warps takes pairs of k-steps
        // and keeps two of them (14 independent sector loads) in flight
        for (int st = warp * 2; st < nsteps; st += 8) {
          double af[2][4], bf[2][3];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = r_lo + (st + u) * 4 + fk;
            const int ch = choff + (i >> 5);                  // front row -> padded unknown index
            const int xi = D.slab_off[D.chunk_blk[ch]] + D.chunk_local[ch] + (i & 31);
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) af[u][mt] = P[(size_t)(j0 + mt * 8 + fr) * ld + i];   // A[m=c][k=i] = L(i, c)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
              const int j = nt * 8 + fr;
              bf[u][nt] = j < k ? x[(size_t)j * NP + xi] : 0.0;                                  // B[k=i][n=j] = x_i^(j)
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
              for (int nt = 0; nt < 3; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], af[u][mt], bf[u][nt]);
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) atomicAdd(&tt[mt * 8 + fr][nt * 8 + fk * 2 + h], -acc[mt][nt][h]);
      }
      __syncthreads();
      if (tid < k) {
        double t[kDP];
#pragma unroll
        for (int c = 0; c < kDP; ++c) t[c] = tt[c][tid];
#pragma unroll
        for (int c = kDP - 1; c >= 0; --c) {
          const double xc = t[c];
#pragma unroll
          for (int i = 0; i < c; ++i) t[i] = fma(-Ld[c][i], xc, t[i]);
        }
#pragma unroll
        for (int c = 0; c < kDP; ++c) x[(size_t)tid * NP + so + j0 + c] = t[c];
      }
      __syncthreads();
    }
  }
}

// xT[cell][j][p] -> interleaved x[g][row][k][32].  grid (ceil(NP/256), cells), block 256
__global__ void k_direct_scatter_x(int NP, int NI, int k, const int *__restrict__ inv_perm, const double *__restrict__ xT,
                                   int cell_lo, double *__restrict__ x) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= NP) return;
  const int row = inv_perm[p];
  const int cell = cell_lo + blockIdx.y, g = cell / kLanes, lane = cell % kLanes;
  if (row < 0) return;
  for (int j = 0; j < k; ++j)
    x[(((size_t)g * NI + row) * k + j) * kLanes + lane] = xT[((size_t)blockIdx.y * k + j) * NP + p];
}

// rows whose solution is fixed to zero (RT_DQ pinned DoF) -- the interleaved x buffer is cleared first.

}  // namespace
}  // namespace msfec
