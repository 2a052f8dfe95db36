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
// grid (ceil(nrows/128), cells), block 128.  Virtual rows v in [j0+32, ld) are solved against the
// diagonal block at (j0, j0).  Every CTA factors the (read-only) diagonal block redundantly in
// shared memory; CTA x==0 publishes it to diagL / dvec.
__global__ void __launch_bounds__(128)
k_direct_panel(double *__restrict__ band, size_t band_stride, long long col_off, int ld, int j0, int pglob, int NP,
               double *__restrict__ diagL, double *__restrict__ dvec, int *__restrict__ bad) {
  __shared__ double Ld[kDP][kDP + 1];
  __shared__ double dinv[kDP];
  const int cell = blockIdx.y, tid = threadIdx.x;
  double *P = band + (size_t)cell * band_stride + col_off;
  for (int idx = tid; idx < kDP * kDP; idx += 128) {
    const int i = idx & 31, p = idx >> 5;
    Ld[i][p] = (i >= p) ? P[(size_t)(j0 + p) * ld + j0 + i] : 0.0;
  }
  __syncthreads();
  if (tid < 32) {
    const int i = tid;
    for (int p = 0; p < kDP; ++p) {
      const double d = Ld[p][p];
      const double l = Ld[i][p] / d;
      __syncwarp();
      if (i > p)
        for (int j = p + 1; j <= i; ++j) Ld[i][j] -= l * Ld[j][p];
      __syncwarp();
      if (i > p) Ld[i][p] = l;
      if (i == p) {
        dinv[p] = 1.0 / d;
        if (!(fabs(d) > 1e-300) || !isfinite(d)) atomicExch(bad, 1);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    double *dl = diagL + ((size_t)cell * NP + pglob) * kDP;
    for (int idx = tid; idx < kDP * kDP; idx += 128) {
      const int i = idx & 31, p = idx >> 5;
      dl[(size_t)p * kDP + i] = Ld[i][p];          // column p of the unit-lower factor (diag holds d_p)
    }
    if (tid < kDP) dvec[(size_t)cell * NP + pglob + tid] = Ld[tid][tid];
  }
  const int v = j0 + kDP + blockIdx.x * 128 + tid;
  if (v >= ld) return;
  double y[kDP];
#pragma unroll
  for (int p = 0; p < kDP; ++p) y[p] = P[(size_t)(j0 + p) * ld + v];
#pragma unroll
  for (int p = 1; p < kDP; ++p) {
    double s = y[p];
#pragma unroll
    for (int q = 0; q < p; ++q) s = fma(-y[q], Ld[p][q], s);
    y[p] = s;
  }
#pragma unroll
  for (int p = 0; p < kDP; ++p) P[(size_t)(j0 + p) * ld + v] = y[p] * dinv[p];
}

// ---- trailing update on FP64 tensor cores ----------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// C(vr, vc) -= sum_p L(vr, p) d_p L(vc, p) over the trailing region vr >= vc >= j0+32 of block
// column s (virtual index space [slab s | slab s+1 | rhs rows]); targets in columns >= bs live in
// block column s+1.  grid (T, T, cells) with T = ceil((ld - j0 - 32) / 64); block 128 (2x2 warps).
__global__ void __launch_bounds__(128)
k_direct_update(double *__restrict__ band, size_t band_stride, long long col_off, int ld, int bs, int bs_next,
                long long col_off_next, int ld_next, int rhs_row_next, int j0, int pglob, int NP,
                const double *__restrict__ dvec) {
  const int ti = blockIdx.x, tj = blockIdx.y;
  if (tj > ti) return;
  constexpr int LDS = kDP + 4;   // padded row stride (doubles)
  __shared__ double Lr[64][LDS];
  __shared__ double Lc[64][LDS];
  const int cell = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double *cb = band + (size_t)cell * band_stride;
  const double *P = cb + col_off;
  const int v0 = j0 + kDP;
  const int rbase = v0 + ti * 64, cbase = v0 + tj * 64;
  const double *dv = dvec + (size_t)cell * NP + pglob;
  // stage the two 64 x 32 operand blocks (column-major in global: contiguous along rows)
  for (int idx = tid; idx < 64 * kDP; idx += 128) {
    const int i = idx & 63, p = idx >> 6;
    const int vr = rbase + i, vc = cbase + i;
    Lr[i][p] = vr < ld ? P[(size_t)(j0 + p) * ld + vr] : 0.0;
    Lc[i][p] = vc < ld ? P[(size_t)(j0 + p) * ld + vc] * dv[p] : 0.0;
  }
  __syncthreads();
  const int wr = warp >> 1, wc = warp & 1;
  const int vr0 = rbase + wr * 32, vc0 = cbase + wc * 32;
  if (vr0 < vc0 || vr0 >= ld || vc0 >= bs + bs_next) return;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
  for (int ks = 0; ks < kDP / 4; ++ks) {
    double af[4], bf[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      af[t] = Lr[wr * 32 + t * 8 + fr][ks * 4 + fk];
      bf[t] = Lc[wc * 32 + t * 8 + fr][ks * 4 + fk];
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
  }
  // destination mapping of this warp's 32x32 sub-block (uniform per warp; all bounds are multiples of 32)
  double *cdst;
  int ldc, roff;
  if (vc0 < bs) { cdst = cb + col_off + (size_t)vc0 * ld; ldc = ld; roff = vr0; }
  else {
    cdst = cb + col_off_next + (size_t)(vc0 - bs) * ld_next; ldc = ld_next;
    roff = vr0 < bs + bs_next ? vr0 - bs : rhs_row_next + (vr0 - bs - bs_next);
  }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double *p = cdst + (size_t)(nt * 8 + fk * 2 + h) * ldc + roff + mt * 8 + fr;
        *p -= acc[mt][nt][h];
      }
}

// ---- backward substitution -------------------------------------------------------------
struct DirectPlanDev {
  int n_slabs, NP;
  const int *bs, *slab_off, *ld;
  const long long *col_off;
};

// L^T x = z.  One CTA (256 threads) per cell; xT[cell][j][NP] is both output and the running
// solution read by later (lower-numbered) panels.
__global__ void __launch_bounds__(256)
k_direct_backward(const double *__restrict__ band, size_t band_stride, DirectPlanDev D, const double *__restrict__ diagL,
                  int k, double *xT) {
  __shared__ double tt[kDP][kMaxK + 1];
  __shared__ double Ld[kDP][kDP + 1];
  const int cell = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NP = D.NP;
  double *x = xT + (size_t)cell * k * NP;
  for (int s = D.n_slabs - 1; s >= 0; --s) {
    const int bs = D.bs[s], ld = D.ld[s], so = D.slab_off[s];
    const int rows_dof = ld - DirectPlan::kRhsRows, rhs_row = rows_dof;
    const double *P = band + (size_t)cell * band_stride + D.col_off[s];
    for (int j0 = bs - kDP; j0 >= 0; j0 -= kDP) {
      // phase 1: t_c = z_c - sum_{i >= j0+32} L(i, c) x_i   (warp w: columns j0 + 4w .. 4w+3)
      for (int cc = 0; cc < 4; ++cc) {
        const int c = j0 + warp * 4 + cc;
        const double *col = P + (size_t)c * ld;
        double acc[kMaxK];
#pragma unroll
        for (int j = 0; j < kMaxK; ++j) acc[j] = 0.0;
        for (int i = j0 + kDP + lane; i < rows_dof; i += 32) {
          const double l = col[i];
#pragma unroll
          for (int j = 0; j < kMaxK; ++j)
            if (j < k) acc[j] = fma(l, x[(size_t)j * NP + so + i], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < kMaxK; ++j) {
          if (j < k) {
            double a = acc[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == j) tt[warp * 4 + cc][j] = col[rhs_row + j] - a;
          }
        }
      }
      const double *dl = diagL + ((size_t)cell * NP + so + j0) * kDP;
      for (int idx = tid; idx < kDP * kDP; idx += 256) {
        const int i = idx & 31, p = idx >> 5;
        Ld[i][p] = dl[(size_t)p * kDP + i];
      }
      __syncthreads();
      // phase 2: unit upper-triangular solve with the diagonal block, one thread per rhs
      if (tid < k) {
        double xs[kDP];
#pragma unroll
        for (int c = kDP - 1; c >= 0; --c) {
          double v = tt[c][tid];
#pragma unroll
          for (int i = c + 1; i < kDP; ++i) v = fma(-Ld[i][c], xs[i], v);
          xs[c] = v;
        }
#pragma unroll
        for (int c = 0; c < kDP; ++c) x[(size_t)tid * NP + so + j0 + c] = xs[c];
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
