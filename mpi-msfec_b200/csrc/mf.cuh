// Batched multifrontal LDL^T kernels (plan: mfplan.h).  Replaces NedRTBasis::solve_direct / solve_iterative
// (reference source/Ned_RT/ned_rt_basis.cc:579-634, :637-847 and the Q / Q_Ned / RT_DQ siblings) for all cells and
// all k right-hand sides at once.
//
// One CTA = one front of one coarse cell; one launch = all fronts of one tree level of all cells of the sub-batch.
//   k_mf_forward   panel [m x s8] in shared memory := matrix entries (per-cell slot values on the shared pattern,
//                  cell-independent coupling entries, padding pivots) + rhs rows + the leading columns of the
//                  children's contribution blocks (extend-add through `cmap`);
//                  LDL^T of the own block / X = A L^-T of the rows below in 8-column steps: the left-looking update
//                  runs on mma.sync.m8n8k4.f64, the 8 x 8 pivot tile is factored redundantly by every warp in
//                  registers (no broadcast, one barrier less), the rows below are solved one thread per row;
//                  factor panel -> global (read again by k_mf_backward only);
//                  contribution block C = sum_children C_child - X L21^T, 8 x 8 tiles on the FP64 tensor cores,
//                  children gathered through `pinv` straight into the accumulators, 16-byte stores.
//   k_mf_backward  x_own = L11^-T (z_own - L21^T x_reached), top-down.
// Everything else of a front lives in shared memory, so HBM sees: slot values and rhs once, every factor panel
// written once and read once, every contribution block written once and read once (MfPlan::bytes).
// Included by engine.cu.
#pragma once

namespace msfec {
namespace {

struct MfDev {
  const MfFront *fronts;
  const MfChild *children;
  const int *front_idx, *own_rows, *cmap, *pinv, *pe_dest, *pe_ref, *ps_dest, *pc_dest, *level_fronts;
  const double *ps_val, *pc_val;
  int kr, NP;
};

constexpr int kMfMaxChildren = 8;

// dynamic shared memory of k_mf_forward (doubles first, then the children's inverse maps)
__host__ __device__ inline size_t mf_fwd_smem_bytes(int m, int ldx, int s8, int nch) {
  return ((size_t)m * ldx + 2 * (size_t)s8 + 8 * (size_t)s8) * sizeof(double) + (size_t)nch * m * sizeof(int);
}
// rows of L21 staged at once by k_mf_backward: the whole block when it is small, else ~48 KB worth (multiple of 8)
__host__ __device__ inline int mf_bwd_chunk(int s8, int u8) {
  const int cap = (6144 / s8) & ~7;
  return u8 < cap ? u8 : (cap < 8 ? 8 : cap);
}
__host__ __device__ inline size_t mf_bwd_smem_bytes(int s8, int u8, int kr) {
  return ((size_t)s8 * s8 + (size_t)s8 * kr + (size_t)u8 * kr + (size_t)mf_bwd_chunk(s8, u8) * s8) * sizeof(double);
}

// grid (fronts of the level, cells of the sub-batch), block NT
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_mf_forward(MfDev M, int lf_off, const double *__restrict__ vals, int n_slots, double kscale,
             const double *__restrict__ b, int NI, int k, int cell_lo, double *__restrict__ Lst, size_t l_stride,
             double *__restrict__ Cst, size_t c_stride, int *__restrict__ bad) {
  extern __shared__ __align__(16) double mf_smem[];
  __shared__ int ch_coff[kMfMaxChildren], ch_ldc[kMfMaxChildren];
  constexpr int NW = NT / 32;
  const int f = M.level_fronts[lf_off + blockIdx.x];
  const MfFront F = M.fronts[f];
  const int cell = blockIdx.y, gcell = cell_lo + cell, g = gcell / kLanes, ln = gcell % kLanes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fr = lane >> 2, fk = lane & 3;
  const int s8 = F.s8, u8 = F.u8, m = F.m, ldx = F.ldx, kr = M.kr;
  const int nch = F.ch_hi - F.ch_lo;
  double *P = mf_smem;                          // [m][ldx]   the panel
  double *dinv = P + (size_t)m * ldx;           // [s8]       1 / d
  double *dval = dinv + s8;                     // [s8]       d
  double *Ld = dval + s8;                       // [s8][8]    unit-lower factors of the 8 x 8 pivot tiles
  int *pinv_s = reinterpret_cast<int *>(Ld + 8 * (size_t)s8);   // [nch][m]

  // ---- assemble the panel --------------------------------------------------------------------------------
  {
    double2 *P2 = reinterpret_cast<double2 *>(P);
    const int n2 = (m * ldx) >> 1;
    for (int i = tid; i < n2; i += NT) P2[i] = make_double2(0.0, 0.0);
  }
  __syncthreads();
  {
    const double *vc = vals + (size_t)g * n_slots * kLanes + ln;
    for (int e = F.pe_lo + tid; e < F.pe_hi; e += NT) {
      const int ref = M.pe_ref[e];
      const double v = vc[(size_t)(ref >> 1) * kLanes];
      P[M.pe_dest[e]] = (ref & 1) ? -v : v;
    }
    for (int e = F.pc_lo + tid; e < F.pc_hi; e += NT) P[M.pc_dest[e]] = M.pc_val[e];
    const double *bc = b + (size_t)g * NI * k * kLanes + ln;
    for (int c = warp; c < s8; c += NW) {
      const int row = M.own_rows[F.row_off + c];
      if (row < 0) continue;
      for (int j = lane; j < k; j += 32) P[(size_t)(s8 + u8 + j) * ldx + c] = bc[((size_t)row * k + j) * kLanes];
    }
  }
  __syncthreads();
  for (int e = F.ps_lo + tid; e < F.ps_hi; e += NT) P[M.ps_dest[e]] += M.ps_val[e] * kscale;
  __syncthreads();
  // children: the leading n_own columns of a child's contribution block belong to this front's own columns
  for (int ci = 0; ci < nch; ++ci) {
    const MfChild ch = M.children[F.ch_lo + ci];
    const int cu8 = M.fronts[ch.front].u8, ccoff = M.fronts[ch.front].c_off;
    const int ldc = cu8 + kr;
    if (tid == 0) { ch_coff[ci] = ccoff; ch_ldc[ci] = ldc; }
    const double *Cc = Cst + (size_t)cell * c_stride + ccoff;
    const int *cmap = M.cmap + ch.cmap_off;
    for (int i = tid; i < m; i += NT) pinv_s[ci * m + i] = M.pinv[ch.pinv_off + i];
    for (int j = warp; j < ch.n_own; j += NW) {
      const int cj = cmap[j];
      const double *col = Cc + (size_t)j * ldc;
#pragma unroll 4
      for (int i = j + lane; i < ldc; i += 32) {
        const int ri = cmap[i];
        const double v = col[i];
        if (ri >= 0) P[(size_t)ri * ldx + cj] += v;
      }
    }
    __syncthreads();
  }

  // ---- factor: 8 columns at a time ---------------------------------------------------------------------------
  const int S = s8 / 8, MT = m / 8;
  for (int q = 0; q < S; ++q) {
    const int c0 = q * 8;
    if (q > 0) {
      // left-looking update of tile column q:  P(R, q) -= X(R, 0:c0) L(q, 0:c0)^T   (L = X D^-1), four row tiles of a
      // warp at a time (independent accumulator chains).  MMA m = column inside the tile, n = row, k = earlier columns
      const double *Arow = P + (size_t)(c0 + fr) * ldx + fk;
      for (int R0 = q + warp; R0 < MT; R0 += 4 * NW) {
        double acc[4][2];
        double *tp[4];
        const double *Brow[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int R = min(R0 + u * NW, MT - 1);
          tp[u] = P + (size_t)(R * 8 + 2 * fk) * ldx + c0 + fr;
          Brow[u] = P + (size_t)(R * 8 + fr) * ldx + fk;
          acc[u][0] = tp[u][0]; acc[u][1] = tp[u][ldx];
        }
        for (int t = 0; t < c0; t += 4) {
          const double a = -Arow[t] * dinv[t + fk];
#pragma unroll
          for (int u = 0; u < 4; ++u) dmma_m8n8k4(acc[u][0], acc[u][1], a, Brow[u][t]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (R0 + u * NW < MT) { tp[u][0] = acc[u][0]; tp[u][ldx] = acc[u][1]; }
      }
      __syncthreads();
    }
    // LDL^T of the pivot tile by warp 0: lane (i = lane & 7) = row, pivot column broadcast by shuffles; the unit-lower
    // factor and the pivots go to shared memory (an earlier version let every thread factor the tile redundantly in
    // registers: no barrier, but 4-8 x the FP64 work, which saturated the FP64 pipe of the small fronts)
    if (warp == 0) {
      const int i = lane & 7;
      double a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = (j <= i) ? P[(size_t)(c0 + i) * ldx + c0 + j] : 0.0;
      bool ok = true;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const double d = __shfl_sync(0xffffffffu, a[p], p);
        ok = ok && (fabs(d) > 1e-300) && isfinite(d);
        const double inv = rcp_newton(d);
        const double l = a[p] * inv;
#pragma unroll
        for (int j = p + 1; j < 8; ++j) {
          const double ajp = __shfl_sync(0xffffffffu, a[p], j);
          if (j <= i) a[j] = fma(-l, ajp, a[j]);
        }
        if (i > p) a[p] = l;
        if (lane == p) { dval[c0 + p] = d; dinv[c0 + p] = inv; }
      }
      if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) Ld[(size_t)(c0 + i) * 8 + j] = (j < i) ? a[j] : 0.0;
      }
      if (!ok && lane == 0) atomicExch(bad, 1);
    }
    __syncthreads();
    // rows below the pivot tile: X = A L^-T, one thread per row, L broadcast from shared memory
    for (int r = c0 + 8 + tid; r < m; r += NT) {
      double *row = P + (size_t)r * ldx + c0;
      const double *Lt = Ld + (size_t)c0 * 8;
      double x[8];
#pragma unroll
      for (int j = 0; j < 8; j += 2) { const double2 v = *reinterpret_cast<const double2 *>(row + j); x[j] = v.x; x[j + 1] = v.y; }
#pragma unroll
      for (int j = 1; j < 8; ++j) {
#pragma unroll
        for (int t = 0; t < j; t += 2) {
          const double2 lv = *reinterpret_cast<const double2 *>(Lt + j * 8 + t);
          x[j] = fma(-x[t], lv.x, x[j]);
          if (t + 1 < j) x[j] = fma(-x[t + 1], lv.y, x[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2 *>(row + j) = make_double2(x[j], x[j + 1]);
    }
    __syncthreads();
  }

  // ---- factor panel -> global: unit-lower L (pivot tiles hold D on the diagonal), L = X D^-1 below --------------
  {
    double *Lo = Lst + (size_t)cell * l_stride + F.l_off;
    for (int r = warp; r < s8; r += NW)
      for (int c = lane; c < s8; c += 32) {
        double v = 0.0;
        if ((r >> 3) == (c >> 3)) v = r == c ? dval[c] : (r > c ? Ld[(size_t)r * 8 + (c & 7)] : 0.0);
        else if (r > c) v = P[(size_t)r * ldx + c] * dinv[c];
        Lo[(size_t)r * s8 + c] = v;
      }
    // rows below the own block: a warp covers 32 / s8 rows at once when the panel is narrow
    const int rpw = s8 >= 32 ? 1 : 32 / s8;
    const int lr = s8 >= 32 ? 0 : lane / s8, lc = s8 >= 32 ? lane : lane - lr * s8;
    if (lr < rpw)
      for (int r = s8 + warp * rpw + lr; r < m; r += NW * rpw)
        for (int c = lc; c < s8; c += 32) Lo[(size_t)r * s8 + c] = P[(size_t)r * ldx + c] * dinv[c];
  }

  // ---- contribution block: C(I, J) = sum_children C_child - X_I L_J^T.  A warp step = one tile column J x four row tiles:
  // the children are gathered through `pinv` straight into the accumulators, the four MMA chains run interleaved ------
  if (u8 > 0) {
    const int UT = u8 / 8, RT = (u8 + kr) / 8, ldc = u8 + kr;
    double *Co = Cst + (size_t)cell * c_stride + F.c_off;
    const double *Cbase = Cst + (size_t)cell * c_stride;
    int J = 0, gi = warp;
    while (J < UT) {
      const int ng = (RT - J + 3) >> 2;
      if (gi >= ng) { gi -= ng; ++J; continue; }
      const int I0 = J + 4 * gi;
      const int colp = s8 + J * 8 + fr;
      double acc[4][2];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u][0] = acc[u][1] = 0.0;
      for (int ci = 0; ci < nch; ++ci) {
        const int *pv = pinv_s + ci * m;
        const int jc = pv[colp];
        if (jc < 0) continue;
        const double *cb = Cbase + ch_coff[ci] + (size_t)jc * ch_ldc[ci];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int I = I0 + u;
          if (I >= RT) continue;
          const int rowp = s8 + I * 8 + 2 * fk;
          const int2 ii = *reinterpret_cast<const int2 *>(pv + rowp);
          if (ii.x >= 0 && rowp >= colp) acc[u][0] += cb[ii.x];
          if (ii.y >= 0 && rowp + 1 >= colp) acc[u][1] += cb[ii.y];
        }
      }
      const double *Arow = P + (size_t)colp * ldx + fk;
      const double *Brow[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) Brow[u] = P + (size_t)(s8 + min(I0 + u, RT - 1) * 8 + fr) * ldx + fk;
      for (int t = 0; t < s8; t += 4) {
        const double a = -Arow[t] * dinv[t + fk];
#pragma unroll
        for (int u = 0; u < 4; ++u) dmma_m8n8k4(acc[u][0], acc[u][1], a, Brow[u][t]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (I0 + u < RT)
          *reinterpret_cast<double2 *>(Co + (size_t)(J * 8 + fr) * ldc + (I0 + u) * 8 + 2 * fk) = make_double2(acc[u][0], acc[u][1]);
      gi += NW;
    }
  }
}

// grid (fronts of the level, cells of the sub-batch), block NT.  xT[cell][k][NP] holds x in the padded elimination order.
template <int NT>
__global__ void __launch_bounds__(NT)
k_mf_backward(MfDev M, int lf_off, int k, const double *__restrict__ Lst, size_t l_stride, double *__restrict__ xT) {
  extern __shared__ __align__(16) double mf_smem[];
  constexpr int NW = NT / 32;
  const int f = M.level_fronts[lf_off + blockIdx.x];
  const MfFront F = M.fronts[f];
  const int cell = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s8 = F.s8, u8 = F.u8, kr = M.kr, NP = M.NP;
  double *L11 = mf_smem;                       // [s8][s8]
  double *ts = L11 + (size_t)s8 * s8;          // [s8][kr]  z, then t, then x
  double *xu = ts + (size_t)s8 * kr;           // [u8][kr]  x of the reached unknowns
  const double *Lp = Lst + (size_t)cell * l_stride + F.l_off;   // [m][s8]
  double *xc = xT + (size_t)cell * k * NP;
  for (int j = warp; j < kr; j += NW) {
    if (j < k) {
      for (int r = lane; r < u8; r += 32) {
        const int p = M.front_idx[F.idx_off + s8 + r];
        xu[r * kr + j] = p >= 0 ? xc[(size_t)j * NP + p] : 0.0;
      }
    } else {
      for (int r = lane; r < u8; r += 32) xu[r * kr + j] = 0.0;
    }
    for (int c = lane; c < s8; c += 32) ts[c * kr + j] = Lp[(size_t)(s8 + u8 + j) * s8 + c];
  }
  for (int idx = tid; idx < s8 * s8; idx += NT) L11[idx] = Lp[idx];
  __syncthreads();
  // t = z - L21^T x_reached.  L21 streams through shared memory in chunks of CH rows (one coalesced copy with many
  // loads in flight; reading it straight from global memory left every thread waiting on its own dependent loads);
  // a thread owns one column c and 4 right-hand sides.
  if (u8 > 0) {
    const double *L21 = Lp + (size_t)s8 * s8;
    double *stage = xu + (size_t)u8 * kr;
    const int CH = mf_bwd_chunk(s8, u8);
    const int items = s8 * (kr / 4);
    for (int r0 = 0; r0 < u8; r0 += CH) {
      const int nr = min(CH, u8 - r0);
      {
        const double2 *src = reinterpret_cast<const double2 *>(L21 + (size_t)r0 * s8);
        double2 *dst = reinterpret_cast<double2 *>(stage);
        const int n2 = (nr * s8) >> 1;
        for (int i = tid; i < n2; i += NT) dst[i] = src[i];
      }
      __syncthreads();
      for (int o = tid; o < items; o += NT) {
        const int jg = o / s8, c = o - jg * s8;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        const double *xr = xu + (size_t)r0 * kr + jg * 4;
#pragma unroll 4
        for (int r = 0; r < nr; ++r) {
          const double l = stage[r * s8 + c];
          const double2 x0 = *reinterpret_cast<const double2 *>(xr + r * kr);
          const double2 x1 = *reinterpret_cast<const double2 *>(xr + r * kr + 2);
          acc[0] = fma(l, x0.x, acc[0]); acc[1] = fma(l, x0.y, acc[1]);
          acc[2] = fma(l, x1.x, acc[2]); acc[3] = fma(l, x1.y, acc[3]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) ts[c * kr + jg * 4 + i] -= acc[i];
      }
      __syncthreads();
    }
  }
  // L11^T x = t, 8 unknowns at a time, last tile first
  for (int p = s8 / 8 - 1; p >= 0; --p) {
    const int c0 = p * 8;
    if (tid < kr) {
      double x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = ts[(c0 + i) * kr + tid];
#pragma unroll
      for (int i = 6; i >= 0; --i)
#pragma unroll
        for (int i2 = i + 1; i2 < 8; ++i2) x[i] = fma(-L11[(size_t)(c0 + i2) * s8 + c0 + i], x[i2], x[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) ts[(c0 + i) * kr + tid] = x[i];
    }
    __syncthreads();
    if (c0 > 0) {
      for (int j = warp; j < kr; j += NW)
        for (int cp = lane; cp < c0; cp += 32) {
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc = fma(L11[(size_t)(c0 + i) * s8 + cp], ts[(c0 + i) * kr + j], acc);
          ts[cp * kr + j] -= acc;
        }
      __syncthreads();
    }
  }
  for (int j = warp; j < k; j += NW)
    for (int c = lane; c < s8; c += 32) xc[(size_t)j * NP + F.own_base + c] = ts[c * kr + j];
}

}  // namespace
}  // namespace msfec
