// Superseded kernel variants of the round-1 banded direct solver, kept for the record (not compiled into the product):
//   k_direct_diag / k_direct_panel / k_direct_update<128,32>   per-panel factorisation of a chunk's diagonal region
//                                                              (7 launches per chunk, bound by the shuffle unit)
//   k_direct_region                                            one-warp FMA-only region kernel (bound by the shared-memory
//                                                              return path)
// Measurements against the shipped k_direct_region_mma / k_direct_trsm<32,4>: profiles/r01_direct_path_v9.md.
// They were selectable through MSFEC_DIRECT_FUSED_REGION = 0 | 1 in round 1; the switches are gone.
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
// factor (column-major, pivots on the diagonal) to diagL, the pivots to dvec and -- for the DMMA triangular
// solves of k_direct_trsm -- the inverse V = L^-1 of the unit-lower factor (column-major) to vinv.
// grid (ceil(cells/4)), block 128
__global__ void __launch_bounds__(128)
k_direct_diag(const double *__restrict__ band, size_t band_stride, long long col_off, int ld, int j0, int pglob, int NP,
              int n_cells, double *__restrict__ diagL, double *__restrict__ dvec, double *__restrict__ vinv,
              int *__restrict__ bad) {
  __shared__ double Ls[4][kDP][kDP + 1];
  const int w = threadIdx.x >> 5;
  const int cell = blockIdx.x * 4 + w, lane = threadIdx.x & 31;
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
  if (vinv == nullptr) return;
  // V = L^-1: lane j solves L v = e_j (column j of V) by forward substitution; L(i, k) is a broadcast read
#pragma unroll
  for (int p = 0; p < kDP; ++p) Ls[w][lane][p] = (p < lane) ? a[p] : 0.0;
  __syncwarp();
  double v[kDP];
#pragma unroll
  for (int i = 0; i < kDP; ++i) {
    double t = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < i; ++k) t = fma(-Ls[w][i][k], v[k], t);
    v[i] = t;
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < kDP; ++i) Ls[w][i][lane] = v[i];          // Ls[i][j] = V(i, j)
  __syncwarp();
  double *vo = vinv + ((size_t)cell * NP + pglob) * kDP;
#pragma unroll
  for (int k = 0; k < kDP; ++k) vo[(size_t)k * kDP + lane] = Ls[w][lane][k];   // column-major: V(n = lane, k)
}


// ---- fused factorisation of a chunk's diagonal region --------------------------------------
// One launch per chunk instead of (diag, panel, strip update) per 32-column panel: ONE WARP per cell factors the
// whole (32 np) x (32 np) lower-triangular region right-looking, lane = row of the current 32 x 32 block:
//   for p:  LDL^T of block (p,p) in registers (the pivot column is broadcast through shared memory: 16 LDS.128 per
//           step instead of 62 SHFL; the per-panel kernel was bound by the shuffle unit), pivots -> dvec, V_p = L^-1 -> vinv;
//           blocks (q,p), q > p:  X = A L_pp^-T by forward substitution, L = X D^-1 back to the band;
//           blocks (q',q), p < q <= q':  C -= X_q'p L_qp^T  (L_qp rows broadcast from shared memory).
// No block-wide barriers (warps are independent), FP64 FMA pipe only.  The band keeps L in the sub-diagonal blocks
// (what k_direct_trsm / k_direct_back_diag read); the diagonal blocks themselves are never read again.
// grid (ceil(cells/kRegionWarps)), block 32 kRegionWarps
constexpr int kRLd = kDP + 2;   // shared-memory row stride: even, so that (k, k+1) pairs load as 16 bytes
constexpr int kRegionWarps = 2; // cells per CTA (17.8 KB of static shared memory per warp)
__global__ void __launch_bounds__(32 * kRegionWarps, 6)
k_direct_region(double *__restrict__ band, size_t band_stride, long long col_off, int ld, int j0, int np, int pglob0, int NP,
                int n_cells, double *__restrict__ dvec, double *__restrict__ vinv, int *__restrict__ bad) {
  __shared__ __align__(16) double Ls_[kRegionWarps][kDP][kRLd];   // rows of L_pp (strictly lower part, zeros elsewhere)
  __shared__ __align__(16) double Lq_[kRegionWarps][kDP][kRLd];   // rows of L_qp for the in-region updates; V staging
  __shared__ __align__(16) double colb_[kRegionWarps][2][kDP];    // pivot-column broadcast, double buffered
  __shared__ __align__(16) double dsm_[kRegionWarps][2][kDP];     // [0]: 1/d, [1]: d of the current panel
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cell = blockIdx.x * kRegionWarps + w;
  if (cell >= n_cells) return;                          // warp-uniform
  double (*Ls)[kRLd] = Ls_[w];
  double (*Lq)[kRLd] = Lq_[w];
  double *P = band + (size_t)cell * band_stride + col_off;
  for (int p = 0; p < np; ++p) {
    const int jp = j0 + p * kDP;                        // first column / row of block (p,p) in the block column
    // ---- LDL^T of the diagonal block -----------------------------------------------------------
    double a[kDP];
#pragma unroll
    for (int k = 0; k < kDP; ++k) a[k] = (k <= lane) ? P[(size_t)(jp + k) * ld + jp + lane] : 0.0;
    double my_d = 0.0;
#pragma unroll
    for (int c = 0; c < kDP; ++c) {
      double *cb = colb_[w][c & 1];
      cb[lane] = a[c];                                  // unscaled column c, one entry per lane (row)
      __syncwarp();
      const double d = cb[c];
      const double dinv = 1.0 / d;
      const double l = a[c] * dinv;
      if (lane == c) {
        my_d = d;
        if (!(fabs(d) > 1e-300) || !isfinite(d)) atomicExch(bad, 1);
      }
#pragma unroll
      for (int jj = (c + 1) / 2; jj < kDP / 2; ++jj) {
        const double2 v = reinterpret_cast<const double2 *>(cb)[jj];
        if (2 * jj > c && 2 * jj <= lane) a[2 * jj] = fma(-l, v.x, a[2 * jj]);
        if (2 * jj + 1 > c && 2 * jj + 1 <= lane) a[2 * jj + 1] = fma(-l, v.y, a[2 * jj + 1]);
      }
      if (lane > c) a[c] = l;
    }
    dvec[(size_t)cell * NP + pglob0 + p * kDP + lane] = my_d;
    dsm_[w][0][lane] = 1.0 / my_d;
    dsm_[w][1][lane] = my_d;
#pragma unroll
    for (int k = 0; k < kDP; ++k) Ls[lane][k] = (k < lane) ? a[k] : 0.0;
    __syncwarp();
    // ---- V_p = L_pp^-1: lane j solves L v = e_j (column j of V) ----------------------------------
    {
      double v[kDP];
#pragma unroll
      for (int i = 0; i < kDP; ++i) {
        double t = (i == lane) ? 1.0 : 0.0, t2 = 0.0;      // two partial sums: shorter dependent FMA chains
#pragma unroll
        for (int kk = 0; kk < i / 2; ++kk) {
          const double2 lv = *reinterpret_cast<const double2 *>(&Ls[i][2 * kk]);
          t = fma(-lv.x, v[2 * kk], t);
          t2 = fma(-lv.y, v[2 * kk + 1], t2);
        }
        if (i & 1) t = fma(-Ls[i][i - 1], v[i - 1], t);
        v[i] = t + t2;
        asm volatile("" ::: "memory");
      }
#pragma unroll
      for (int i = 0; i < kDP; ++i) Lq[i][lane] = v[i];                 // Lq[i][j] = V(i, j)
      __syncwarp();
      double *vo = vinv + ((size_t)cell * NP + pglob0 + p * kDP) * kDP;
#pragma unroll
      for (int k = 0; k < kDP; ++k) vo[(size_t)k * kDP + lane] = Lq[lane][k];   // column-major: V(n = lane, k)
      __syncwarp();
    }
    if (p + 1 == np) break;
    // ---- blocks below: X = A L_pp^-T, L = X D^-1 ---------------------------------------------------
    for (int q = p + 1; q < np; ++q) {
      double *B = P + (size_t)jp * ld + j0 + q * kDP + lane;             // row `lane` of block (q, p)
      double y[kDP];
#pragma unroll
      for (int k = 0; k < kDP; ++k) y[k] = B[(size_t)k * ld];
#pragma unroll
      for (int k = 1; k < kDP; ++k) {
        double t = y[k], t2 = 0.0;
#pragma unroll
        for (int mm = 0; mm < k / 2; ++mm) {
          const double2 lv = *reinterpret_cast<const double2 *>(&Ls[k][2 * mm]);
          t = fma(-y[2 * mm], lv.x, t);
          t2 = fma(-y[2 * mm + 1], lv.y, t2);
        }
        if (k & 1) t = fma(-y[k - 1], Ls[k][k - 1], t);
        y[k] = t + t2;
        asm volatile("" ::: "memory");
      }
#pragma unroll
      for (int k = 0; k < kDP; ++k) B[(size_t)k * ld] = y[k] * dsm_[w][0][k];
    }
    __syncwarp();                                        // the L blocks written above are re-read by other lanes
    // ---- right-looking update of the rest of the region ---------------------------------------------
    for (int q = p + 1; q < np; ++q) {
      {
        const double *B = P + (size_t)jp * ld + j0 + q * kDP + lane;     // row `lane` of L_qp
#pragma unroll
        for (int k = 0; k < kDP; ++k) Lq[lane][k] = B[(size_t)k * ld];
      }
      __syncwarp();
      for (int q2 = q; q2 < np; ++q2) {
        const double *B = P + (size_t)jp * ld + j0 + q2 * kDP + lane;    // row `lane` of L_q2p
        double x[kDP];
#pragma unroll
        for (int k = 0; k < kDP; ++k) x[k] = B[(size_t)k * ld] * dsm_[w][1][k];     // X = L D
        double *C = P + (size_t)(j0 + q * kDP) * ld + j0 + q2 * kDP + lane;         // row `lane` of block (q2, q)
        // 4 target columns at a time: 4 independent FMA chains, each L_qp pair is one 16-byte broadcast load
#pragma unroll 2
        for (int jb = 0; jb < kDP; jb += 4) {
          double t[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) t[u] = C[(size_t)(jb + u) * ld];
#pragma unroll
          for (int kk = 0; kk < kDP / 2; ++kk) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const double2 lv = *reinterpret_cast<const double2 *>(&Lq[jb + u][2 * kk]);
              t[u] = fma(-x[2 * kk], lv.x, t[u]);
              t[u] = fma(-x[2 * kk + 1], lv.y, t[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) C[(size_t)(jb + u) * ld] = t[u];
        }
      }
      __syncwarp();                                      // Lq is restaged for the next q
    }
    __syncwarp();                                        // updated blocks are re-read with other lane mappings
  }
}


// Row-parallel triangular solve of the virtual rows v in [j0+32, ld) (rest of slab s, slab s+1 and the rhs
// rows: forward substitution is fused into the factorisation) against the factored diagonal block.
// grid (ceil(nrows/128), cells), block 128
__global__ void __launch_bounds__(128)
k_direct_panel(double *__restrict__ band, size_t band_stride, long long col_off, int ld, int row_hi, int j0, int pglob,
               int NP, const double *__restrict__ diagL, const double *__restrict__ dvec, double *__restrict__ ybuf,
               int slot, int ldy) {
  __shared__ double Ld[kDP][kDP + 1];
  __shared__ double dinv[kDP];
  const int cell = blockIdx.y, tid = threadIdx.x;
  double *P = band + (size_t)cell * band_stride + col_off;
  const int v = j0 + kDP + blockIdx.x * 128 + tid;
  double y[kDP];
  if (v < row_hi) {
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
  if (v >= row_hi) return;
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
k_direct_update(double *__restrict__ band, size_t band_stride, DirectPlanDev D, int s, int jsrc, int nq, int yslot0,
                int vc_lo, int vc_hi, int row_hi, const double *__restrict__ ybuf, int ldy) {
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
  const int T = (row_hi - r_first + TM - 1) / TM;   // rows >= row_hi are left alone (diagonal-region strips)
  const int Z = gridDim.y;
  if ((int)blockIdx.y >= T) return;
  // column operands: -y = -(L D) of the nq source panels, written by k_direct_panel into the window scratch
  const double *Yc = ybuf + ((size_t)cell * kMaxWindow + yslot0) * kDP * ldy;
  for (int q = 0; q < nq; ++q) stage_block<TN, NT>(Lc + (size_t)q * kDP * LDC, Yc + (size_t)q * kDP * ldy, ldy, row_hi, 0, cbase, tid);
  stage_block<TM, NT>(Lr, P, ld, row_hi, jsrc, r_first + blockIdx.y * TM, tid);
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
    const bool active = col_ok && vr0 >= vc0 && vr0 < row_hi;
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
      if (q + 1 < nq) stage_block<TM, NT>(Lr + (buf ^ 1) * kDP * LDR, P, ld, row_hi, jsrc + (q + 1) * kDP, rbase, tid);
      else if (ti + Z < T) stage_block<TM, NT>(Lr + (buf ^ 1) * kDP * LDR, P, ld, row_hi, jsrc, rbase + Z * TM, tid);
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


