// Batched direct solver kernels: block-sparse LDL^T without pivoting (see DirectPlan in topology.h),
// replacing NedRTBasis::solve_direct (reference ned_rt_basis.cc:579-634, UMFPACK re-factorised for
// each of the k right-hand sides) for all cells and all k rhs at once.
//
// Per cell the "band" holds, for every block column s, a column-major panel of ld_s rows
// [block s | reached blocks | 32 rhs rows] x bs_s columns (the rhs ride along as rows, so the forward
// substitution is part of the factorisation).  A block column is processed in CHUNKS of <= 3 panels of
// 32 columns (engine.cu: Engine::solve_direct_batch):
//   k_direct_fill_fused  band := matrix entries / 0 in one coalesced pass
//   k_direct_region_mma  LDL^T of the chunk's diagonal region (one warp per cell), V = L^-1 of every 32 x 32 pivot block
//   k_direct_trsm        all rows below the region in one pass: X_p = (A_p - sum X_q L(p,q)^T) V_p^T on
//                        mma.sync.m8n8k4.f64, writes L = X D^-1 and -X (window scratch)
//   k_direct_update_s    C -= L D L^T for everything behind the chunk (the dominant kernel): one 64x64 tile
//                        per CTA, 32x32 per warp, operands through a 4-stage cp.async ring
//   k_direct_back_gemm / k_direct_back_diag   L^T x = z per block column, last to first
// Included by engine.cu (and by profiles/microbench_update.cu).
#pragma once

namespace msfec {
namespace {

constexpr int kDP = 32;   // panel width (DirectPlan::kPanel)
constexpr int kMaxWindow = 6;   // max panels per chunk (slots of the window scratch)

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

// Fused zero + fill: every band entry is written exactly once (value or 0) with 16-byte stores, driven by
// a per-entry code table shared by all cells (0: zero, 1: per-cell slot reference, 2: cell-independent value
// x kscale, 3: constant, 4: right-hand side (row, j), 7: never read, not written).  Replaces memset + the four scatter kernels above
// (whose 8-byte scattered writes cost a read-modify-write per sector).
// grid (ceil(n_entries / 512), ceil(cells / kFillCells)), block 256; band entries in pairs
__device__ __forceinline__ double fill_value(int code, const double *__restrict__ vals_cell, const double *__restrict__ sval,
                                             double kscale, const double *__restrict__ kval, const double *__restrict__ b_cell,
                                             int k) {
  const int t = code & 7, a = code >> 3;
  if (t == 0) return 0.0;
  if (t == 1) { const double v = vals_cell[(size_t)(a >> 1) * kLanes]; return (a & 1) ? -v : v; }
  if (t == 2) return sval[a] * kscale;
  if (t == 3) return kval[a];
  return b_cell[((size_t)(a >> 5) * k + (a & 31)) * kLanes];
}

constexpr int kFillCells = 4;    // cells per CTA (16 cells per CTA with a zero fast path measured 8.78 vs 8.59 ms per 4096 cells)
__global__ void __launch_bounds__(256)
k_direct_fill_fused(const int2 *__restrict__ code, long long n_pairs, const double *__restrict__ vals, int n_slots,
                    const double *__restrict__ sval, double kscale, const double *__restrict__ kval,
                    const double *__restrict__ b, int NI, int k, int cell_lo, int n_cells, double *__restrict__ band,
                    size_t band_stride) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= n_pairs) return;
  const int2 c = code[e];
  if (c.x == 7 && c.y == 7) return;                        // never-read upper block: leave unwritten
#pragma unroll
  for (int u = 0; u < kFillCells; ++u) {
    const int ci = blockIdx.y * kFillCells + u;
    if (ci >= n_cells) break;
    double2 v = make_double2(0.0, 0.0);
    if (c.x | c.y) {
      const int cell = cell_lo + ci, g = cell / kLanes, lane = cell % kLanes;
      const double *vc = vals + (size_t)g * n_slots * kLanes + lane;
      const double *bc = b + (size_t)g * NI * k * kLanes + lane;
      v.x = fill_value(c.x, vc, sval, kscale, kval, bc, k);
      v.y = fill_value(c.y, vc, sval, kscale, kval, bc, k);
    }
    *reinterpret_cast<double2 *>(band + (size_t)ci * band_stride + 2 * e) = v;
  }
}

// ---- factorisation of a chunk's diagonal region ---------------------------------------------
// One launch per chunk, block products on the FP64 tensor cores (an FMA-only version was bound by the shared-memory
// return path, per-panel launches by the shuffle unit: profiles/r01_superseded_kernels.cuh).  One warp per CTA; per panel p:
//   LDL^T of block (p,p) in registers (pivot column broadcast through shared memory), V_p = L_pp^-1 (lane = column);
//   X_qp = A_qp V_p^T for the blocks below as one 32x32x32 DMMA product each (zero k-steps of the triangular V_p
//   skipped), L = X D^-1 to the band, X kept in shared memory;
//   C(q2,q) -= X_q2p L_qp^T on DMMA with L = X D^-1 formed in the fragment load; diagonal targets skip the 8x8 tiles
//   above the diagonal.
// MMA m = band column, n = band row, so accumulator pairs are consecutive band rows (16-byte global accesses).
// Shared memory: max(np, 2) blocks of 32 x 36 doubles (V_p + one per block row below; the first of those holds the
// rows of L_pp while V_p is computed).   grid (cells), block 32, dynamic shared memory region_mma_smem(np)
constexpr int kRMLd = kDP + 4;                     // row stride 36 = 4 mod 16: conflict-free fragment loads
constexpr size_t region_mma_smem(int np, bool keep_x) { return ((size_t)(keep_x ? (np < 2 ? 2 : np) : 3) * kDP * kRMLd + 4 * kDP) * sizeof(double); }

__device__ __forceinline__ double rcp_newton(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));   // ~2^-20; three Newton steps give full double precision
  r = fma(fma(-d, r, 1.0), r, r);
  r = fma(fma(-d, r, 1.0), r, r);
  r = fma(fma(-d, r, 1.0), r, r);
  return r;
}

// KEEP_X = true: the X blocks of all block rows below the panel stay in shared memory (np blocks of 9 KB: 7 CTAs/SM at
// np = 3 but only 4 at np = 5).  KEEP_X = false: three blocks for any np -- the updates re-stage L_qp / L_q2p from the
// band (L2 hits) and form X = L D in the fragment load, which keeps 7 warps per SM for the 5-panel chunks.
template <bool KEEP_X>
__global__ void __launch_bounds__(32, 8)
k_direct_region_mma(double *__restrict__ band, size_t band_stride, long long col_off, int ld, int j0, int np, int pglob0,
                    int NP, double *__restrict__ dvec, double *__restrict__ vinv, int *__restrict__ bad) {
  extern __shared__ __align__(16) double rm_smem[];
  double *Vs = rm_smem;                                   // [32][36]  V_p(i, j)
  double *S1 = rm_smem + kDP * kRMLd;                     // [np-1][32][36]  L_pp rows, then X blocks of rows q > p
  double *colb = rm_smem + (size_t)(KEEP_X ? (np < 2 ? 2 : np) : 3) * kDP * kRMLd;      // [2][32]
  double *dsm = colb + 2 * kDP;                           // [2][32]  1/d and d of the current panel
  const int cell = blockIdx.x, lane = threadIdx.x;
  const int fr = lane >> 2, fk = lane & 3;
  double *P = band + (size_t)cell * band_stride + col_off;
  for (int p = 0; p < np; ++p) {
    const int jp = j0 + p * kDP;
    // ---- LDL^T of the diagonal block, lane = row ---------------------------------------------------
    double my_d = 0.0;
    {
      double a[kDP];
#pragma unroll
      for (int k = 0; k < kDP; ++k) a[k] = (k <= lane) ? P[(size_t)(jp + k) * ld + jp + lane] : 0.0;
#pragma unroll
      for (int c = 0; c < kDP; ++c) {
        double *cb = colb + (c & 1) * kDP;
        cb[lane] = a[c];
        __syncwarp();
        const double d = cb[c];
        const double l = a[c] * rcp_newton(d);
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
      dsm[lane] = 1.0 / my_d;
      dsm[kDP + lane] = my_d;
#pragma unroll
      for (int k = 0; k < kDP; ++k) S1[lane * kRMLd + k] = (k < lane) ? a[k] : 0.0;
    }
    __syncwarp();
    // ---- V_p = L_pp^-1: lane j solves L v = e_j ------------------------------------------------------
    {
      double v[kDP];
#pragma unroll
      for (int i = 0; i < kDP; ++i) {
        double t = (i == lane) ? 1.0 : 0.0, t2 = 0.0;
#pragma unroll
        for (int kk = 0; kk < i / 2; ++kk) {
          const double2 lv = *reinterpret_cast<const double2 *>(&S1[i * kRMLd + 2 * kk]);
          t = fma(-lv.x, v[2 * kk], t);
          t2 = fma(-lv.y, v[2 * kk + 1], t2);
        }
        if (i & 1) t = fma(-S1[i * kRMLd + i - 1], v[i - 1], t);
        v[i] = t + t2;
        asm volatile("" ::: "memory");
      }
#pragma unroll
      for (int i = 0; i < kDP; ++i) Vs[i * kRMLd + lane] = v[i];
    }
    __syncwarp();
    {
      double *vo = vinv + ((size_t)cell * NP + pglob0 + p * kDP) * kDP;
#pragma unroll
      for (int k = 0; k < kDP; ++k) vo[(size_t)k * kDP + lane] = Vs[lane * kRMLd + k];   // column-major V(n = lane, k)
    }
    if (p + 1 == np) break;
    double dinv_m[4], ndinv_k[kDP / 4];
#pragma unroll
    for (int t = 0; t < 4; ++t) dinv_m[t] = dsm[t * 8 + fr];
#pragma unroll
    for (int ks = 0; ks < kDP / 4; ++ks) ndinv_k[ks] = KEEP_X ? -dsm[ks * 4 + fk] : -dsm[kDP + ks * 4 + fk];   // -1/d_k | -d_k
    // ---- blocks below: X_qp = A_qp V_p^T ----------------------------------------------------------------
    for (int q = p + 1; q < np; ++q) {
      double *Xq = S1 + (KEEP_X ? (size_t)(q - p - 1) * kDP * kRMLd : 0);
      {
        const double *B = P + (size_t)jp * ld + j0 + q * kDP + lane;       // row `lane` of A_qp
#pragma unroll
        for (int k = 0; k < kDP; ++k) Xq[lane * kRMLd + k] = B[(size_t)k * ld];
      }
      __syncwarp();
      double acc[4][4][2];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < kDP / 4; ++ks) {
        double bf[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) bf[t] = Xq[(t * 8 + fr) * kRMLd + ks * 4 + fk];        // B[k = m'][n = i] = A(i, m')
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          if (ks >= 2 * mt + 2) continue;                                                   // V(k, m') = 0 for m' > k
          const double af = Vs[(mt * 8 + fr) * kRMLd + ks * 4 + fk];                        // A[m = k][k = m'] = V(k, m')
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], af, bf[nt]);
        }
      }
      __syncwarp();                                                          // A_qp fully consumed: overwrite with X_qp
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int k = mt * 8 + fr, i = nt * 8 + fk * 2;
          if (KEEP_X) {
            Xq[i * kRMLd + k] = acc[mt][nt][0];
            Xq[(i + 1) * kRMLd + k] = acc[mt][nt][1];
          }
          *reinterpret_cast<double2 *>(P + (size_t)(jp + k) * ld + j0 + q * kDP + i) =
              make_double2(acc[mt][nt][0] * dinv_m[mt], acc[mt][nt][1] * dinv_m[mt]);
        }
    }
    __syncwarp();
    // ---- right-looking update of the rest of the region: C(q2, q) -= X_q2p L_qp^T ----------------------------
    for (int q = p + 1; q < np; ++q) {
      const double *Lq = S1 + (KEEP_X ? (size_t)(q - p - 1) * kDP * kRMLd : 0);   // X_qp (KEEP_X) or L_qp
      if (!KEEP_X) {
        __syncwarp();                                                          // previous readers of both slots are done
        const double *B = P + (size_t)jp * ld + j0 + q * kDP + lane;           // row `lane` of L_qp
#pragma unroll
        for (int k = 0; k < kDP; ++k) S1[lane * kRMLd + k] = B[(size_t)k * ld];
        __syncwarp();
      }
      for (int q2 = q; q2 < np; ++q2) {
        const double *Xq2 = S1 + (KEEP_X ? (size_t)(q2 - p - 1) * kDP * kRMLd : (q2 == q ? 0 : (size_t)kDP * kRMLd));
        if (!KEEP_X && q2 != q) {
          __syncwarp();
          double *S2 = S1 + (size_t)kDP * kRMLd;
          const double *B = P + (size_t)jp * ld + j0 + q2 * kDP + lane;        // row `lane` of L_q2p
#pragma unroll
          for (int k = 0; k < kDP; ++k) S2[lane * kRMLd + k] = B[(size_t)k * ld];
          __syncwarp();
        }
        double *C = P + (size_t)(j0 + q * kDP + fr) * ld + j0 + q2 * kDP + fk * 2;
        const bool dg = q2 == q;
        double acc[4][4][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            if (dg && nt < mt) continue;
            const double2 v = *reinterpret_cast<const double2 *>(C + (size_t)(mt * 8) * ld + nt * 8);
            acc[mt][nt][0] = v.x; acc[mt][nt][1] = v.y;
          }
#pragma unroll
        for (int ks = 0; ks < kDP / 4; ++ks) {
          double af[4], bf[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (KEEP_X) {
              af[t] = Lq[(t * 8 + fr) * kRMLd + ks * 4 + fk] * ndinv_k[ks];                 // A[m = j][k] = -L(j, k) = -X(j, k) / d_k
              bf[t] = Xq2[(t * 8 + fr) * kRMLd + ks * 4 + fk];                              // B[k][n = i] = X(i, k)
            } else {
              af[t] = Lq[(t * 8 + fr) * kRMLd + ks * 4 + fk];                               // A[m = j][k] = L(j, k)
              bf[t] = Xq2[(t * 8 + fr) * kRMLd + ks * 4 + fk] * ndinv_k[ks];                // B[k][n = i] = -X(i, k) = -L(i, k) d_k
            }
          }
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              if (dg && nt < mt) continue;
              dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
            }
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            if (dg && nt < mt) continue;
            *reinterpret_cast<double2 *>(C + (size_t)(mt * 8) * ld + nt * 8) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
          }
      }
    }
    __syncwarp();                                          // next panel re-reads band and shared memory with other lane mappings
  }
}

// ---- trailing update on FP64 tensor cores ----------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Device copy of the block structure (DirectPlan): all lookups below are warp-uniform.
struct DirectPlanDev {
  int n_slabs, NP;
  const int *bs, *slab_off, *ld, *front_rows;
  const long long *col_off;
  const int *chunk_off, *chunk_blk, *chunk_local, *front_pos;
};

// ---- trailing update, streaming version -------------------------------------------------
// C(vr, vc) += sum_k L(vr, k) * Ys(vc, k) for everything behind a chunk, with
// Ys = -(L D) from the window scratch (slots yslot0 .. yslot0+nq-1) and L from columns jsrc .. jsrc+32nq-1 of
// block column s.  One 64x64 C tile per CTA (4 warps of 32x32), both operands streamed through a 4-stage
// cp.async ring of K = 16 slices, so any K fits (the whole block column is applied to the reached blocks in
// ONE pass: every C entry behind a block column is read and written once per block column, not once per
// window).  The MMA m-dimension is the band COLUMN, so a thread's two accumulator entries are two
// consecutive band rows: C moves with 16-byte accesses.  Shared-memory row stride 68 (= 4 mod 16) makes the
// fragment loads conflict-free.
// grid (tiles of the trapezoid vc in [vc_lo, vc_hi), vr in [vc, ld), cells); block 128.
template <int TM, int TN, int KC, int ST>
constexpr size_t update_s_smem() { return (size_t)ST * KC * (TM + 4 + TN + 4) * sizeof(double); }

// number of tiles of the trapezoid (host side; the kernel decodes the same enumeration)
template <int TM, int TN>
inline int update_s_tiles(int ld, int vc_lo, int c_hi) {
  int n = 0;
  for (int cbase = vc_lo; cbase < c_hi; cbase += TN) {
    const int r_first = cbase - (cbase - vc_lo) % TM;
    n += (ld - r_first + TM - 1) / TM;
  }
  return n;
}

template <int TM, int TN, int KC, int ST, int MINB>
__global__ void __launch_bounds__((TM / 32) * (TN / 32) * 32, MINB)
k_direct_update_s(double *__restrict__ band, size_t band_stride, DirectPlanDev D, int s, int jsrc, int nq, int yslot0,
                  int vc_lo, int vc_hi, const double *__restrict__ ybuf, int ldy) {
  constexpr int NT = (TM / 32) * (TN / 32) * 32, LDL = TM + 4, LDY = TN + 4, STAGE = KC * (LDL + LDY);
  constexpr int WC = TN / 32;
  extern __shared__ __align__(16) double upd_smem[];
  const int cell = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ld = D.ld[s], front_rows = D.front_rows[s];
  // tile decode: column tile tc owns the row tiles from its own first row on (TM-aligned relative to vc_lo)
  int cbase = vc_lo, rbase;
  {
    int t = blockIdx.x;
    for (;; cbase += TN) {
      const int r_first = cbase - (cbase - vc_lo) % TM;
      const int T = (ld - r_first + TM - 1) / TM;
      if (t < T) { rbase = r_first + t * TM; break; }
      t -= T;
    }
  }
  const long long col_off = D.col_off[s];
  const int choff = D.chunk_off[s];
  double *cb = band + (size_t)cell * band_stride;
  const double *Lsrc = cb + col_off + (size_t)jsrc * ld;
  const double *Ysrc = ybuf + ((size_t)cell * kMaxWindow + yslot0) * kDP * ldy;
  const int nk = nq * (kDP / KC);
  auto load_stage = [&](int kt) {
    double *dst = upd_smem + (size_t)(kt % ST) * STAGE;
#pragma unroll
    for (int t = 0; t < KC * TM / 2 / NT; ++t) {
      const int c = tid + t * NT;
      const int k = c / (TM / 2), i = (c % (TM / 2)) * 2;
      double *d = dst + k * LDL + i;
      if (rbase + i < ld) cp_async16(d, Lsrc + (size_t)(kt * KC + k) * ld + rbase + i);
      else { d[0] = 0.0; d[1] = 0.0; }
    }
#pragma unroll
    for (int t = 0; t < KC * TN / 2 / NT; ++t) {
      const int c = tid + t * NT;
      const int k = c / (TN / 2), i = (c % (TN / 2)) * 2;
      double *d = dst + KC * LDL + k * LDY + i;
      if (cbase + i < ld) cp_async16(d, Ysrc + (size_t)(kt * KC + k) * ldy + cbase + i);
      else { d[0] = 0.0; d[1] = 0.0; }
    }
  };
#pragma unroll
  for (int st = 0; st < ST - 1; ++st) {
    if (st < nk) load_stage(st);
    cp_async_commit();
  }
  const int wr = warp / WC, wc = warp % WC;
  const int fr = lane >> 2, fk = lane & 3;
  const int vc0 = cbase + wc * 32, vr0 = rbase + wr * 32;
  const bool active = vc0 < vc_hi && vc0 < front_rows && vr0 >= vc0 && vr0 < ld;
  const bool diagw = vr0 == vc0;
  double *cdst = cb;
  int ldc = ld;
  if (active) {
    const int cblk = D.chunk_blk[choff + (vc0 >> 5)];
    int roff = vr0;
    if (cblk == s) cdst = cb + col_off + (size_t)vc0 * ld;
    else {
      ldc = D.ld[cblk];
      cdst = cb + D.col_off[cblk] + (size_t)D.chunk_local[choff + (vc0 >> 5)] * ldc;
      const int rb = D.chunk_blk[choff + (vr0 >> 5)];
      roff = rb < 0 ? D.front_rows[cblk] + (vr0 - front_rows)
                    : D.front_pos[cblk * D.n_slabs + rb] + D.chunk_local[choff + (vr0 >> 5)];
    }
    cdst += (size_t)fr * ldc + roff + fk * 2;
  }
  double acc[4][4][2];
  if (active) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const double2 v = *reinterpret_cast<const double2 *>(cdst + (size_t)(mt * 8) * ldc + nt * 8);
        acc[mt][nt][0] = v.x; acc[mt][nt][1] = v.y;
      }
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<ST - 2>();
    __syncthreads();
    if (kt + ST - 1 < nk) load_stage(kt + ST - 1);
    cp_async_commit();
    if (active) {
      const double *Ls = upd_smem + (size_t)(kt % ST) * STAGE + wr * 32 + fr;
      const double *Ys = upd_smem + (size_t)(kt % ST) * STAGE + KC * LDL + wc * 32 + fr;
#pragma unroll
      for (int ks = 0; ks < KC / 4; ++ks) {
        const int kk = ks * 4 + fk;
        double af[4], bf[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          af[t] = Ys[kk * LDY + t * 8];
          bf[t] = Ls[kk * LDL + t * 8];
        }
        if (diagw) {
          // warp tile on the diagonal: 8x8 sub-tiles strictly above it (row tile nt < column tile mt) are never read
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = mt; nt < 4; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        } else {
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
      }
    }
  }
  cp_async_wait<0>();
  if (active) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        *reinterpret_cast<double2 *>(cdst + (size_t)(mt * 8) * ldc + nt * 8) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
  }
}


// ---- rows below a chunk: triangular solve on FP64 tensor cores ---------------------------
// A chunk is np <= 6 consecutive panels (W = 32 np columns, first column jc0) of block column s whose W x W
// diagonal region is already factored (sub-diagonal blocks L(p,q) in the band, V_p = L(p,p)^-1 in vinv, pivots
// in dvec).  For the rows below it, A_rows = X L_dd^T with X = L_rows D, solved block-wise
//   X_p = (A_p - sum_{q<p} X_q L(p,q)^T) V_p^T
// entirely with mma.sync.m8n8k4.f64: the MMA m-dimension is the panel column, n the row, so accumulator pairs
// are consecutive band rows (16-byte stores).  One CTA = 32 rows x W columns held in shared memory as
// [column][row]; the 32x32 operand blocks L(p,q), V_p stream through a 3-stage cp.async ring.  Writes
// L = X D^-1 in place and -X to the window scratch (slot p) for the chunk update.  Every band entry of these
// rows is read once and written once (the 32-wide panel kernels moved it ~5 times).
// grid ((ld - row_lo) / 32, cells), block 128 (warp = 16 columns x 16 rows of the current panel)
constexpr int kTBs = kDP + 4;
template <int TR>
constexpr size_t trsm_smem_bytes(int np) {   // independent of the warp count
  return ((size_t)np * kDP * (TR + 4) + 3 * kDP * kTBs + (size_t)np * kDP) * sizeof(double);
}

// TR rows per CTA (32 or 64), NW warps (4 or 8): a warp owns two 8-column tiles x TR / (NW/2) rows of the current panel.
// TR = 64 halves the number of times the 32x32 operand blocks are re-streamed from L2; with 4 warps it doubles the
// independent MMAs per fragment load, with 8 warps it keeps 16-row warp tiles and 16 warps/SM at 2 CTAs/SM.
// The last CTA of a launch may own only 32 valid rows (row counts are multiples of 32).
template <int TR, int NW>
__global__ void __launch_bounds__(32 * NW)
k_direct_trsm(double *__restrict__ band, size_t band_stride, long long col_off, int ld, int jc0, int np, int row_lo,
              int pglob0, int NP, const double *__restrict__ vinv, const double *__restrict__ dvec,
              double *__restrict__ ybuf, int ldy) {
  constexpr int NT = 32 * NW, kTLd = TR + 4, WR = TR / (NW / 2), NTL = WR / 8;   // threads, smem row stride, rows and n-tiles per warp
  extern __shared__ __align__(16) double trsm_smem[];
  double *Xs = trsm_smem;                                  // [32 np][kTLd]   A, then X
  double *Bs = Xs + (size_t)np * kDP * kTLd;               // [3][32][kTBs]   operand ring, [k][n]
  double *dinv = Bs + 3 * kDP * kTBs;                      // [32 np]
  const int cell = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = row_lo + blockIdx.x * TR;
  double *P = band + (size_t)cell * band_stride + col_off;
  const int W = np * kDP;
  // stage the A tile: W columns x TR rows = TR/2 chunks of 16 bytes per column
  for (int c = tid; c < W * (TR / 2); c += NT) {
    const int k = c / (TR / 2), i = (c % (TR / 2)) * 2;
    double *d = Xs + k * kTLd + i;
    if (r0 + i < ld) cp_async16(d, P + (size_t)(jc0 + k) * ld + r0 + i);
    else { d[0] = 0.0; d[1] = 0.0; }
  }
  cp_async_commit();
  for (int c = tid; c < W; c += NT) dinv[c] = 1.0 / dvec[(size_t)cell * NP + pglob0 + c];
  const int nblk = np * (np + 1) / 2;
  // block sequence: for p: L(p,0) .. L(p,p-1), V_p
  auto stage_blk = [&](int b, int p, int q) {
    double *dst = Bs + (size_t)(b % 3) * kDP * kTBs;
    const double *src;
    int lds;
    if (q < p) { src = P + (size_t)(jc0 + q * kDP) * ld + jc0 + p * kDP; lds = ld; }
    else { src = vinv + ((size_t)cell * NP + pglob0 + p * kDP) * kDP; lds = kDP; }
#pragma unroll
    for (int t = 0; t < 512 / NT; ++t) {
      const int c = tid + t * NT;
      const int k = c >> 4, n = (c & 15) * 2;
      cp_async16(dst + k * kTBs + n, src + (size_t)k * lds + n);
    }
  };
  int sp = 0, sq = 0;                                      // (p, q) of the next block to stage
  auto advance = [&](int &p, int &q) { if (q < p) ++q; else { ++p; q = 0; } };
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    if (b < nblk) { stage_blk(b, sp, sq); advance(sp, sq); }
    cp_async_commit();
  }
  const int wm = warp / (NW / 2), wn = warp % (NW / 2);    // wm: column tiles (MMA m), wn: row slice (MMA n)
  const int fr = lane >> 2, fk = lane & 3;
  // a warp owns the 8-column tiles {wm, 3 - wm} of the panel: V_p is lower triangular, so tile t needs only the
  // k-steps below 8 (t + 1); pairing tile t with 3 - t gives both warps 10 of the 16 k-steps (5/8 of the MMAs)
  const int m0[2] = {wm * 8, (3 - wm) * 8};
  const bool rows_valid = r0 + wn * WR < ld;               // warp-uniform (valid rows come in multiples of 32)
  double acc[2][NTL][2];
  int p = 0, q = 0;
  for (int b = 0; b < nblk; ++b) {
    cp_async_wait<1>();
    __syncthreads();
    if (b + 2 < nblk) { stage_blk(b + 2, sp, sq); advance(sp, sq); }
    cp_async_commit();
    const double *Bb = Bs + (size_t)(b % 3) * kDP * kTBs + fr;
    if (q == 0 && p > 0) {
      // acc := A_p (own 16 columns x WR rows)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) {
          const double2 v = *reinterpret_cast<const double2 *>(Xs + (size_t)(p * kDP + m0[mt] + fr) * kTLd + wn * WR + nt * 8 + fk * 2);
          acc[mt][nt][0] = v.x; acc[mt][nt][1] = v.y;
        }
    }
    if (q < p) {
      // acc -= L(p,q) X_q^T   (m = column of panel p, k = column of panel q, n = row)
      const double *Xq = Xs + (size_t)(q * kDP) * kTLd + wn * WR + fr;
#pragma unroll
      for (int ks = 0; ks < kDP / 4; ++ks) {
        const int kk = ks * 4 + fk;
        double af[2], bf[NTL];
#pragma unroll
        for (int t = 0; t < 2; ++t) af[t] = -Bb[kk * kTBs + m0[t]];
#pragma unroll
        for (int t = 0; t < NTL; ++t) bf[t] = Xq[kk * kTLd + t * 8];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
      }
      if (q == p - 1) {
        // T_p complete: publish it as an MMA operand for the V_p product (visible after the next barrier) in the
        // slot of A_p, which every warp has already moved into its accumulators (own tile positions only)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt)
            *reinterpret_cast<double2 *>(Xs + (size_t)(p * kDP + m0[mt] + fr) * kTLd + wn * WR + nt * 8 + fk * 2) =
                make_double2(acc[mt][nt][0], acc[mt][nt][1]);
      }
    } else {
      // X_p = V_p T_p  (m = column n' of panel p, k = column of T_p, n = row); T_p (T_0 = A_0) sits in slot p of Xs
      const double *Tq = Xs + (size_t)(p * kDP) * kTLd + wn * WR + fr;
      double x[2][NTL][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) x[mt][nt][0] = x[mt][nt][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < kDP / 4; ++ks) {
        const int kk = ks * 4 + fk;
        double bf[NTL];
#pragma unroll
        for (int t = 0; t < NTL; ++t) bf[t] = Tq[kk * kTLd + t * 8];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (ks * 4 >= m0[mt] + 8) continue;              // V_p[m][k] = 0 for k > m (warp-uniform)
          const double af = Bb[kk * kTBs + m0[mt]];
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) dmma_m8n8k4(x[mt][nt][0], x[mt][nt][1], af, bf[nt]);
        }
      }
      __syncthreads();                                     // everyone has read T_p before it is overwritten by X_p
      double *Y = ybuf + ((size_t)cell * kMaxWindow + p) * kDP * ldy;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int col = m0[mt] + fr;                       // column inside panel p
        const double di = dinv[p * kDP + col];
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) {
          const int row = wn * WR + nt * 8 + fk * 2;
          const double x0 = x[mt][nt][0], x1 = x[mt][nt][1];
          *reinterpret_cast<double2 *>(Xs + (size_t)(p * kDP + col) * kTLd + row) = make_double2(x0, x1);
          if (rows_valid) {
            *reinterpret_cast<double2 *>(P + (size_t)(jc0 + p * kDP + col) * ld + r0 + row) = make_double2(x0 * di, x1 * di);
            *reinterpret_cast<double2 *>(Y + (size_t)col * ldy + r0 + row) = make_double2(-x0, -x1);
          }
        }
      }
    }
    advance(p, q);
  }
  cp_async_wait<0>();
}

// ---- backward substitution, chunk by chunk ------------------------------------------------
// L^T x = z, chunks of <= 3 panels in reverse elimination order, two launches per chunk:
//  k_direct_back_gemm   T = z - L(rows below the chunk's diagonal region, chunk columns)^T x  for all panels of the
//                       chunk in parallel: grid (panels, cells).  The rows stream through a 4-stage cp.async ring
//                       of 32-row slices, so the band is read at HBM speed with long contiguous runs.
//  k_direct_back_diag   x = L_dd^-T T inside the diagonal region: one warp per cell, panels last to first,
//                       x_p = V_p^T (T_p - sum_{q>p} L(q,p)^T x_q), all products on mma.sync.m8n8k4.f64.
// xT[cell][j][NP] holds T between the two launches and x afterwards.
constexpr int kBLd = kDP + 4, kBStages = 4;
constexpr size_t kBackGemmSmem = (size_t)kBStages * (kDP + 24) * kBLd * sizeof(double);

__global__ void __launch_bounds__(128)
k_direct_back_gemm(const double *__restrict__ band, size_t band_stride, DirectPlanDev D, int s, int panel0, int i_lo, int k,
                   double *xT) {
  extern __shared__ __align__(16) double bg_smem[];        // [stage][32 cols + 24 rhs][kBLd]
  const int cell = blockIdx.y, j0 = (panel0 + blockIdx.x) * kDP, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fr = lane >> 2, fk = lane & 3;
  const int ld = D.ld[s], so = D.slab_off[s], choff = D.chunk_off[s], rows_dof = D.front_rows[s];
  const int NP = D.NP;
  const double *P = band + (size_t)cell * band_stride + D.col_off[s];
  double *x = xT + (size_t)cell * k * NP;
  const int nst = (rows_dof - i_lo) / kDP;                 // 32-row slices below the chunk's diagonal region
  // rhs rows >= k of every stage stay zero
  for (int i = tid; i < kBStages * 24 * kBLd; i += 128) {
    const int st = i / (24 * kBLd), r = i % (24 * kBLd);
    if (r / kBLd >= k) bg_smem[(size_t)st * (kDP + 24) * kBLd + kDP * kBLd + r] = 0.0;
  }
  auto load_stage = [&](int t) {
    double *dst = bg_smem + (size_t)(t % kBStages) * (kDP + 24) * kBLd;
    const int i0 = i_lo + t * kDP;                         // first front row of the slice
    const int ch = choff + (i0 >> 5);
    const int xi = D.slab_off[D.chunk_blk[ch]] + D.chunk_local[ch];   // padded unknown index of that row
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = tid + u * 128;
      const int col = c >> 4, i = (c & 15) * 2;
      cp_async16(dst + col * kBLd + i, P + (size_t)(j0 + col) * ld + i0 + i);
    }
    for (int c = tid; c < k * 16; c += 128) {
      const int j = c >> 4, i = (c & 15) * 2;
      cp_async16(dst + (kDP + j) * kBLd + i, x + (size_t)j * NP + xi + i);
    }
  };
#pragma unroll
  for (int t = 0; t < kBStages - 1; ++t) {
    if (t < nst) load_stage(t);
    cp_async_commit();
  }
  double acc[3][2];
#pragma unroll
  for (int nt = 0; nt < 3; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
  for (int t = 0; t < nst; ++t) {
    cp_async_wait<kBStages - 2>();
    __syncthreads();
    if (t + kBStages - 1 < nst) load_stage(t + kBStages - 1);
    cp_async_commit();
    const double *Ls = bg_smem + (size_t)(t % kBStages) * (kDP + 24) * kBLd + (warp * 8 + fr) * kBLd;
    const double *Xs = bg_smem + (size_t)(t % kBStages) * (kDP + 24) * kBLd + (kDP + fr) * kBLd;
#pragma unroll
    for (int ks = 0; ks < kDP / 4; ++ks) {
      const double a = Ls[ks * 4 + fk];                    // A[m = column][k = row] = L(row, column)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) dmma_m8n8k4(acc[nt][0], acc[nt][1], a, Xs[nt * 8 * kBLd + ks * 4 + fk]);
    }
  }
  cp_async_wait<0>();
  const int c = warp * 8 + fr;
#pragma unroll
  for (int nt = 0; nt < 3; ++nt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = nt * 8 + fk * 2 + h;
      if (j < k) x[(size_t)j * NP + so + j0 + c] = P[(size_t)(j0 + c) * ld + rows_dof + j] - acc[nt][h];
    }
}

// grid (ceil(cells/4)), block 128 (warp per cell); panels [p_lo, p_hi) of block column s.  from_z: no rows below (T = z).
__global__ void __launch_bounds__(128)
k_direct_back_diag(const double *__restrict__ band, size_t band_stride, DirectPlanDev D, int s, int p_lo, int p_hi, int k,
                   int n_cells, const double *__restrict__ vinv, int from_z, double *xT) {
  __shared__ double ts[4][kDP][24 + 1];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cell = blockIdx.x * 4 + w;
  if (cell >= n_cells) return;
  const int fr = lane >> 2, fk = lane & 3;
  const int ld = D.ld[s], so = D.slab_off[s], rows_dof = D.front_rows[s], NP = D.NP;
  const double *P = band + (size_t)cell * band_stride + D.col_off[s];
  double *x = xT + (size_t)cell * k * NP;
  for (int p = p_hi - 1; p >= p_lo; --p) {
    const int j0 = p * kDP;
    double acc[4][3][2];
    // t := T_p (m = column c, n = rhs j)
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = mt * 8 + fr, j = nt * 8 + fk * 2 + h;
          acc[mt][nt][h] = j < k ? (from_z ? P[(size_t)(j0 + c) * ld + rows_dof + j] : x[(size_t)j * NP + so + j0 + c]) : 0.0;
        }
    // t -= L(q,p)^T x_q for the panels q > p of this block column
    for (int q = p + 1; q < p_hi; ++q) {
      const double *Lq = P + (size_t)(j0 + fr) * ld + q * kDP + fk;
      const double *xq = x + (size_t)fr * NP + so + q * kDP + fk;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        double af[4][4], bf[4][3];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = (half * 4 + u) * 4;
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) af[u][mt] = -Lq[(size_t)(mt * 8) * ld + r];
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) bf[u][nt] = (nt * 8 + fr < k) ? xq[(size_t)(nt * 8) * NP + r] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], af[u][mt], bf[u][nt]);
      }
    }
    // x_p = V_p^T t: t becomes the B operand through shared memory
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) ts[w][mt * 8 + fr][nt * 8 + fk * 2 + h] = acc[mt][nt][h];
    __syncwarp();
    double xo[4][3][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) xo[mt][nt][0] = xo[mt][nt][1] = 0.0;
    const double *Vp = vinv + ((size_t)cell * NP + so + j0) * kDP;    // V(n, kk) at [kk * 32 + n]
#pragma unroll
    for (int ks = 0; ks < kDP / 4; ++ks) {
      const int kk = ks * 4 + fk;
      double af[4], bf[3];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) af[mt] = Vp[(size_t)(mt * 8 + fr) * kDP + kk];   // A[m = c][k = kk] = V(kk, c)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) bf[nt] = ts[w][kk][nt * 8 + fr];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) dmma_m8n8k4(xo[mt][nt][0], xo[mt][nt][1], af[mt], bf[nt]);
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = nt * 8 + fk * 2 + h;
          if (j < k) x[(size_t)j * NP + so + j0 + mt * 8 + fr] = xo[mt][nt][h];
        }
    __syncwarp();
  }
}

// xT[cell][j][p] -> interleaved x[g][row][k][32]: a CTA transposes 32 cells x 32 padded indices per right-hand side
// through shared memory, so both the per-cell reads and the cell-interleaved writes are 256-byte coalesced (the
// one-thread-per-entry version wrote 8 bytes per 256-byte line).  cell_lo is a multiple of 32.
// grid (NP/32, ceil(cells/32)), block (32, 8)
__global__ void __launch_bounds__(256)
k_direct_scatter_x(int NP, int NI, int k, const int *__restrict__ inv_perm, const double *__restrict__ xT,
                   int cell_lo, int n_cells, double *__restrict__ x) {
  __shared__ double tile[kLanes][kDP + 1];
  const int lane = threadIdx.x, w = threadIdx.y;
  const int p0 = blockIdx.x * kDP, c0 = blockIdx.y * kLanes;
  const int g = (cell_lo + c0) / kLanes;
  int rows[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) rows[t] = inv_perm[p0 + w + 8 * t];
  for (int j = 0; j < k; ++j) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int c = w + 8 * t;                              // cell inside the group
      tile[c][lane] = (c0 + c < n_cells) ? xT[((size_t)(c0 + c) * k + j) * NP + p0 + lane] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (rows[t] >= 0) x[(((size_t)g * NI + rows[t]) * k + j) * kLanes + lane] = tile[lane][w + 8 * t];
    __syncthreads();
  }
}

// rows whose solution is fixed to zero (RT_DQ pinned DoF) -- the interleaved x buffer is cleared first.

}  // namespace
}  // namespace msfec
