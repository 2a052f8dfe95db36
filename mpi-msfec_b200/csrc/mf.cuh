// Batched multifrontal LDL^T kernels (plan: mfplan.h).  Replaces NedRTBasis::solve_direct / solve_iterative
// (reference source/Ned_RT/ned_rt_basis.cc:579-634, :637-847 and the Q / Q_Ned / RT_DQ siblings) for all cells and
// all k right-hand sides at once.
//
// One CTA = one front of one coarse cell; one launch = all fronts of one tree level of all cells of the sub-batch.
//   k_mf_forward   panel [m x s8] in shared memory := matrix entries (per-cell slot values on the shared pattern,
//                  cell-independent coupling entries, padding pivots) + rhs rows + the leading columns of the
//                  children's contribution blocks (extend-add through `cmap`);
//                  LDL^T of the own block / X = A L^-T of the rows below in 8-column steps: the left-looking update
//                  runs on mma.sync.m8n8k4.f64, the 8 x 8 pivot tile is factored by the warp that owns it (lane = row,
//                  shuffles) and its unit-lower factor inverted in place, the rows below are solved with the inverse on
//                  the tensor cores too;
//                  factor record -> global (read again by k_mf_backward only);
//                  contribution block C = sum_children C_child - X L21^T, 8 x 8 tiles on the FP64 tensor cores,
//                  children gathered through `pinv` straight into the accumulators -- or, where several fronts share
//                  an SM, streamed through a shared-memory ring by cp.async.bulk + mbarrier (template flag STG) --
//                  16-byte stores.
//   k_mf_backward  x_own = L11^-T (z_own - L21^T x_reached), top-down; positions, record and solution rows arrive as
//                  cp.async groups, product and back-substitution run on the FP64 tensor cores.
//   k_mf_scatter_x padded per-cell solution -> cell-interleaved layout.
// Everything else of a front lives in shared memory, so HBM sees: slot values and rhs once, every factor record
// written once and read once, every contribution block written once and read once (MfPlan::bytes_fwd / _bwd).
// Included by engine.cu.
#pragma once

namespace msfec {
namespace {

// One record per (level, position): everything a CTA needs to know about its front and its children, fetched with ONE round
// trip (the tables level_fronts -> fronts -> children -> fronts of the plan are four dependent ones).  Built by Engine::upload_mf.
struct MfRec {
  MfFront F;
  int32_t ch_ldc[8], ch_coff[8], ch_nown[8], ch_cmap[8], ch_pinv[8];   // children: column length u8 + kr, c_off, n_own, cmap_off, pinv_off
};

struct MfDev {
  const MfRec *recs;
  const MfFront *fronts;
  const MfChild *children;
  const int *front_idx, *own_rows, *cmap, *pinv, *pe_dest, *pe_ref, *ps_dest, *pc_dest, *level_fronts;
  const double *ps_val, *pc_val;
  int kr, NP;
};

// Phase switches for cost attribution (profiles/tools/phase_cost.sh; results in profiles/r02_mf_kernels.md): compiled in only
// with -DMSFEC_MF_PHASE_SWITCHES, the production library carries none of it.
#ifdef MSFEC_MF_PHASE_SWITCHES
__device__ int g_mf_dbg = 0;
#define MF_DBG_LOAD() const int dbg = g_mf_dbg
#else
#define MF_DBG_LOAD() constexpr int dbg = 0
#endif
constexpr int kMfMaxChildren = 8;
constexpr int kMfMaxStageBufs = 8;
constexpr int kMfFastChildren = 2;   // children handled by the unrolled gather of the contribution update

// Per-front record in the factor storage (written by k_mf_forward, read by k_mf_backward): the shared-memory image of
// the forward kernel, copied verbatim (one flat coalesced copy each way, no index arithmetic):
//   [ panel m x ldx | 1/d (s8) | d (s8) | INVERSES of the unit-lower factors of the 8 x 8 pivot tiles (s8 x 8) ]
// Panel rows below a pivot tile hold X = L D (unscaled); the consumers multiply by 1/d.
__host__ __device__ inline int mf_record_doubles(int m, int ldx, int s8) { return m * ldx + 10 * s8; }

// dynamic shared memory of k_mf_forward: the record, then the children's inverse maps
__host__ __device__ inline size_t mf_fwd_smem_bytes(int m, int ldx, int s8, int nch) {
  return (size_t)mf_record_doubles(m, ldx, s8) * sizeof(double) + (size_t)nch * m * sizeof(int);
}
// k_mf_forward with staged children (STG): the record, kMfStageBufs stage buffers of 8 child columns (ldc_max doubles each),
// the children's inverse maps, one presence flag per (tile column of the front, child)
__host__ __device__ inline size_t mf_fwd_st_smem_bytes(int m, int ldx, int s8, int u8, int nch, int ldc_max) {
  const size_t flags = ((size_t)((s8 + u8) / 8) * nch + 7) / 8 * 8;
  return ((size_t)mf_record_doubles(m, ldx, s8) + (size_t)kMfStageBufs * 8 * ldc_max) * sizeof(double) + (size_t)nch * m * sizeof(int) + flags;
}

// ---- mbarrier / bulk asynchronous copy (TMA, 1-D) primitives -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}
// global -> shared, completion counted in bytes on an mbarrier (16-byte aligned addresses, size a multiple of 16)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global as one bulk group
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// k_mf_backward: the record, a chunk of x of the reached unknowns (<= kMfBwdChunk rows x (kr + 4)), t / x of the own
// unknowns (s8 x (kr + 4)); the positions of the reached unknowns of two chunks sit in static shared memory
constexpr int kMfBwdChunk = 64;
constexpr int kMfBwdWarpSolve = 48;   // fronts up to this many own columns: barrier-free back-substitution, one warp per 8 right-hand sides
__host__ __device__ inline size_t mf_bwd_smem_bytes(int m, int ldx, int s8, int u8, int kr) {
  const int ch = u8 < kMfBwdChunk ? u8 : kMfBwdChunk;
  return ((size_t)mf_record_doubles(m, ldx, s8) + (size_t)(ch + s8) * (kr + 4)) * sizeof(double);
}

// grid (fronts of the level, cells of the sub-batch), block NT.  S > 0: every front of the level has s8 = 8 S own
// columns (compile-time panel width: the k-loops unroll and the A fragments of a tile column stay in registers);
// S = 0: run-time width.
// STG: the children's contribution blocks are not gathered element by element from global memory (a dependent round trip to
// HBM per step: 55-60 % of the stall samples of the levels above the leaves) but STREAMED: one elected thread issues 1-D bulk
// copies (cp.async.bulk, the TMA unit) of the child columns that fall into one 8-column tile column of this front, one stage
// = (tile column, child), kMfStageBufs stages in flight, completion on an mbarrier per buffer.  All warps work on the same
// tile column (row tiles dealt round-robin), pick their entries out of the staged columns through `pinv`, and release the
// buffer with the barrier that ends the stage.  The first stages are in flight while the panel is assembled and factored.
template <int NT, int MINB, int S, bool STG>
__global__ void __launch_bounds__(NT, MINB)
k_mf_forward(MfDev M, int lf_off, const double *__restrict__ vals, int n_slots, double kscale,
             const double *__restrict__ b, int NI, int k, int cell_lo, double *__restrict__ Lst, size_t l_stride,
             double *__restrict__ Cst, size_t c_stride, int *__restrict__ bad, int nbuf) {
  extern __shared__ __align__(16) double mf_smem[];
  __shared__ int ch_coff[kMfMaxChildren], ch_ldc[kMfMaxChildren], ch_nown[kMfMaxChildren], ch_cmap[kMfMaxChildren], ch_pinv[kMfMaxChildren];
  __shared__ __align__(8) unsigned long long full_bar[kMfMaxStageBufs];
  constexpr int NW = NT / 32;
  constexpr int KA = S > 0 ? 2 * S : 1;          // k-steps of a full-width product
  constexpr int TPS = (S > 0 && S <= 2) ? 8 : 4; // row tiles per step of the contribution update: narrow panels have almost no
                                                 // MMA work to hide the children gathers behind, so more loads are put in flight
  constexpr int TW = 4;                          // STG: row tiles of one tile column per warp (launch: row tiles <= TW * NW)
  MF_DBG_LOAD();
  const MfRec *R = M.recs + lf_off + blockIdx.x;
  const MfFront F = R->F;
  const int cell = blockIdx.y, gcell = cell_lo + cell, g = gcell / kLanes, ln = gcell % kLanes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fr = lane >> 2, fk = lane & 3;
  const int s8 = S > 0 ? 8 * S : F.s8, ldx = s8 + 4;
  const int u8 = F.u8, kr = M.kr, m = s8 + u8 + kr;
  const int nch = F.ch_hi - F.ch_lo;
  double *P = mf_smem;                          // [m][ldx]   the panel
  double *dinv = P + m * ldx;                   // [s8]       1 / d
  double *dval = dinv + s8;                     // [s8]       d
  double *Ld = dval + s8;                       // [s8][8]    unit-lower factors of the 8 x 8 pivot tiles
  const int ldcm = STG ? F.ldc_max : 0;
  double *stg = Ld + 8 * s8;                    // STG: [kMfStageBufs][8][ldcm]  staged child columns
  int *pinv_s = reinterpret_cast<int *>(stg + (STG ? nbuf : 0) * 8 * ldcm);   // [nch][m]
  unsigned char *has = reinterpret_cast<unsigned char *>(pinv_s + nch * m);   // STG: [(s8 + u8) / 8][nch]
  const int rec2 = mf_record_doubles(m, ldx, s8) >> 1;
  const double *Cbase = Cst + (size_t)cell * c_stride;
  const int TC = (s8 + u8) / 8;

  // STG: issue stage (T, ci) into buffer `buf`: the child columns that map into tile column T, each from its (even) diagonal row
  // to the end (lower triangle + rhs rows), 16-byte aligned on both sides
  auto issue_stage = [&](int T, int ci, int buf) {      // threads 0 .. 7: one child column each
    const int ldc = ch_ldc[ci];
    const double *Cc = Cbase + ch_coff[ci];
    uint32_t bytes = 0;
    const int c = tid;
    const int jc = pinv_s[ci * m + T * 8 + c];
    if (jc >= 0) {
      const int j0 = jc & ~1;
      bytes = (uint32_t)(ldc - j0) * sizeof(double);
      bulk_g2s(stg + (size_t)(buf * 8 + c) * ldcm, Cc + (size_t)jc * ldc + j0, bytes, &full_bar[buf]);
    }
    mbar_arrive_expect_tx(&full_bar[buf], bytes);
  };
  int pT = 0, pci = 0;                          // producer cursor (thread 0): next stage to issue
  auto issue_next = [&](int buf) {
    while (pT < TC) {
      const int T = pT, ci = pci;
      if (++pci == nch) { pci = 0; ++pT; }
      if (has[T * nch + ci]) { issue_stage(T, ci, buf); return; }
    }
  };
  int s_buf = 0;                                // ring position and phase parity of the next stage to consume (all threads agree)
  uint32_t s_par = 0;

  // ---- assemble the panel --------------------------------------------------------------------------------
  {
    double2 *P2 = reinterpret_cast<double2 *>(P);
    for (int i = tid; i < rec2; i += NT) P2[i] = make_double2(0.0, 0.0);
  }
  if (tid < nch) {
    ch_coff[tid] = R->ch_coff[tid]; ch_ldc[tid] = R->ch_ldc[tid]; ch_nown[tid] = R->ch_nown[tid];
    ch_cmap[tid] = R->ch_cmap[tid]; ch_pinv[tid] = R->ch_pinv[tid];
  }
  if (STG) {
    for (int ci = 0; ci < nch; ++ci) {
      const int po = R->ch_pinv[ci];
      for (int i = tid; i < m; i += NT) pinv_s[ci * m + i] = M.pinv[po + i];
    }
    if (tid == 0) {
      for (int bq = 0; bq < nbuf; ++bq) mbar_init(&full_bar[bq], 8);
      mbar_fence_init();
    }
  }
  __syncthreads();
  if (STG) {
    for (int t = tid; t < TC * nch; t += NT) {
      const int T = t / nch, ci = t - T * nch;
      const int *pv = pinv_s + ci * m + T * 8;
      bool any = false;
#pragma unroll
      for (int c = 0; c < 8; ++c) any = any || pv[c] >= 0;
      has[t] = any ? 1 : 0;
    }
    __syncthreads();
    if (tid < 8)
      for (int bq = 0; bq < nbuf; ++bq) issue_next(bq);
  }
  {
    // per-cell inputs: slot values and lifted right-hand sides, four entries per thread and pass so that the index loads and
    // then the value loads of a pass are in flight together (two round trips per pass; one pass covers the leaf fronts)
    const double *vc = vals + (size_t)g * n_slots * kLanes + ln;
    for (int e0 = F.pe_lo + tid; e0 < ((dbg & 16) ? F.pe_lo : F.pe_hi); e0 += 4 * NT) {
      int ref[4], dst[4];
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = min(e0 + u * NT, F.pe_hi - 1);
        ref[u] = M.pe_ref[e]; dst[u] = M.pe_dest[e];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = vc[(size_t)(ref[u] >> 1) * kLanes];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (e0 + u * NT < F.pe_hi) P[dst[u]] = (ref[u] & 1) ? -v[u] : v[u];
    }
    for (int e = F.pc_lo + tid; e < F.pc_hi; e += NT) P[M.pc_dest[e]] = M.pc_val[e];
    const double *bc = b + (size_t)g * NI * k * kLanes + ln;
    const int n_in = (dbg & 16) ? 0 : s8 * k;
    for (int o0 = tid; o0 < n_in; o0 += 4 * NT) {
      int row[4], c[4], j[4];
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int o = min(o0 + u * NT, n_in - 1);
        c[u] = o / k; j[u] = o - c[u] * k;
        row[u] = M.own_rows[F.row_off + c[u]];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = row[u] >= 0 ? bc[((size_t)row[u] * k + j[u]) * kLanes] : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (o0 + u * NT < n_in && row[u] >= 0) P[(s8 + u8 + j[u]) * ldx + c[u]] = v[u];
    }
  }
  __syncthreads();
  for (int e0 = F.ps_lo + tid; e0 < F.ps_hi; e0 += 2 * NT) {
    const int e1 = e0 + NT;
    const int d0 = M.ps_dest[e0], d1 = M.ps_dest[min(e1, F.ps_hi - 1)];
    const double v0 = M.ps_val[e0], v1 = M.ps_val[min(e1, F.ps_hi - 1)];
    P[d0] += v0 * kscale;
    if (e1 < F.ps_hi) P[d1] += v1 * kscale;
  }
  __syncthreads();
  if (STG) {
    // children, own tile columns of the front: every panel entry collects its child entries from the staged columns
    for (int q = 0; q < s8 / 8; ++q)
      for (int ci = 0; ci < nch; ++ci) {
        if (!has[q * nch + ci]) continue;
        const int buf = s_buf;
        mbar_wait(&full_bar[buf], s_par);
        const int *pv = pinv_s + ci * m;
        const double *sb = stg + (size_t)buf * 8 * ldcm;
        for (int o = q * 64 + tid; o < m * 8; o += NT) {
          const int c = o & 7, r = o >> 3, col = q * 8 + c;
          const int jc = pv[col], ir = pv[r];
          if (jc >= 0 && ir >= 0 && r >= col) P[r * ldx + col] += sb[c * ldcm + ir - (jc & ~1)];
        }
        __syncthreads();                         // the buffer is free again (and the panel complete for the next child)
        if (tid < 8) issue_next(buf);
        if (++s_buf == nbuf) { s_buf = 0; s_par ^= 1u; }
      }
  } else {
  // children: the leading n_own columns of a child's contribution block belong to this front's own columns
  for (int ci = 0; ci < nch; ++ci) {
    const int ldc = ch_ldc[ci];
    const double *Cc = Cst + (size_t)cell * c_stride + ch_coff[ci];
    const int *cmap = M.cmap + ch_cmap[ci];
    const int po = ch_pinv[ci];
    for (int i = tid; i < m; i += NT) pinv_s[ci * m + i] = M.pinv[po + i];
    for (int j = warp; j < ((dbg & 8) ? 0 : ch_nown[ci]); j += NW) {
      const int cj = cmap[j];
      const double *col = Cc + (size_t)j * ldc;
      // eight independent loads in flight per lane (two columns per pass with sixteen loads measured slower: registers)
      for (int i0 = j + lane; i0 < ldc; i0 += 256) {
        double v[8];
        int ri[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + 32 * u;
          ri[u] = i < ldc ? cmap[i] : -1;
          v[u] = i < ldc ? col[i] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) if (ri[u] >= 0) P[ri[u] * ldx + cj] += v[u];
      }
    }
    __syncthreads();
  }
  }

  // ---- factor: 8 columns at a time.  Row tile R of the panel belongs to warp R mod NW for the whole factorisation. ---------
  const int NS = s8 / 8, MT = m / 8;
#pragma unroll
  for (int q = 0; q < ((dbg & 64) ? 0 : (S > 0 ? S : NS)); ++q) {
    const int c0 = q * 8;
    const int pw = q % NW;                                   // the warp that owns the pivot tile
    const int Rf = q + ((warp - q) % NW + NW) % NW;          // this warp's first row tile >= q
    if (q > 0) {
      // left-looking update of tile column q:  P(R, q) -= X(R, 0:c0) L(q, 0:c0)^T   (L = X D^-1), four row tiles of a
      // warp at a time (independent accumulator chains).  MMA m = column inside the tile, n = row, k = earlier columns
      const double *Arow = P + (c0 + fr) * ldx + fk;
      double af[KA];
      if (S > 0) {
#pragma unroll
        for (int t = 0; t < KA; ++t) af[t] = t < 2 * q ? -Arow[4 * t] * dinv[4 * t + fk] : 0.0;
      }
      for (int R0 = Rf; R0 < MT; R0 += 4 * NW) {
        double acc[4][2];
        double *tp[4];
        const double *Brow[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int R = R0 + u * NW < MT ? R0 + u * NW : R0;   // past the end: the warp's own first tile again (result discarded)
          tp[u] = P + (R * 8 + 2 * fk) * ldx + c0 + fr;
          Brow[u] = P + (R * 8 + fr) * ldx + fk;
          acc[u][0] = tp[u][0]; acc[u][1] = tp[u][ldx];
        }
        if (S > 0) {
#pragma unroll
          for (int t = 0; t < KA; ++t)
            if (t < 2 * q) {
#pragma unroll
              for (int u = 0; u < 4; ++u) dmma_m8n8k4(acc[u][0], acc[u][1], af[t], Brow[u][4 * t]);
            }
        } else {
          for (int t = 0; t < c0; t += 4) {
            const double a = -Arow[t] * dinv[t + fk];
#pragma unroll
            for (int u = 0; u < 4; ++u) dmma_m8n8k4(acc[u][0], acc[u][1], a, Brow[u][t]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (R0 + u * NW < MT) { tp[u][0] = acc[u][0]; tp[u][ldx] = acc[u][1]; }
      }
    }
    // LDL^T of the pivot tile by the warp that has just updated it (the others are still updating their tiles): lane (i = lane & 7)
    // = row, pivot column broadcast by shuffles; pivots to shared memory, then the unit-lower factor is inverted in place (lane =
    // column, forward substitution on the unit vectors): the rows below are solved with it on the tensor cores
    if (warp == pw) {
      __syncwarp();
      const int i = lane & 7;
      double a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = (j <= i) ? P[(c0 + i) * ldx + c0 + j] : 0.0;
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
      double *Lt = Ld + c0 * 8;
      if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) Lt[i * 8 + j] = (j < i) ? a[j] : 0.0;
      }
      if (!ok && lane == 0) atomicExch(bad, 1);
      __syncwarp();
      double y[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        double v = (r == i) ? 1.0 : 0.0;                     // column i of L^-1
#pragma unroll
        for (int t = 0; t < r; ++t) v = fma(-Lt[r * 8 + t], y[t], v);
        y[r] = v;
      }
      __syncwarp();
      if (lane < 8) {
#pragma unroll
        for (int r = 0; r < 8; ++r) Lt[r * 8 + i] = y[r];
      }
    }
    __syncthreads();
    // rows below the pivot tile: X = A L^-T on the tensor cores (m = column, n = row, k = column of A), in place: a warp
    // reads the 8 x 8 tile it owns as B fragments and writes it back as accumulators
    {
      const double l0 = Ld[(c0 + fr) * 8 + fk], l1 = Ld[(c0 + fr) * 8 + 4 + fk];
      for (int R0 = (Rf > q ? Rf : Rf + NW); R0 < MT; R0 += 4 * NW) {
        double acc[4][2];
        double *tp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int R = R0 + u * NW < MT ? R0 + u * NW : R0;   // past the end: the warp's own first tile again (result discarded)
          const double *Br = P + (R * 8 + fr) * ldx + c0 + fk;
          tp[u] = P + (R * 8 + 2 * fk) * ldx + c0 + fr;
          acc[u][0] = acc[u][1] = 0.0;
          const double b0 = Br[0], b1 = Br[4];
          dmma_m8n8k4(acc[u][0], acc[u][1], l0, b0);
          dmma_m8n8k4(acc[u][0], acc[u][1], l1, b1);
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (R0 + u * NW < MT) { tp[u][0] = acc[u][0]; tp[u][ldx] = acc[u][1]; }
      }
    }
    __syncthreads();
  }

  // ---- factor record -> global: the shared-memory image as it is ---------------------------------------------------
  if (STG) {
    // one bulk copy shared -> global by the TMA unit (the generic-proxy writes of the factorisation are fenced first); nothing
    // below writes the record region, the issuing thread waits for the read side before the CTA leaves
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      char *dst = reinterpret_cast<char *>(Lst + (size_t)cell * l_stride + F.l_off);
      const char *src = reinterpret_cast<const char *>(P);
      const uint32_t total = (uint32_t)rec2 * 16u;
      for (uint32_t o = 0; o < total; o += 32768u) bulk_s2g(dst + o, src + o, min(32768u, total - o));
      bulk_commit();
    }
  } else {
    double2 *Lo = reinterpret_cast<double2 *>(Lst + (size_t)cell * l_stride + F.l_off);
    const double2 *P2 = reinterpret_cast<const double2 *>(P);
    for (int i = tid; i < ((dbg & 4) ? 0 : rec2); i += NT) Lo[i] = P2[i];
  }

  if (STG) {
    // ---- contribution block, children staged: all warps on tile column J, warp w takes the row tiles J + w, J + w + NW, ... ----
    const int UT = u8 / 8, RT = (u8 + kr) / 8, ldc = u8 + kr;
    double *Co = Cst + (size_t)cell * c_stride + F.c_off;
    for (int J = 0; J < UT; ++J) {
      const int colp = s8 + J * 8 + fr;
      double acc[TW][2];
#pragma unroll
      for (int u = 0; u < TW; ++u) acc[u][0] = acc[u][1] = 0.0;
      for (int ci = 0; ci < nch; ++ci) {
        if (!has[(s8 / 8 + J) * nch + ci]) continue;
        const int buf = s_buf;
        mbar_wait(&full_bar[buf], s_par);
        const int *pv = pinv_s + ci * m;
        const int jc = pv[colp];
        if (jc >= 0) {
          const double *sc = stg + (size_t)(buf * 8 + fr) * ldcm - (jc & ~1);
#pragma unroll
          for (int u = 0; u < TW; ++u) {
            const int I = J + warp + u * NW;
            if (I >= RT) continue;
            const int rowp = s8 + I * 8 + 2 * fk;
            const int2 ii = *reinterpret_cast<const int2 *>(pv + rowp);
            if (ii.x >= 0 && rowp >= colp) acc[u][0] += sc[ii.x];
            if (ii.y >= 0 && rowp + 1 >= colp) acc[u][1] += sc[ii.y];
          }
        }
        __syncthreads();
        if (tid < 8) issue_next(buf);
        if (++s_buf == nbuf) { s_buf = 0; s_par ^= 1u; }
      }
      if (J + warp >= RT) continue;
      const double *Arow = P + colp * ldx + fk;
      const double *Brow[TW];
#pragma unroll
      for (int u = 0; u < TW; ++u) Brow[u] = P + (s8 + min(J + warp + u * NW, RT - 1) * 8 + fr) * ldx + fk;
      if (S > 0) {
#pragma unroll
        for (int t = 0; t < KA; ++t) {
          const double a = -Arow[4 * t] * dinv[4 * t + fk];
#pragma unroll
          for (int u = 0; u < TW; ++u) dmma_m8n8k4(acc[u][0], acc[u][1], a, Brow[u][4 * t]);
        }
      } else {
        for (int t = 0; t < s8; t += 4) {
          const double a = -Arow[t] * dinv[t + fk];
#pragma unroll
          for (int u = 0; u < TW; ++u) dmma_m8n8k4(acc[u][0], acc[u][1], a, Brow[u][t]);
        }
      }
#pragma unroll
      for (int u = 0; u < TW; ++u) {
        const int I = J + warp + u * NW;
        if (I < RT) *reinterpret_cast<double2 *>(Co + (size_t)(J * 8 + fr) * ldc + I * 8 + 2 * fk) = make_double2(acc[u][0], acc[u][1]);
      }
    }
    if (tid == 0) bulk_wait_read();
    return;
  }

  // ---- contribution block: C(I, J) = sum_children C_child - X_I L_J^T.  A warp owns tile columns J (dealt in snake
  // order, long and short columns alternate); per column the A fragments are scaled once and kept in registers (S > 0);
  // TPS row tiles per step: the children are gathered through `pinv` straight into the accumulators and the four MMA
  // chains run interleaved ---------------------------------------------------------------------------------------------
  if (u8 > 0 && !(dbg & 32)) {
    const int UT = u8 / 8, RT = (u8 + kr) / 8, ldc = u8 + kr;
    double *Co = Cst + (size_t)cell * c_stride + F.c_off;
    for (int rnd = 0; rnd * NW < UT; ++rnd) {
      const int J = rnd * NW + ((rnd & 1) ? NW - 1 - warp : warp);
      if (J >= UT) continue;
      const int colp = s8 + J * 8 + fr;
      const double *Arow = P + colp * ldx + fk;
      double af[KA];
      if (S > 0) {
#pragma unroll
        for (int t = 0; t < KA; ++t) af[t] = -Arow[4 * t] * dinv[4 * t + fk];
      }
      // fronts of the dissection tree have two children, four where the separators of two leaf pairs were merged into their
      // parent's (fast path); more take the loop below
      int jc[kMfFastChildren];
      const double *cb[kMfFastChildren];
#pragma unroll
      for (int ci = 0; ci < kMfFastChildren; ++ci) {
        jc[ci] = (ci < nch && !(dbg & 1)) ? pinv_s[ci * m + colp] : -1;
        cb[ci] = jc[ci] >= 0 ? Cbase + ch_coff[ci] + (size_t)jc[ci] * ch_ldc[ci] : Cbase;
#ifdef MSFEC_MF_PHASE_SWITCHES
        if (dbg & 128) cb[ci] = jc[ci] >= 0 ? Cst + ch_coff[ci] + (size_t)jc[ci] * ch_ldc[ci] : Cst;   // every cell reads cell 0's blocks: L2 hits
        if (dbg & 256) cb[ci] = Cst + (jc[ci] & 7) * 32;                                              // a 2 KB window: L1 hits
#endif
      }
      for (int I0 = J; I0 < RT; I0 += TPS) {
        double acc[TPS][2];
#pragma unroll
        for (int u = 0; u < TPS; ++u) acc[u][0] = acc[u][1] = 0.0;
#pragma unroll
        for (int ci = 0; ci < kMfFastChildren; ++ci) {
          if (jc[ci] < 0) continue;
          const int *pv = pinv_s + ci * m;
#pragma unroll
          for (int u = 0; u < TPS; ++u) {
            const int I = I0 + u;
            if (I >= RT) continue;
            const int rowp = s8 + I * 8 + 2 * fk;
            const int2 ii = *reinterpret_cast<const int2 *>(pv + rowp);
            if (ii.x >= 0 && rowp >= colp) acc[u][0] += cb[ci][ii.x];
            if (ii.y >= 0 && rowp + 1 >= colp) acc[u][1] += cb[ci][ii.y];
          }
        }
        for (int ci = kMfFastChildren; ci < nch; ++ci) {
          const int *pv = pinv_s + ci * m;
          const int jcc = pv[colp];
          if (jcc < 0) continue;
          const double *cbb = Cbase + ch_coff[ci] + (size_t)jcc * ch_ldc[ci];
          for (int u = 0; u < TPS; ++u) {
            const int I = I0 + u;
            if (I >= RT) continue;
            const int rowp = s8 + I * 8 + 2 * fk;
            if (pv[rowp] >= 0 && rowp >= colp) acc[u][0] += cbb[pv[rowp]];
            if (pv[rowp + 1] >= 0 && rowp + 1 >= colp) acc[u][1] += cbb[pv[rowp + 1]];
          }
        }
        const double *Brow[TPS];
#pragma unroll
        for (int u = 0; u < TPS; ++u) Brow[u] = P + (s8 + min(I0 + u, RT - 1) * 8 + fr) * ldx + fk;
        if (S > 0) {
#pragma unroll
          for (int t = 0; t < KA; ++t) {
#pragma unroll
            for (int u = 0; u < TPS; ++u) dmma_m8n8k4(acc[u][0], acc[u][1], af[t], Brow[u][4 * t]);
          }
        } else {
          for (int t = 0; t < s8; t += 4) {
            const double a = -Arow[t] * dinv[t + fk];
#pragma unroll
            for (int u = 0; u < TPS; ++u) dmma_m8n8k4(acc[u][0], acc[u][1], a, Brow[u][t]);
          }
        }
#pragma unroll
        for (int u = 0; u < TPS; ++u)
          if (I0 + u < RT && !(dbg & 2))
            *reinterpret_cast<double2 *>(Co + (size_t)(J * 8 + fr) * ldc + (I0 + u) * 8 + 2 * fk) = make_double2(acc[u][0], acc[u][1]);
      }
    }
  }
}

// grid (fronts of the level, cells of the sub-batch), block NT.  xT[cell][NP][kr] holds x in the padded elimination order
// (the kr right-hand sides of one unknown are contiguous: the gather of the reached unknowns reads whole 192-byte rows).
// The front's record comes back into shared memory with one flat copy; t = D^-1 (X_z - X21^T x_reached) runs on the FP64
// tensor cores (m = own column, n = right-hand side, k = reached unknown), then L11^T x = t tile by tile.
template <int NT>
__global__ void __launch_bounds__(NT)
k_mf_backward(MfDev M, int lf_off, int k, const double *__restrict__ Lst, size_t l_stride, double *__restrict__ xT) {
  extern __shared__ __align__(16) double mf_smem[];
  constexpr int NW = NT / 32;
  const MfFront F = M.recs[lf_off + blockIdx.x].F;
  const int cell = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fr = lane >> 2, fk = lane & 3;
  const int s8 = F.s8, u8 = F.u8, kr = M.kr, NP = M.NP, ldx = s8 + 4, m = s8 + u8 + kr, ldt = kr + 4;
  double *P = mf_smem;                          // [m][ldx]
  const double *dinv = P + m * ldx;             // [s8]
  const double *Ld = dinv + 2 * s8;             // [s8][8]   inverses of the unit-lower pivot tiles
  const int rec = mf_record_doubles(m, ldx, s8);
  const int CH = u8 < kMfBwdChunk ? u8 : kMfBwdChunk;
  double *xu = P + rec;                         // [CH][ldt]  x of the reached unknowns, one chunk of rows at a time
  double *ts = xu + CH * ldt;                   // [s8][ldt]  t, then x of the own unknowns
  __shared__ int idx_s[2 * kMfBwdChunk];        // [2][CH]  padded positions of the reached unknowns, chunk c in half c & 1
  double *xc = xT + (size_t)cell * kr * NP;
  {
    // everything comes in asynchronously (cp.async: no registers, no dependent round trips): first the positions of the reached
    // unknowns, then the record; the gather of x_reached below is issued as soon as the positions have landed
    for (int r = tid; r < CH; r += NT) cp_async4(idx_s + r, M.front_idx + F.idx_off + s8 + r);
    cp_async_commit();
    const double2 *src = reinterpret_cast<const double2 *>(Lst + (size_t)cell * l_stride + F.l_off);
    double2 *dst = reinterpret_cast<double2 *>(P);
    for (int i = tid; i < (rec >> 1); i += NT) cp_async16(dst + i, src + i);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
  }
  const int CT = s8 / 8, JT = kr / 8, kr2 = kr >> 1;
  // t(c, j) = (X_z(j, c) - sum_r X21(r, c) x_reached(r, j)) / d_c : 8 x 8 tiles (c-tile, j-tile) over the warps, four
  // independent accumulator chains over the reached unknowns of a chunk; partial sums of the chunks meet in ts
  int r_lo = 0, ib = 0;
  do {
    const int nr = min(CH, u8 - r_lo);
    if (r_lo > 0) __syncthreads();              // the previous chunk has been consumed
    // x of the reached unknowns: one row of kr right-hand sides (192 bytes, contiguous in xT[cell][position][kr]) per unknown,
    // 16 bytes per asynchronous copy, all of them in flight at once
    for (int o = tid; o < nr * kr2; o += NT) {
      const int r = o / kr2, q = o - r * kr2;
      const int p = idx_s[ib * CH + r];
      double2 *d2 = reinterpret_cast<double2 *>(xu + r * ldt) + q;
      if (p >= 0) cp_async16(d2, reinterpret_cast<const double2 *>(xc + (size_t)p * kr) + q);
      else *d2 = make_double2(0.0, 0.0);
    }
    // the positions of the next chunk travel with this group
    for (int r = tid; r < min(CH, u8 - r_lo - CH); r += NT) cp_async4(idx_s + (ib ^ 1) * CH + r, M.front_idx + F.idx_off + s8 + r_lo + CH + r);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    for (int tix = warp; tix < CT * JT; tix += NW) {
      const int ct = tix / JT, jt = tix - ct * JT;
      double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
      const double *A = P + (s8 + r_lo + fk) * ldx + ct * 8 + fr;    // A[m = c][k = r] = X21(r, c)
      const double *B = xu + fk * ldt + jt * 8 + fr;                 // B[k = r][n = j] = x_reached(r, j)
      int r0 = 0;
      for (; r0 + 16 <= nr; r0 += 16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma_m8n8k4(a0[q], a1[q], A[(r0 + 4 * q) * ldx], B[(r0 + 4 * q) * ldt]);
      }
      for (; r0 < nr; r0 += 4) dmma_m8n8k4(a0[0], a1[0], A[r0 * ldx], B[r0 * ldt]);
      const double s0 = (a0[0] + a0[1]) + (a0[2] + a0[3]), s1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
      const int c = ct * 8 + fr, j = jt * 8 + 2 * fk;
      double v0 = (r_lo == 0 ? P[(s8 + u8 + j) * ldx + c] : ts[c * ldt + j]) - s0;
      double v1 = (r_lo == 0 ? P[(s8 + u8 + j + 1) * ldx + c] : ts[c * ldt + j + 1]) - s1;
      if (r_lo + CH >= u8) { const double di = dinv[c]; v0 *= di; v1 *= di; }   // last chunk: t = D^-1 (...)
      ts[c * ldt + j] = v0; ts[c * ldt + j + 1] = v1;
    }
    r_lo += CH; ib ^= 1;
  } while (r_lo < u8);
  __syncthreads();
  // L11^T x = t, 8 unknowns at a time, last tile first; below a pivot tile the panel holds X = L D, the record carries the
  // inverses of the unit-lower pivot tiles:  x = Linv^T t  and the update of the earlier tiles are 8 x 8 products on the tensor cores
  if (s8 <= kMfBwdWarpSolve) {
    // narrow fronts: the right-hand sides are independent, so each warp takes one tile of 8 of them through the whole
    // back-substitution without a CTA barrier
    for (int jt = warp; jt < JT; jt += NW) {
      const int j = jt * 8 + 2 * fk;
      for (int p = CT - 1; p >= 0; --p) {
        const int c0 = p * 8;
        const double *B = ts + (c0 + fk) * ldt + jt * 8 + fr;
        {
          double a0 = 0.0, a1 = 0.0;
          const double *A = Ld + (c0 + fk) * 8 + fr;          // A[m = i][k] = Linv(k, i)
          dmma_m8n8k4(a0, a1, A[0], B[0]);
          dmma_m8n8k4(a0, a1, A[32], B[4 * ldt]);
          __syncwarp();
          ts[(c0 + fr) * ldt + j] = a0; ts[(c0 + fr) * ldt + j + 1] = a1;
          __syncwarp();
        }
        for (int cpt = 0; cpt < p; ++cpt) {
          double a0 = 0.0, a1 = 0.0;
          const double *A = P + (c0 + fk) * ldx + cpt * 8 + fr;
          dmma_m8n8k4(a0, a1, A[0], B[0]);
          dmma_m8n8k4(a0, a1, A[4 * ldx], B[4 * ldt]);
          const int c = cpt * 8 + fr;
          const double di = dinv[c];
          ts[c * ldt + j] -= a0 * di; ts[c * ldt + j + 1] -= a1 * di;
        }
        __syncwarp();
      }
    }
    __syncthreads();
  } else
  for (int p = s8 / 8 - 1; p >= 0; --p) {
    // wide fronts: the tiles of the update go to all warps, two barriers per pivot tile
    const int c0 = p * 8;
    for (int jt = warp; jt < JT; jt += NW) {
      double a0 = 0.0, a1 = 0.0;
      const double *A = Ld + (c0 + fk) * 8 + fr;
      const double *B = ts + (c0 + fk) * ldt + jt * 8 + fr;
      dmma_m8n8k4(a0, a1, A[0], B[0]);
      dmma_m8n8k4(a0, a1, A[32], B[4 * ldt]);
      __syncwarp();
      ts[(c0 + fr) * ldt + jt * 8 + 2 * fk] = a0; ts[(c0 + fr) * ldt + jt * 8 + 2 * fk + 1] = a1;
    }
    __syncthreads();
    if (c0 > 0) {
      for (int tix = warp; tix < p * JT; tix += NW) {
        const int cpt = tix / JT, jt = tix - cpt * JT;
        double a0 = 0.0, a1 = 0.0;
        const double *A = P + (c0 + fk) * ldx + cpt * 8 + fr;
        const double *B = ts + (c0 + fk) * ldt + jt * 8 + fr;
        dmma_m8n8k4(a0, a1, A[0], B[0]);
        dmma_m8n8k4(a0, a1, A[4 * ldx], B[4 * ldt]);
        const int c = cpt * 8 + fr, j = jt * 8 + 2 * fk;
        const double di = dinv[c];
        ts[c * ldt + j] -= a0 * di; ts[c * ldt + j + 1] -= a1 * di;
      }
      __syncthreads();
    }
  }
  // own solution: s8 consecutive rows of xT, one contiguous block (columns k .. kr are zero right-hand sides: x = 0)
  double *xo = xc + (size_t)F.own_base * kr;
  for (int o = tid; o < s8 * kr2; o += NT) {
    const int c = o / kr2, q = o - c * kr2;
    reinterpret_cast<double2 *>(xo)[o] = *reinterpret_cast<const double2 *>(ts + c * ldt + 2 * q);
  }
}

// xT[cell][NP][kr] (padded elimination order, per cell) -> x[g][row][k][32] (interior numbering, cell-interleaved), transposed
// through shared memory so that both sides move whole sectors.  cell_lo is a multiple of 32, NP of 8, kr even.
// grid (NP / 4, ceil(cells / 32)), block (32, 8): two threads read the k values of one (cell, position) as 16-byte pairs,
// then warp w writes the rows (position, rhs j = w, w + 8, ...) with the 32 cells of the group side by side.
__global__ void __launch_bounds__(256)
k_mf_scatter_x(int NP, int NI, int k, int kr, const int *__restrict__ inv_perm, const double *__restrict__ xT, int cell_lo,
               int n_cells, double *__restrict__ x) {
  constexpr int PC = 4;                                    // positions per block (21 KB of shared memory: 10 CTAs per SM)
  __shared__ double tile[kLanes][PC * kMaxK + 1];          // [cell][position][rhs < k]
  __shared__ int rows[PC];
  const int lane = threadIdx.x, w = threadIdx.y, t = w * 32 + lane;
  const int p0 = blockIdx.x * PC, c0 = blockIdx.y * kLanes;
  const int g = (cell_lo + c0) / kLanes;
  if (t < PC) rows[t] = inv_perm[p0 + t];
  {
    const int half = t >> 7, c = (t & 127) >> 2, pp = t & 3;
    const bool ok = c0 + c < n_cells;
    const double2 *src = reinterpret_cast<const double2 *>(xT + ((size_t)(c0 + c) * NP + p0 + pp) * kr);
    double *dst = &tile[c][pp * k];
    for (int q = half; 2 * q < k; q += 2) {
      const double2 v = ok ? src[q] : make_double2(0.0, 0.0);
      dst[2 * q] = v.x;
      if (2 * q + 1 < k) dst[2 * q + 1] = v.y;
    }
  }
  __syncthreads();
  for (int pp = 0; pp < PC; ++pp) {
    const int row = rows[pp];
    if (row < 0) continue;
    for (int j = w; j < k; j += 8) x[(((size_t)g * NI + row) * k + j) * kLanes + lane] = tile[lane][pp * k + j];
  }
}

}  // namespace
}  // namespace msfec
