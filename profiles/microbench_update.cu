// Micro-benchmark of the trailing-update kernels of the batched block LDL^T (csrc/direct.cuh) on a synthetic
// band with the block structure of a C5 plane block column (bs 192, reaching a 160-layer and a 192-plane).
// Build + run (B200):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o /tmp/mb profiles/microbench_update.cu && /tmp/mb [cells]
// Prints time and FP64 TFLOP/s (algorithmic flops of the lower-trapezoid target region) per kernel variant and
// cross-checks every variant against the first one (max abs difference of the updated band).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CUDA_OK(x)                                                                      \
  do {                                                                                  \
    cudaError_t e_ = (x);                                                               \
    if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); std::exit(1); } \
  } while (0)

namespace msfec {
namespace {
constexpr int kLanes = 32;
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
}  // namespace
}  // namespace msfec
#include "../mpi-msfec_b200/csrc/direct.cuh"
namespace msfec {
namespace {
#include "r01_superseded_kernels.cuh"   // k_direct_update<TM, TN> (round-1 variant, out of the product since round 2)
}  // namespace
}  // namespace msfec

namespace msfec {
namespace {
// EXPERIMENT (not shipped): measured 18.3 / 20.2 / 20.7 TFLOP/s at K = 96 / 160 / 192 against 27.1 / 28.3 / 28.6 for the
// shared-memory ring of k_direct_update_s -- 8 poorly coalesced fragment loads per 16 DMMAs bind the LSU.
// ---- trailing update, warp-private version ---------------------------------------------------
// Same operation and tile enumeration as k_direct_update_s<64,64>, but every warp feeds its 32x32 DMMA tile straight
// from global memory (L1/L2) into MMA fragments: no shared memory, no block-wide barriers, no cp.async ring.  A fragment
// load of a warp touches 4 k-rows x 64 contiguous bytes (full sectors); the two warps of a CTA that share an operand
// hit in L1.  The next k-step's 8 fragment values are prefetched into registers while the current 16 DMMAs issue.
// grid (tiles, cells), block 128
__global__ void __launch_bounds__(128, 4)
k_direct_update_w(double *__restrict__ band, size_t band_stride, DirectPlanDev D, int s, int jsrc, int nq, int yslot0,
                  int vc_lo, int vc_hi, const double *__restrict__ ybuf, int ldy) {
  constexpr int TM = 64, TN = 64;
  const int cell = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = D.ld[s], front_rows = D.front_rows[s];
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
  const int wr = warp >> 1, wc = warp & 1;
  const int fr = lane >> 2, fk = lane & 3;
  const int vc0 = cbase + wc * 32, vr0 = rbase + wr * 32;
  if (!(vc0 < vc_hi && vc0 < front_rows && vr0 >= vc0 && vr0 < ld)) return;     // warp-uniform
  const bool diagw = vr0 == vc0;
  const long long col_off = D.col_off[s];
  const int choff = D.chunk_off[s];
  double *cb = band + (size_t)cell * band_stride;
  const double *Lp = cb + col_off + (size_t)jsrc * ld + (size_t)fk * ld + vr0 + fr;                      // B[k][n = row]
  const double *Yp = ybuf + ((size_t)cell * kMaxWindow + yslot0) * kDP * ldy + (size_t)fk * ldy + vc0 + fr;   // A[m = col][k]
  double *cdst;
  int ldc = ld;
  {
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
  const int nk = nq * (kDP / 4);
  double af[4], bf[4], an[4], bn[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) { af[t] = Yp[t * 8]; bf[t] = Lp[t * 8]; }
  double acc[4][4][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const double2 v = *reinterpret_cast<const double2 *>(cdst + (size_t)(mt * 8) * ldc + nt * 8);
      acc[mt][nt][0] = v.x; acc[mt][nt][1] = v.y;
    }
  for (int ks = 0; ks < nk; ++ks) {
    if (ks + 1 < nk) {
      const double *yp = Yp + (size_t)(ks + 1) * 4 * ldy, *lp = Lp + (size_t)(ks + 1) * 4 * ld;
#pragma unroll
      for (int t = 0; t < 4; ++t) { an[t] = yp[t * 8]; bn[t] = lp[t * 8]; }
    }
    if (diagw) {
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
#pragma unroll
    for (int t = 0; t < 4; ++t) { af[t] = an[t]; bf[t] = bn[t]; }
  }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      *reinterpret_cast<double2 *>(cdst + (size_t)(mt * 8) * ldc + nt * 8) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
}


}  // namespace
}  // namespace msfec

using namespace msfec;

template <typename T> T *upload(const std::vector<T> &v) {
  T *d; CUDA_OK(cudaMalloc(&d, v.size() * sizeof(T)));
  CUDA_OK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

__global__ void k_fill(double *p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)i * 2654435761u + seed; x ^= x >> 16; x *= 2246822519u; x ^= x >> 13;
    p[i] = (double)(x & 0xffff) / 65536.0 - 0.5;
  }
}
__global__ void k_maxdiff(const double *a, const double *b, size_t n, double *out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  double m = 0;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmax(m, fabs(a[i] - b[i]));
  atomicMax((unsigned long long *)out, (unsigned long long)__double_as_longlong(m));
}

int main(int argc, char **argv) {
  const int cells = argc > 1 ? std::atoi(argv[1]) : 1024;
  const int reps = 5;
  std::vector<int> bs = {192, 160, 192}, off = {0, 192, 352}, front = {544, 352, 192}, ld = {576, 384, 224};
  std::vector<long long> col = {0, 576LL * 192, 576LL * 192 + 384LL * 160};
  const size_t band_doubles = 576 * 192 + 384 * 160 + 224 * 192;
  std::vector<int> choff = {0}, chblk, chloc;
  const int nB = 3;
  std::vector<int> fpos(nB * nB, -1);
  for (int s = 0; s < nB; ++s) {
    int rows = 0;
    for (int b = s; b < nB; ++b) {
      fpos[s * nB + b] = rows;
      for (int i = 0; i < bs[b]; i += 32) { chblk.push_back(b); chloc.push_back(i); }
      rows += bs[b];
    }
    chblk.push_back(-1); chloc.push_back(0);
    choff.push_back((int)chblk.size());
  }
  DirectPlanDev D{nB, 544, upload(bs), upload(off), upload(ld), upload(front), upload(col), upload(choff), upload(chblk), upload(chloc), upload(fpos)};
  const int ldy = 576;
  double *band0, *band, *ref, *ybuf, *dmax;
  const size_t nband = (size_t)cells * band_doubles, ny = (size_t)cells * kMaxWindow * kDP * ldy;
  CUDA_OK(cudaMalloc(&band0, nband * 8)); CUDA_OK(cudaMalloc(&band, nband * 8)); CUDA_OK(cudaMalloc(&ref, nband * 8));
  CUDA_OK(cudaMalloc(&ybuf, ny * 8)); CUDA_OK(cudaMalloc(&dmax, 8));
  k_fill<<<2048, 256>>>(band0, nband, 1u); k_fill<<<2048, 256>>>(ybuf, ny, 7u);
  CUDA_OK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bool have_ref = false; int ref_nq = 0;
  auto run = [&](const char *name, int nq, int vc_lo, auto launch) {
    const double R = 576 - vc_lo, Cn = 544 - vc_lo;
    const double flops = 2.0 * 32 * nq * (Cn * R - Cn * (Cn - 1) / 2.0) * cells;
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
      CUDA_OK(cudaMemcpy(band, band0, nband * 8, cudaMemcpyDeviceToDevice));
      CUDA_OK(cudaEventRecord(e0)); launch(); CUDA_OK(cudaEventRecord(e1));
      CUDA_OK(cudaEventSynchronize(e1));
      CUDA_OK(cudaGetLastError());
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
    }
    double md = -1;
    if (!have_ref || ref_nq != nq * 1000 + vc_lo) { CUDA_OK(cudaMemcpy(ref, band, nband * 8, cudaMemcpyDeviceToDevice)); have_ref = true; ref_nq = nq * 1000 + vc_lo; }
    else {
      CUDA_OK(cudaMemset(dmax, 0, 8));
      k_maxdiff<<<2048, 256>>>(band, ref, nband, dmax);
      CUDA_OK(cudaMemcpy(&md, dmax, 8, cudaMemcpyDeviceToHost));
    }
    std::printf("%-34s nq %d vc_lo %3d  %8.3f ms  %6.2f TFLOP/s  maxdiff %g\n", name, nq, vc_lo, best, flops / best * 1e-9, md);
  };
#define OLD(TM, TN, nq, vc_lo)                                                                                         \
  {                                                                                                                    \
    CUDA_OK(cudaFuncSetAttribute(k_direct_update<TM, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_smem_bytes<TM, TN>(kMaxWindow))); \
    run("resident<" #TM "," #TN ">", nq, vc_lo, [&] {                                                                  \
      const int Tc = (544 - vc_lo + TN - 1) / TN, T = (576 - vc_lo + TM - 1) / TM;                                     \
      int Z = std::max(1, std::min(T, (4 * 148 * 4 + Tc * cells - 1) / (Tc * cells)));                                 \
      k_direct_update<TM, TN><<<dim3(Tc, Z, cells), (TM / 32) * (TN / 32) * 32, update_smem_bytes<TM, TN>(nq)>>>(band, band_doubles, D, 0, 0, nq, 0, vc_lo, 544, 576, ybuf, ldy); \
    });                                                                                                                \
  }
#define NEW(TM, TN, KC, ST, MB, nq, vc_lo)                                                                             \
  {                                                                                                                    \
    CUDA_OK(cudaFuncSetAttribute(k_direct_update_s<TM, TN, KC, ST, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_s_smem<TM, TN, KC, ST>())); \
    run("stream<" #TM "," #TN "," #KC "," #ST "," #MB ">", nq, vc_lo, [&] {                                            \
      k_direct_update_s<TM, TN, KC, ST, MB><<<dim3(update_s_tiles<TM, TN>(576, vc_lo, 544), cells), (TM / 32) * (TN / 32) * 32, update_s_smem<TM, TN, KC, ST>()>>>( \
          band, band_doubles, D, 0, 0, nq, 0, vc_lo, 544, ybuf, ldy);                                                  \
    });                                                                                                                \
  }
#define WARP(nq, vc_lo)                                                                                                \
  run("warp-private (no smem)", nq, vc_lo, [&] {                                                                        \
    k_direct_update_w<<<dim3(update_s_tiles<64, 64>(576, vc_lo, 544), cells), 128>>>(band, band_doubles, D, 0, 0, nq, 0, vc_lo, 544, ybuf, ldy); \
  });
  for (int nq : {3, 5, 6}) {
    NEW(64, 64, 8, 4, 4, nq, 192)
    WARP(nq, 192)
  }
  for (int nq : {3, 4, 6}) {
    if (nq <= 4) OLD(64, 64, nq, 192)
    NEW(64, 64, 8, 4, 4, nq, 192)
    NEW(64, 64, 16, 2, 4, nq, 192)
    NEW(64, 64, 16, 4, 3, nq, 192)
    NEW(128, 64, 8, 4, 2, nq, 192)
    NEW(128, 32, 8, 4, 4, nq, 192)
  }
  std::printf("done\n");
  return 0;
}
