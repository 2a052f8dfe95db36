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
