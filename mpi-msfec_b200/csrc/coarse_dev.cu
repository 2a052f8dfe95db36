// Device solve of the global coarse system: C ABI in include/msfec_coarse.h.
//
// Reference: the Trilinos solves of *Multiscale::solve_iterative -- Schur-complement CG with an inner CG on block (0,0)
// (source/Ned_RT/ned_rt_global.cc:330-461, q_ned_global.cc:331-470, rt_dq_global.cc:330-470) and CG for Q
// (source/Q/q_global.cc:327-363).  Same scheme and preconditioners as the host stand-in (host/coarse.cpp), with every
// vector and the four CSR blocks resident in HBM.  Per CG iteration: one SpMV (8 lanes per row, shuffle reduction), one
// dot kernel, one fused update kernel (x += alpha p, r -= alpha Ap, z = D^-1 r, ||r||^2 and r.z in the same pass) and
// one direction kernel; alpha and beta are formed ON THE DEVICE from the reduced dot products of the previous kernels,
// the host reads ||r||^2 only (the stopping test).  Dot products: per-block partial sums in a fixed partition, summed
// in index order by the last block to finish (threadfence + ticket), i.e. deterministic.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/msfec_coarse.h"

namespace {

thread_local std::string g_coarse_err;

#define CU_OK(call)                                                                                  \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 148 * 8;   // partial-sum slots; one full wave of 8 CTAs per SM
constexpr int kLanesPerRow = 8;

// device scalars of one CG instance
enum { S_RZ0 = 0, S_RZ1 = 1, S_PAP = 2, S_RR = 3, S_BB = 4, S_COUNT = 8 };

struct CsrDev {
  int n_rows = 0, n_cols = 0;
  int32_t *ptr = nullptr, *col = nullptr;
  double *val = nullptr;
};

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double warp_part[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();                                   // warp_part may still be read from a previous call
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x < 32) {
    s = threadIdx.x < kThreads / 32 ? warp_part[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  }
  return s;                                          // valid in thread 0
}

// Every block deposits its partial sum; the last block to arrive adds them up in index order and resets the ticket.
__device__ __forceinline__ void finish_sum(double v, double *partials, unsigned *ticket, double *out) {
  __shared__ bool last;
  const double s = block_sum(v);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double t = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += kThreads) t += partials[i];
    t = block_sum(t);
    if (threadIdx.x == 0) { *out = t; *ticket = 0u; }
  }
}

// y = alpha * A x (+ addend), kLanesPerRow lanes per row
__global__ void __launch_bounds__(kThreads)
k_coarse_spmv(CsrDev A, const double *__restrict__ x, double alpha, const double *addend, double *y) {
  const int sub = threadIdx.x % kLanesPerRow;
  const int rows_per_block = kThreads / kLanesPerRow;
  for (int row0 = blockIdx.x * rows_per_block; row0 < A.n_rows; row0 += gridDim.x * rows_per_block) {
    const int row = row0 + threadIdx.x / kLanesPerRow;
    double s = 0.0;
    if (row < A.n_rows)
      for (int e = A.ptr[row] + sub; e < A.ptr[row + 1]; e += kLanesPerRow) s = fma(A.val[e], x[A.col[e]], s);
#pragma unroll
    for (int o = kLanesPerRow / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, kLanesPerRow);
    if (row < A.n_rows && sub == 0) y[row] = alpha * s + (addend ? addend[row] : 0.0);
  }
}

__global__ void __launch_bounds__(kThreads)
k_coarse_dot(int n, const double *__restrict__ a, const double *__restrict__ b, double *partials, unsigned *ticket, double *out) {
  double s = 0.0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) s = fma(a[i], b[i], s);
  finish_sum(s, partials, ticket, out);
}

// x = 0, r = b, z = D^-1 r, p = z;  r.z -> S_RZ0, b.b -> S_BB
__global__ void __launch_bounds__(kThreads)
k_coarse_cg_init(int n, const double *__restrict__ b, const double *__restrict__ dinv, double *x, double *r, double *z, double *p,
                 double *partials, unsigned *tickets, double *scal) {
  double rz = 0.0, bb = 0.0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const double bi = b[i], zi = dinv[i] * bi;
    x[i] = 0.0; r[i] = bi; z[i] = zi; p[i] = zi;
    rz = fma(bi, zi, rz); bb = fma(bi, bi, bb);
  }
  finish_sum(rz, partials, tickets, scal + S_RZ0);
  finish_sum(bb, partials + kMaxBlocks, tickets + 1, scal + S_BB);
}

// alpha = rz_old / pAp (device scalars);  x += alpha p, r -= alpha Ap, z = D^-1 r;  ||r||^2 -> S_RR, r.z -> rz_new
__global__ void __launch_bounds__(kThreads)
k_coarse_cg_update(int n, const double *__restrict__ p, const double *__restrict__ Ap, const double *__restrict__ dinv,
                   double *x, double *r, double *z, double *partials, unsigned *tickets, double *scal, int rz_old, int rz_new) {
  const double alpha = scal[rz_old] / scal[S_PAP];
  double rr = 0.0, rz = 0.0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, Ap[i], r[i]);
    const double zi = dinv[i] * ri;
    r[i] = ri; z[i] = zi;
    rr = fma(ri, ri, rr); rz = fma(ri, zi, rz);
  }
  finish_sum(rr, partials, tickets, scal + S_RR);
  finish_sum(rz, partials + kMaxBlocks, tickets + 1, scal + rz_new);
}

// beta = rz_new / rz_old;  p = z + beta p
__global__ void __launch_bounds__(kThreads)
k_coarse_cg_direction(int n, const double *__restrict__ z, double *p, const double *__restrict__ scal, int rz_old, int rz_new) {
  const double beta = scal[rz_new] / scal[rz_old];
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) p[i] = fma(beta, p[i], z[i]);
}

struct Solver {
  cudaStream_t stream = nullptr;
  std::vector<void *> owned;
  double *h_pinned = nullptr;
  long long launches = 0;

  ~Solver() {
    for (void *p : owned) cudaFree(p);
    if (h_pinned) cudaFreeHost(h_pinned);
    if (stream) cudaStreamDestroy(stream);
  }
  template <typename T> T *alloc(size_t n) {
    void *p = nullptr;
    CU_OK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    owned.push_back(p);
    return (T *)p;
  }
  template <typename T> T *upload(const T *h, size_t n) {
    T *d = alloc<T>(n);
    if (n) CU_OK(cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, stream));
    return d;
  }
  CsrDev upload(const msfec_coarse_csr &A) {
    CsrDev D;
    D.n_rows = A.n_rows; D.n_cols = A.n_cols;
    const size_t nnz = (size_t)A.ptr[A.n_rows];
    D.ptr = upload(A.ptr, (size_t)A.n_rows + 1); D.col = upload(A.col, nnz); D.val = upload(A.val, nnz);
    return D;
  }
  static int blocks(int n) { return std::max(1, std::min(kMaxBlocks, (n + kThreads - 1) / kThreads)); }
  void spmv(const CsrDev &A, const double *x, double alpha, const double *addend, double *y) {
    const int rpb = kThreads / kLanesPerRow;
    k_coarse_spmv<<<std::max(1, std::min(65535, (A.n_rows + rpb - 1) / rpb)), kThreads, 0, stream>>>(A, x, alpha, addend, y);
    ++launches;
  }

  // workspace + scalars of one CG instance (the inner and the outer iteration each own one)
  struct Work {
    int n = 0;
    double *r, *z, *p, *Ap, *partials, *scal;
    unsigned *tickets;
  };
  Work make_work(int n) {
    Work w;
    w.n = n;
    w.r = alloc<double>(n); w.z = alloc<double>(n); w.p = alloc<double>(n); w.Ap = alloc<double>(n);
    w.partials = alloc<double>(2 * kMaxBlocks); w.scal = alloc<double>(S_COUNT); w.tickets = alloc<unsigned>(2);
    CU_OK(cudaMemsetAsync(w.tickets, 0, 2 * sizeof(unsigned), stream));
    CU_OK(cudaMemsetAsync(w.scal, 0, S_COUNT * sizeof(double), stream));
    return w;
  }
  double read_scalar(const double *d) {
    CU_OK(cudaMemcpyAsync(h_pinned, d, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CU_OK(cudaStreamSynchronize(stream));
    return *h_pinned;
  }

  // Jacobi-preconditioned CG on device vectors; returns the iteration count, negative when the cap was reached
  int cg(Work &w, const std::function<void(const double *, double *)> &op, const double *dinv, const double *b, double *x,
         double rtol, int max_it) {
    const int n = w.n, nb = blocks(n);
    k_coarse_cg_init<<<nb, kThreads, 0, stream>>>(n, b, dinv, x, w.r, w.z, w.p, w.partials, w.tickets, w.scal);
    ++launches;
    const double bn = std::sqrt(read_scalar(w.scal + S_BB));
    if (bn == 0.0) return 0;
    for (int it = 1; it <= max_it; ++it) {
      const int rz_old = (it - 1) & 1, rz_new = it & 1;    // S_RZ0 / S_RZ1 alternate
      op(w.p, w.Ap);
      k_coarse_dot<<<nb, kThreads, 0, stream>>>(n, w.p, w.Ap, w.partials, w.tickets, w.scal + S_PAP);
      k_coarse_cg_update<<<nb, kThreads, 0, stream>>>(n, w.p, w.Ap, dinv, x, w.r, w.z, w.partials, w.tickets, w.scal, rz_old, rz_new);
      launches += 2;
      const double rr = read_scalar(w.scal + S_RR);
      if (!(rr == rr)) throw std::runtime_error("coarse CG broke down (NaN residual)");
      if (std::sqrt(rr) <= rtol * bn) return it;
      k_coarse_cg_direction<<<nb, kThreads, 0, stream>>>(n, w.z, w.p, w.scal, rz_old, rz_new);
      ++launches;
    }
    return -max_it;
  }
};

void check_csr(const msfec_coarse_csr *A, int rows, int cols, const char *name) {
  if (!A || !A->ptr || (A->ptr[A->n_rows] && (!A->col || !A->val)))
    throw std::invalid_argument(std::string("msfec_coarse_solve_device: block ") + name + " is missing");
  if (A->n_rows != rows || A->n_cols != cols)
    throw std::invalid_argument(std::string("msfec_coarse_solve_device: block ") + name + " has the wrong shape");
}

}  // namespace

extern "C" const char *msfec_coarse_last_error(void) { return g_coarse_err.c_str(); }

extern "C" int msfec_coarse_solve_device(int device, const msfec_coarse_csr *A00, const msfec_coarse_csr *A01,
                                         const msfec_coarse_csr *A10, const msfec_coarse_csr *A11, const double *b0,
                                         const double *b1, double rtol_inner, double rtol_outer, double *x0, double *x1,
                                         msfec_coarse_stats *stats) {
  try {
    if (!A00 || !b0 || !x0) throw std::invalid_argument("msfec_coarse_solve_device: null argument");
    const int n0 = A00->n_rows, n1 = A10 ? A10->n_rows : 0;
    check_csr(A00, n0, n0, "(0,0)");
    if (n1 > 0) {
      if (!b1 || !x1) throw std::invalid_argument("msfec_coarse_solve_device: null argument");
      check_csr(A01, n0, n1, "(0,1)"); check_csr(A10, n1, n0, "(1,0)"); check_csr(A11, n1, n1, "(1,1)");
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count)
      throw std::runtime_error("no CUDA device for the coarse solve (there is no CPU fallback in this call)");
    CU_OK(cudaSetDevice(device));
    Solver S;
    CU_OK(cudaStreamCreate(&S.stream));
    CU_OK(cudaMallocHost(&S.h_pinned, sizeof(double)));
    cudaEvent_t ev0, ev1;
    CU_OK(cudaEventCreate(&ev0)); CU_OK(cudaEventCreate(&ev1));
    CU_OK(cudaEventRecord(ev0, S.stream));

    // Jacobi preconditioners (host, one pass over the blocks)
    std::vector<double> d0(n0), ds(n1);
    for (int r = 0; r < n0; ++r) {
      double d = 0;
      for (int e = A00->ptr[r]; e < A00->ptr[r + 1]; ++e) if (A00->col[e] == r) d = A00->val[e];
      d0[r] = d != 0 ? 1.0 / d : 1.0;
    }
    for (int r = 0; r < n1; ++r) {
      double d = 0;
      for (int e = A11->ptr[r]; e < A11->ptr[r + 1]; ++e) if (A11->col[e] == r) d += A11->val[e];
      for (int e = A10->ptr[r]; e < A10->ptr[r + 1]; ++e) d += A10->val[e] * A10->val[e] * d0[A10->col[e]];
      ds[r] = d > 0 ? 1.0 / d : 1.0;
    }
    const CsrDev D00 = S.upload(*A00);
    double *d_d0 = S.upload(d0.data(), n0), *d_b0 = S.upload(b0, n0), *d_x0 = S.alloc<double>(n0);
    Solver::Work w_in = S.make_work(n0);
    long long inner_total = 0;
    auto inv00 = [&](const double *rhs, double *out) {
      const int it = S.cg(w_in, [&](const double *v, double *y) { S.spmv(D00, v, 1.0, nullptr, y); }, d_d0, rhs, out, rtol_inner,
                          20 * n0 + 100);
      if (it < 0) throw std::runtime_error("coarse solve: inner CG on block (0,0) did not converge");
      inner_total += it;
    };
    int outer = 0;
    if (n1 == 0) {
      inv00(d_b0, d_x0);
    } else {
      const CsrDev D01 = S.upload(*A01), D10 = S.upload(*A10), D11 = S.upload(*A11);
      double *d_ds = S.upload(ds.data(), n1), *d_b1 = S.upload(b1, n1), *d_x1 = S.alloc<double>(n1);
      double *t0 = S.alloc<double>(n0), *t1 = S.alloc<double>(n0), *rhsS = S.alloc<double>(n1);
      Solver::Work w_out = S.make_work(n1);
      // S u = b1 - A10 A00^-1 b0,   S = A11 - A10 A00^-1 A01
      inv00(d_b0, t0);
      S.spmv(D10, t0, -1.0, d_b1, rhsS);
      outer = S.cg(w_out, [&](const double *v, double *y) {
        S.spmv(D01, v, 1.0, nullptr, t0);
        inv00(t0, t1);
        S.spmv(D11, v, 1.0, nullptr, y);
        S.spmv(D10, t1, -1.0, y, y);
      }, d_ds, rhsS, d_x1, rtol_outer, 20 * n1 + 100);
      if (outer < 0) throw std::runtime_error("coarse solve: Schur-complement CG did not converge");
      // sigma = A00^-1 (b0 - A01 u)
      S.spmv(D01, d_x1, -1.0, d_b0, t0);
      inv00(t0, d_x0);
      CU_OK(cudaMemcpyAsync(x1, d_x1, (size_t)n1 * sizeof(double), cudaMemcpyDeviceToHost, S.stream));
    }
    CU_OK(cudaMemcpyAsync(x0, d_x0, (size_t)n0 * sizeof(double), cudaMemcpyDeviceToHost, S.stream));
    CU_OK(cudaEventRecord(ev1, S.stream));
    CU_OK(cudaStreamSynchronize(S.stream));
    CU_OK(cudaGetLastError());
    float ms = 0;
    CU_OK(cudaEventElapsedTime(&ms, ev0, ev1));
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    if (stats) { stats->outer_iterations = outer; stats->inner_iterations = inner_total; stats->kernel_launches = S.launches; stats->ms_device = ms; }
    return 0;
  } catch (const std::exception &e) {
    g_coarse_err = e.what();
    return 1;
  }
}
