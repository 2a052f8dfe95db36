// libmsfec_comm.so: NCCL exchange steps of the host driver (C ABI: include/msfec_comm.h).
// Replaces the Trilinos ghost import `locally_relevant_solution = distributed_solution`
// (reference source/Ned_RT/ned_rt_global.cc:460) and the MPI reductions of the reference's *Multiscale classes.
#include "../../include/msfec_comm.h"

#include <arpa/inet.h>
#include <cuda_runtime.h>
#include <nccl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>

struct msfec_comm {
  int rank = 0, world = 1, device = 0;
  ncclComm_t nccl = nullptr;
  cudaStream_t stream = nullptr;
  double *d_send = nullptr, *d_recv = nullptr;
  size_t cap_send = 0, cap_recv = 0;
};

namespace {

thread_local std::string g_err;

int fail(const std::string &m) { g_err = m; return 1; }

#define CK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::runtime_error(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)
#define CK_NCCL(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) throw std::runtime_error(std::string(#x) + ": " + ncclGetErrorString(r_)); } while (0)

void send_all(int fd, const void *buf, size_t n) {
  const char *p = (const char *)buf;
  while (n) { const ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL); if (w <= 0) throw std::runtime_error("rendezvous: send failed"); p += w; n -= (size_t)w; }
}
void recv_all(int fd, void *buf, size_t n) {
  char *p = (char *)buf;
  while (n) { const ssize_t r = ::recv(fd, p, n, 0); if (r <= 0) throw std::runtime_error("rendezvous: connection closed"); p += r; n -= (size_t)r; }
}

// rank 0 serves the id to world - 1 peers; the others fetch it (retrying while rank 0 is not listening yet)
void exchange_id(int rank, int world, ncclUniqueId &id) {
  const char *addr_env = std::getenv("MASTER_ADDR");
  const std::string addr = addr_env && *addr_env ? addr_env : "127.0.0.1";
  const int port = (std::getenv("MASTER_PORT") ? std::atoi(std::getenv("MASTER_PORT")) : 29500) + 17;
  if (rank == 0) {
    const int ls = ::socket(AF_INET, SOCK_STREAM, 0);
    if (ls < 0) throw std::runtime_error("rendezvous: socket()");
    int one = 1;
    ::setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in sa{};
    sa.sin_family = AF_INET; sa.sin_addr.s_addr = htonl(INADDR_ANY); sa.sin_port = htons((uint16_t)port);
    if (::bind(ls, (sockaddr *)&sa, sizeof(sa)) != 0 || ::listen(ls, world) != 0) { ::close(ls); throw std::runtime_error("rendezvous: cannot listen on port " + std::to_string(port)); }
    for (int p = 1; p < world; ++p) {
      const int fd = ::accept(ls, nullptr, nullptr);
      if (fd < 0) { ::close(ls); throw std::runtime_error("rendezvous: accept()"); }
      send_all(fd, &id, sizeof(id));
      ::close(fd);
    }
    ::close(ls);
  } else {
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_INET; hints.ai_socktype = SOCK_STREAM;
    if (::getaddrinfo(addr.c_str(), std::to_string(port).c_str(), &hints, &res) != 0 || !res) throw std::runtime_error("rendezvous: cannot resolve " + addr);
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
      const int fd = ::socket(AF_INET, SOCK_STREAM, 0);
      if (fd >= 0 && ::connect(fd, res->ai_addr, res->ai_addrlen) == 0) {
        try { recv_all(fd, &id, sizeof(id)); } catch (...) { ::close(fd); ::freeaddrinfo(res); throw; }
        ::close(fd);
        break;
      }
      if (fd >= 0) ::close(fd);
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(180)) { ::freeaddrinfo(res); throw std::runtime_error("rendezvous: rank 0 not reachable at " + addr + ":" + std::to_string(port)); }
      std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
    ::freeaddrinfo(res);
  }
}

void reserve(msfec_comm *c, size_t n_send, size_t n_recv) {
  if (n_send > c->cap_send) { cudaFree(c->d_send); CK_CUDA(cudaMalloc(&c->d_send, n_send * sizeof(double))); c->cap_send = n_send; }
  if (n_recv > c->cap_recv) { cudaFree(c->d_recv); CK_CUDA(cudaMalloc(&c->d_recv, n_recv * sizeof(double))); c->cap_recv = n_recv; }
}

int allreduce(msfec_comm *c, double *inout, size_t count, ncclRedOp_t op) {
  if (!c || !inout) return fail("null argument");
  if (count == 0) return 0;
  try {
    CK_CUDA(cudaSetDevice(c->device));
    reserve(c, count, count);
    CK_CUDA(cudaMemcpyAsync(c->d_send, inout, count * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK_NCCL(ncclAllReduce(c->d_send, c->d_recv, count, ncclDouble, op, c->nccl, c->stream));
    CK_CUDA(cudaMemcpyAsync(inout, c->d_recv, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  } catch (const std::exception &e) { return fail(e.what()); }
}

}  // namespace

extern "C" {

const char *msfec_comm_last_error(void) { return g_err.c_str(); }

int msfec_comm_create(int rank, int world, int device, msfec_comm **out) {
  if (!out || world < 1 || rank < 0 || rank >= world) return fail("bad argument");
  *out = nullptr;
  msfec_comm *c = new msfec_comm();
  c->rank = rank; c->world = world; c->device = device;
  try {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) throw std::runtime_error("no CUDA device available");
    if (device < 0 || device >= n_dev) throw std::runtime_error("CUDA device index out of range");
    CK_CUDA(cudaSetDevice(device));
    ncclUniqueId id;
    std::memset(&id, 0, sizeof(id));
    if (rank == 0) CK_NCCL(ncclGetUniqueId(&id));
    if (world > 1) exchange_id(rank, world, id);
    CK_NCCL(ncclCommInitRank(&c->nccl, world, id, rank));
    CK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    *out = c;
    return 0;
  } catch (const std::exception &e) {
    const std::string msg = e.what();
    msfec_comm_destroy(c);
    return fail(msg);
  }
}

void msfec_comm_destroy(msfec_comm *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->nccl) ncclCommDestroy(c->nccl);
  cudaFree(c->d_send); cudaFree(c->d_recv);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int msfec_comm_rank(const msfec_comm *c) { return c ? c->rank : -1; }
int msfec_comm_world(const msfec_comm *c) { return c ? c->world : -1; }

int msfec_comm_allgather(msfec_comm *c, const double *send, size_t count, double *recv) {
  if (!c || !send || !recv) return fail("null argument");
  if (count == 0) return 0;
  try {
    CK_CUDA(cudaSetDevice(c->device));
    reserve(c, count, count * c->world);
    CK_CUDA(cudaMemcpyAsync(c->d_send, send, count * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK_NCCL(ncclAllGather(c->d_send, c->d_recv, count, ncclDouble, c->nccl, c->stream));
    CK_CUDA(cudaMemcpyAsync(recv, c->d_recv, count * c->world * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  } catch (const std::exception &e) { return fail(e.what()); }
}

int msfec_comm_allreduce_sum(msfec_comm *c, double *inout, size_t count) { return allreduce(c, inout, count, ncclSum); }
int msfec_comm_allreduce_max(msfec_comm *c, double *inout, size_t count) { return allreduce(c, inout, count, ncclMax); }

}  // extern "C"
