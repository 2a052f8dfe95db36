/*
 * msfec_comm.h -- C ABI of the inter-rank exchange steps of the MsFEC host driver (libmsfec_comm.so).
 *
 * The basis build itself needs no inter-GPU traffic (one rank per GPU, contiguous Morton chunks of coarse cells).  The
 * two cross-rank steps of the reference's *Multiscale::run() are
 *   - the exchange of coarse data around the global coarse solve: the reference assembles a distributed Trilinos
 *     matrix and imports ghost values of the coarse solution (`locally_relevant_solution = distributed_solution`,
 *     source/Ned_RT/ned_rt_global.cc:460; same in q_global.cc, q_ned_global.cc, rt_dq_global.cc).  Here every rank
 *     contributes the element matrices of its owned cells with ONE ncclAllGather, assembles and solves the (small)
 *     coarse system redundantly and reads the weights of its own cells -- no second exchange is needed;
 *   - reductions over ranks (error / solution norms, timing and iteration statistics): ncclAllReduce, sum, FP64.
 * Both run over NCCL (NVLink 5 / NVSwitch on a B200 box) on device staging buffers; the caller passes host buffers.
 *
 * Rendezvous without MPI: rank 0 creates the ncclUniqueId and serves it over TCP on MASTER_ADDR:MASTER_PORT + 17 (the
 * environment torchrun / torch.distributed.run sets; defaults 127.0.0.1:29500), the other ranks fetch it there.
 *
 * Every function returns 0 on success; msfec_comm_last_error() describes the last failure of the calling thread.
 */
#ifndef MSFEC_COMM_H_
#define MSFEC_COMM_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct msfec_comm msfec_comm;

/* rank / world: this process and the number of processes; device: CUDA device of this rank. */
int msfec_comm_create(int rank, int world, int device, msfec_comm **out);
void msfec_comm_destroy(msfec_comm *c);
int msfec_comm_rank(const msfec_comm *c);
int msfec_comm_world(const msfec_comm *c);

/* recv[world][count] <- send[count] of every rank (ncclAllGather, FP64).  Host buffers. */
int msfec_comm_allgather(msfec_comm *c, const double *send, size_t count, double *recv);

/* inout[count] <- sum over ranks (ncclAllReduce, ncclSum, FP64).  Host buffer. */
int msfec_comm_allreduce_sum(msfec_comm *c, double *inout, size_t count);

/* inout[count] <- max over ranks (ncclAllReduce, ncclMax, FP64): wall-clock / iteration statistics. */
int msfec_comm_allreduce_max(msfec_comm *c, double *inout, size_t count);

const char *msfec_comm_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MSFEC_COMM_H_ */
