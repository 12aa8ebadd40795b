// cm_transport.h — rank context and the exchange transports behind commit().
//
// The reference moves update data with UPC++ rpc / rput / rpc_ff and synchronises with
// reduce_all + progress spin + barrier (include/convergent_matrix.hpp:163-173, :226-244,
// :734-741). Here one commit is: counts all-gather -> all-to-all-v of the per-owner messages
// -> barrier. Two transports implement that:
//   NcclTransport      one process per GPU, grouped ncclSend/ncclRecv over NVLink (production)
//   LoopbackTransport  threads-as-ranks on ONE device, D2D copies (single-GPU test transport;
//                      same kernels, same message layout)
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

#include <string>
#include <vector>

namespace cmb {

int fail(int code, const char *fmt, ...);  // records the thread's last error, returns code
const char *last_error();

#define CM_CUDA_TRY(expr)                                                                  \
  do {                                                                                     \
    cudaError_t cm_e_ = (expr);                                                            \
    if (cm_e_ != cudaSuccess)                                                              \
      return ::cmb::fail(2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(cm_e_),     \
                         __FILE__, __LINE__);                                              \
  } while (0)

#define CM_TRY(expr)                 \
  do {                               \
    int cm_rc_ = (expr);             \
    if (cm_rc_ != 0) return cm_rc_;  \
  } while (0)

struct XferSpec {
  void *ptr;
  size_t bytes;
};

class Transport {
 public:
  virtual ~Transport() {}
  virtual const char *name() const = 0;
  // every rank contributes `bytes` from d_send; d_recv receives nranks*bytes (rank order)
  virtual int allgather(const void *d_send, void *d_recv, size_t bytes, cudaStream_t s) = 0;
  // personalised exchange: sends[r] goes to rank r, recvs[r] is filled by rank r; the self
  // entry is ignored. May be called with several "lanes" (values / metadata / coo) at once.
  virtual int alltoallv(const std::vector<std::vector<XferSpec>> &sends,
                        const std::vector<std::vector<XferSpec>> &recvs, cudaStream_t s) = 0;
  // returns when every rank's stream work up to this call is complete
  virtual int barrier(cudaStream_t s) = 0;
  // stream-ordered barrier: work enqueued on s after this call starts only when every rank's
  // stream has reached its own stream_barrier (no host synchronisation where the transport allows)
  virtual int stream_barrier(cudaStream_t s) = 0;
  // small host-side all-gather (IPC handles, pointers)
  virtual int host_allgather(const void *send, void *recv, size_t bytes, cudaStream_t s) = 0;
  // true if all ranks share this process's address space (peer pointers usable directly)
  virtual bool same_process() const = 0;
};

struct RankCtx {
  int rank = 0;
  int nranks = 1;
  int device = 0;
  Transport *tp = nullptr;
};

// current thread's context (loopback: bound per thread; nccl / single: process default)
RankCtx *current_ctx();

}  // namespace cmb
