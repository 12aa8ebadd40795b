// tests/emu/include/cuda_runtime.h — TEST DOUBLE of the CUDA runtime (test infrastructure only).
//
// Purpose: compile the product's kernel sources (convergent-matrix_b200/csrc/*.cu) with the HOST
// compiler so that their logic — grouping, scans, packing, item building, the column-tile
// scatter-add, the commit protocol over the loopback transport — can be exercised by the CPU test
// suite (`-m "not gpu"`) in a container that has no GPU. It plays the same role for CUDA that
// oracle/upcxx/upcxx.hpp plays for UPC++.
//
// It is NOT a backend: the result is tests/emu/libcm_b200_emu.so, which only tests/ loads, by
// explicit path. libcm_b200.so (the product) is built by nvcc for sm_100a only and never sees this
// header; it has no CPU fallback.
//
// Execution model: a kernel launch runs synchronously on the calling host thread, block after
// block; the threads of a block are fibers scheduled round-robin, __syncthreads() and the warp
// collectives are rendezvous points between them. "Device" memory is host memory (filled with a
// garbage pattern at allocation), streams and events are ordering no-ops.
#pragma once

#ifndef CM_EMU
#error "tests/emu/include is only for -DCM_EMU builds of the kernel-logic test library"
#endif

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

// ---- CUDA language surface --------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static thread_local  // blocks of one host thread run one after the other
#define __constant__ static
#define __align__(n) alignas(n)

struct uint3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace cmemu {
struct ThreadCoords {
  uint3 threadIdx, blockIdx;
  dim3 blockDim, gridDim;
};
ThreadCoords &coords();  // of the fiber that is running on this host thread
void syncthreads();
int syncthreads_or(int pred);
unsigned match_any(unsigned long long value);
unsigned char *dyn_smem();
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()> &thread_body);

template <class K>
struct Launch {
  dim3 g, b;
  size_t smem;
  K kernel;
  template <class... A>
  void operator()(A... args) const {
    run_grid(g, b, smem, [&]() { kernel(args...); });
  }
};
template <class S, class K>
Launch<K> make_launch(dim3 g, dim3 b, size_t smem, S /*stream*/, K kernel) {
  return Launch<K>{g, b, smem, kernel};
}
}  // namespace cmemu

#define threadIdx (::cmemu::coords().threadIdx)
#define blockIdx (::cmemu::coords().blockIdx)
#define blockDim (::cmemu::coords().blockDim)
#define gridDim (::cmemu::coords().gridDim)

#define CM_LAUNCH(grid, block, smem, stream, ...) ::cmemu::make_launch(grid, block, smem, stream, &__VA_ARGS__)
#define CM_DYN_SMEM(name) unsigned char *name = ::cmemu::dyn_smem()

inline void __syncthreads() { ::cmemu::syncthreads(); }
inline int __syncthreads_or(int p) { return ::cmemu::syncthreads_or(p); }
template <class T>
inline unsigned __match_any_sync(unsigned /*full mask*/, T v) {
  return ::cmemu::match_any((unsigned long long)v);
}
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <class T>
inline T __ldg(const T *p) {
  return *p;
}
// one fiber runs at a time per host thread and rank threads never share a block: plain updates
template <class T>
inline T atomicAdd(T *p, T v) {
  T old = *p;
  *p = old + v;
  return old;
}
inline int atomicOr(int *p, int v) {
  int old = *p;
  *p = old | v;
  return old;
}
template <class T>
inline T min(T a, T b) {
  return b < a ? b : a;
}
template <class T>
inline T max(T a, T b) {
  return a < b ? b : a;
}

// ---- runtime API surface used by csrc/ ---------------------------------------------------
typedef int cudaError_t;
enum : int {
  cudaSuccess = 0,
  cudaErrorInvalidValue = 1,
  cudaErrorMemoryAllocation = 2,
  cudaErrorNotReady = 600,
  cudaErrorNotSupported = 801
};
typedef struct cmemu_stream_st *cudaStream_t;
typedef struct cmemu_event_st *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1, cudaEnableDefault = 0 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrCanFlushRemoteWrites = 98 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
struct cudaPointerAttributes {
  cudaMemoryType type;
};
struct cudaIpcMemHandle_t {
  char reserved[64];
};
struct cudaDeviceProp {
  char name[256];
  int major, minor;
};

const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDevice(int *dev);
cudaError_t cudaSetDevice(int dev);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int dev);
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int dev);
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi);
cudaError_t cudaMalloc(void **p, size_t bytes);
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t bytes);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t s);
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                              cudaMemcpyKind k, cudaStream_t s);
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned flags, int prio);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamQuery(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaEventQuery(cudaEvent_t e);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p);
cudaError_t cudaGetDriverEntryPoint(const char *sym, void **fn, unsigned long long flags, cudaDriverEntryPointQueryResult *q);

template <class K>
inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) {
  return cudaSuccess;
}
template <class K>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) {
  *n = 2;
  return cudaSuccess;
}
template <class T>
inline cudaError_t cudaMemcpyToSymbol(T &sym, const void *src, size_t n) {
  memcpy(&sym, src, n);
  return cudaSuccess;
}
