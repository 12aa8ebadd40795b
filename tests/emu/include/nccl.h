// tests/emu/include/nccl.h — type-only stand-in (csrc/cm_transport.cu binds NCCL with dlopen at run
// time and only in the one-process-per-GPU mode, which the kernel-logic tests never enter).
#pragma once
#include <cuda_runtime.h>
typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclInt32 = 2 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
