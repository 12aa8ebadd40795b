// tests/emu/emu_rt.cpp — the fiber scheduler and runtime stubs behind tests/emu/include/cuda_runtime.h.
// TEST INFRASTRUCTURE ONLY (see that header): it lets the CPU test-suite execute the product's
// kernel sources for their logic; it is linked into tests/emu/libcm_b200_emu.so and nothing else.
#include <cuda_runtime.h>
#undef threadIdx  // the kernel-side spellings are macros over coords(); here they are plain members
#undef blockIdx
#undef blockDim
#undef gridDim
#include <sched.h>
#include <stdio.h>
#include <sys/mman.h>

#include <chrono>
#include <vector>

namespace cmemu {

// ---- fibers: a minimal cooperative context switch -------------------------------------------
#if defined(__x86_64__)
extern "C" void cmemu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl cmemu_switch
.type cmemu_switch,@function
cmemu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cmemu_switch,.-cmemu_switch
)");
#else
#error "tests/emu needs an x86-64 host (fiber switch is written for the System V x86-64 ABI)"
#endif

constexpr size_t kStackBytes = 256 * 1024;
constexpr int kMaxThreads = 1024;

struct Fiber {
  void *sp = nullptr;
  bool done = false;
  uint3 tid{0, 0, 0};
};

struct WarpState {
  unsigned long long val[32];
  unsigned result[32];
  int arrived = 0;
  long gen = 0;
};

struct BlockState {
  char *stacks = nullptr;  // kMaxThreads stacks, allocated once per host thread
  std::vector<Fiber> fibers;
  std::vector<WarpState> warps;
  std::vector<unsigned char> smem;
  const std::function<void()> *body = nullptr;
  void *sched_sp = nullptr;
  int nthreads = 0, alive = 0, cur = -1;
  int bar_arrived = 0;
  long bar_gen = 0;
  int or_acc = 0, or_result = 0;
  long progress = 0;  // bumped by every barrier release / fiber exit (deadlock watchdog)
  ThreadCoords tc;
  ~BlockState();  // rank threads come and go (one set per test scenario): give the stacks back
};

BlockState::~BlockState() {
  if (stacks) munmap(stacks, kStackBytes * kMaxThreads);
}

static thread_local BlockState tl_block;

ThreadCoords &coords() { return tl_block.tc; }
// dynamic shared memory starts on a 1024-byte boundary, as on the device
unsigned char *dyn_smem() {
  return reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tl_block.smem.data()) + 1023) & ~(uintptr_t)1023);
}

static void yield() {
  BlockState &b = tl_block;
  Fiber &f = b.fibers[b.cur];
  cmemu_switch(&f.sp, b.sched_sp);
}

static void barrier_release(BlockState &b) {
  b.bar_arrived = 0;
  b.or_result = b.or_acc;
  b.or_acc = 0;
  ++b.bar_gen;
  ++b.progress;
}

void syncthreads() {
  BlockState &b = tl_block;
  const long gen = b.bar_gen;
  if (++b.bar_arrived == b.alive) {
    barrier_release(b);
  } else {
    while (b.bar_gen == gen) yield();
  }
}

int syncthreads_or(int pred) {
  BlockState &b = tl_block;
  if (pred) b.or_acc = 1;
  syncthreads();
  return b.or_result;  // stable until every thread has arrived at the next barrier
}

unsigned match_any(unsigned long long value) {
  BlockState &b = tl_block;
  const int lin = b.cur;
  WarpState &w = b.warps[lin >> 5];
  const int lane = lin & 31;
  const int lanes = std::min(32, b.nthreads - (lin & ~31));
  w.val[lane] = value;
  const long gen = w.gen;
  if (++w.arrived == lanes) {
    for (int i = 0; i < lanes; ++i) {
      unsigned m = 0;
      for (int j = 0; j < lanes; ++j)
        if (w.val[j] == w.val[i]) m |= 1u << j;
      w.result[i] = m;
    }
    w.arrived = 0;
    ++w.gen;
    ++b.progress;
  } else {
    while (w.gen == gen) yield();
  }
  return w.result[lane];
}

static void fiber_main() {
  BlockState &b = tl_block;
  (*b.body)();
  Fiber &f = b.fibers[b.cur];
  f.done = true;
  --b.alive;
  ++b.progress;
  // threads that leave early no longer take part in block barriers
  if (b.alive > 0 && b.bar_arrived == b.alive) barrier_release(b);
  cmemu_switch(&f.sp, b.sched_sp);
  abort();  // a finished fiber is never resumed
}

static void prepare_fiber(BlockState &b, int i) {
  char *top = b.stacks + (size_t)(i + 1) * kStackBytes;
  void **sp = reinterpret_cast<void **>(top);
  *--sp = nullptr;                                 // fake return address of fiber_main (16-byte aligned slot above)
  *--sp = reinterpret_cast<void *>(&fiber_main);   // popped by cmemu_switch's ret
  for (int k = 0; k < 6; ++k) *--sp = nullptr;     // rbp rbx r12 r13 r14 r15
  b.fibers[i].sp = sp;
  b.fibers[i].done = false;
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()> &thread_body) {
  BlockState &b = tl_block;
  if (b.cur >= 0) {
    fprintf(stderr, "[cm emu] nested kernel launch\n");
    abort();
  }
  const int nthreads = (int)(block.x * block.y * block.z);
  if (nthreads <= 0 || nthreads > kMaxThreads || grid.x == 0 || grid.y == 0 || grid.z == 0) {
    fprintf(stderr, "[cm emu] bad launch configuration: grid %u x %u x %u, block %d threads\n", grid.x, grid.y, grid.z,
            nthreads);
    abort();
  }
  if (smem > 227 * 1024) {
    fprintf(stderr, "[cm emu] %zu bytes of dynamic shared memory exceed the 227 KB of an sm_100 block\n", smem);
    abort();
  }
  if (!b.stacks) {
    void *p = mmap(nullptr, kStackBytes * kMaxThreads, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE,
                   -1, 0);
    if (p == MAP_FAILED) {
      perror("[cm emu] mmap of fiber stacks");
      abort();
    }
    b.stacks = static_cast<char *>(p);
  }
  b.body = &thread_body;
  b.nthreads = nthreads;
  b.fibers.assign(nthreads, Fiber());
  b.warps.assign((nthreads + 31) / 32, WarpState());
  b.smem.assign((smem ? smem : 1) + 1024, 0xCD);  // shared memory starts out undefined on a GPU (+ room to align its start)
  b.tc.blockDim = block;
  b.tc.gridDim = grid;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        b.tc.blockIdx = uint3{bx, by, bz};
        for (int i = 0; i < nthreads; ++i) {
          prepare_fiber(b, i);
          b.fibers[i].tid = uint3{(unsigned)i % block.x, ((unsigned)i / block.x) % block.y, (unsigned)i / (block.x * block.y)};
        }
        for (auto &w : b.warps) w.arrived = 0;
        b.alive = nthreads;
        b.bar_arrived = 0;
        b.or_acc = 0;
        long idle_passes = 0;
        while (b.alive > 0) {
          const long before = b.progress;
          for (int i = 0; i < nthreads; ++i) {
            Fiber &f = b.fibers[i];
            if (f.done) continue;
            b.cur = i;
            b.tc.threadIdx = f.tid;
            cmemu_switch(&b.sched_sp, f.sp);
          }
          if (b.progress == before) {
            if (++idle_passes > 4) {
              fprintf(stderr, "[cm emu] deadlock: block (%u,%u,%u): %d threads alive, %d at the block barrier\n", bx, by, bz,
                      b.alive, b.bar_arrived);
              abort();
            }
          } else {
            idle_passes = 0;
          }
        }
      }
  b.cur = -1;
  b.body = nullptr;
}

}  // namespace cmemu

// ---- runtime stubs: "device" memory is host memory, everything completes immediately ----------
struct cmemu_stream_st {
  int unused;
};
struct cmemu_event_st {
  double t_ms;
};

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

const char *cudaGetErrorString(cudaError_t e) {
  switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorMemoryAllocation: return "out of memory";
    case cudaErrorNotReady: return "not ready";
    case cudaErrorNotSupported: return "not supported by the CUDA test double";
    default: return "invalid value";
  }
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaGetDevice(int *dev) {
  *dev = 0;
  return cudaSuccess;
}
cudaError_t cudaSetDevice(int dev) { return dev == 0 ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetDeviceCount(int *n) {
  *n = 1;
  return cudaSuccess;
}
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
  memset(p, 0, sizeof(*p));
  snprintf(p->name, sizeof(p->name), "cm emu (CPU fibers, kernel-logic tests only)");
  p->major = 10;
  p->minor = 0;
  return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int) {
  if (a != cudaDevAttrMultiProcessorCount) {
    *v = 0;
    return cudaSuccess;
  }
  const char *e = getenv("CM_EMU_SMS");
  *v = e ? atoi(e) : 3;  // few "SMs": persistent kernels really loop, odd count exercises remainders
  return cudaSuccess;
}
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) {
  *lo = 0;
  *hi = -1;
  return cudaSuccess;
}
static cudaError_t alloc_bytes(void **p, size_t bytes) {
  *p = nullptr;
  if (posix_memalign(p, 256, bytes ? bytes : 256) != 0) return cudaErrorMemoryAllocation;
  memset(*p, 0xA5, bytes);  // cudaMalloc'ed memory is not zero
  return cudaSuccess;
}
cudaError_t cudaMalloc(void **p, size_t bytes) { return alloc_bytes(p, bytes); }
cudaError_t cudaFree(void *p) {
  free(p);
  return cudaSuccess;
}
cudaError_t cudaMallocHost(void **p, size_t bytes) { return alloc_bytes(p, bytes); }
cudaError_t cudaFreeHost(void *p) {
  free(p);
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) {
  memmove(dst, src, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) {
  memmove(dst, src, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                              cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < height; ++r)
    memmove(static_cast<char *>(dst) + r * dpitch, static_cast<const char *>(src) + r * spitch, width);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) {
  memset(p, v, n);
  return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
  *s = new cmemu_stream_st();
  return cudaSuccess;
}
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned f, int) { return cudaStreamCreateWithFlags(s, f); }
cudaError_t cudaStreamDestroy(cudaStream_t s) {
  delete s;
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) {
  *e = new cmemu_event_st{0.0};
  return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
  e->t_ms = now_ms();
  return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = (float)(b->t_ms - a->t_ms);
  return cudaSuccess;
}
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) {
  delete e;
  return cudaSuccess;
}
// one process, one address space: the loopback transport never needs IPC handles
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaErrorNotSupported; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) {
  a->type = cudaMemoryTypeUnregistered;
  return cudaSuccess;
}
// cuStreamWaitValue32(stream, addr, value, GEQ): streams complete immediately here, so the wait is a
// host spin until another rank thread's kernel has stored the value
static int emu_stream_wait_value32(cudaStream_t, unsigned long long addr, uint32_t value, unsigned) {
  volatile uint32_t *p = reinterpret_cast<volatile uint32_t *>(static_cast<uintptr_t>(addr));
  const double t0 = now_ms();
  while ((int32_t)(*p - value) < 0) {
    sched_yield();
    if (now_ms() - t0 > 30000.0) {  // a rank never signalled: report instead of hanging the test-suite
      fprintf(stderr, "[cm emu] cuStreamWaitValue32 timed out: *%p = %u, waiting for >= %u\n", (void *)p, *p, value);
      abort();
    }
  }
  __atomic_thread_fence(__ATOMIC_SEQ_CST);
  return 0;
}
cudaError_t cudaGetDriverEntryPoint(const char *sym, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *q) {
  if (!strcmp(sym, "cuStreamWaitValue32")) {
    *fn = reinterpret_cast<void *>(&emu_stream_wait_value32);
    *q = cudaDriverEntryPointSuccess;
    return cudaSuccess;
  }
  *fn = nullptr;
  *q = cudaDriverEntryPointSymbolNotFound;
  return cudaSuccess;
}
