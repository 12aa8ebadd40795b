// cm_transport.cu — runtime (rank contexts) and the two exchange transports.
#include "cm_transport.h"

#include <dlfcn.h>
#include <nccl.h>  // types only: libnccl.so.2 is bound at run time with dlopen (no link dep)
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>

#include "../../include/cm_b200.h"

namespace cmb {

// ------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------
static thread_local char tl_err[1024] = "";

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_err, sizeof(tl_err), fmt, ap);
  va_end(ap);
  if (getenv("CM_B200_VERBOSE")) fprintf(stderr, "[cm_b200] error %d: %s\n", code, tl_err);
  return code;
}
const char *last_error() { return tl_err; }

// ------------------------------------------------------------------------------------
// NCCL binding (dlopen so that a torch-loaded libnccl.so.2 is shared, and so that the
// library loads on machines without NCCL when only one rank is used)
// ------------------------------------------------------------------------------------
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int load_nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.handle) return CM_OK;
  const char *names[] = {getenv("CM_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    if (!n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(CM_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define CM_BIND(field, sym)                                                        \
  *(void **)(&g_nccl.field) = dlsym(h, sym);                                       \
  if (!g_nccl.field) return fail(CM_ERR_NCCL, "libnccl.so.2 lacks symbol %s", sym)
  CM_BIND(GetUniqueId, "ncclGetUniqueId");
  CM_BIND(CommInitRank, "ncclCommInitRank");
  CM_BIND(CommDestroy, "ncclCommDestroy");
  CM_BIND(AllGather, "ncclAllGather");
  CM_BIND(AllReduce, "ncclAllReduce");
  CM_BIND(Send, "ncclSend");
  CM_BIND(Recv, "ncclRecv");
  CM_BIND(GroupStart, "ncclGroupStart");
  CM_BIND(GroupEnd, "ncclGroupEnd");
  CM_BIND(GetErrorString, "ncclGetErrorString");
  CM_BIND(GetVersion, "ncclGetVersion");
#undef CM_BIND
  g_nccl.handle = h;
  return CM_OK;
}

#define CM_NCCL_TRY(expr)                                                                       \
  do {                                                                                          \
    ncclResult_t cm_n_ = (expr);                                                                \
    if (cm_n_ != ncclSuccess)                                                                   \
      return ::cmb::fail(CM_ERR_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(cm_n_));    \
  } while (0)

class NcclTransport final : public Transport {
 public:
  ncclComm_t comm = nullptr;
  int rank, nranks;
  void *d_scratch = nullptr;  // small device scratch for host_allgather / barrier
  size_t scratch_bytes = 0;

  NcclTransport(int r, int n) : rank(r), nranks(n) {}
  ~NcclTransport() override {
    if (d_scratch) cudaFree(d_scratch);
    if (comm) g_nccl.CommDestroy(comm);
  }
  const char *name() const override { return "nccl"; }
  bool same_process() const override { return false; }

  int init(const void *uid) {
    ncclUniqueId id;
    memcpy(&id, uid, sizeof(id));
    CM_NCCL_TRY(g_nccl.CommInitRank(&comm, nranks, id, rank));
    return CM_OK;
  }
  int ensure_scratch(size_t bytes) {
    if (bytes <= scratch_bytes) return CM_OK;
    if (d_scratch) cudaFree(d_scratch);
    d_scratch = nullptr;
    scratch_bytes = 0;
    CM_CUDA_TRY(cudaMalloc(&d_scratch, bytes));
    scratch_bytes = bytes;
    return CM_OK;
  }
  int allgather(const void *d_send, void *d_recv, size_t bytes, cudaStream_t s) override {
    CM_NCCL_TRY(g_nccl.AllGather(d_send, d_recv, bytes, ncclInt8, comm, s));
    return CM_OK;
  }
  int alltoallv(const std::vector<std::vector<XferSpec>> &sends, const std::vector<std::vector<XferSpec>> &recvs,
                cudaStream_t s) override {
    // one group for all lanes and peers: NCCL fuses them into one kernel over NVLink
    CM_NCCL_TRY(g_nccl.GroupStart());
    for (size_t lane = 0; lane < sends.size(); ++lane) {
      for (int r = 0; r < nranks; ++r) {
        if (r == rank) continue;
        if (sends[lane][r].bytes)
          CM_NCCL_TRY(g_nccl.Send(sends[lane][r].ptr, sends[lane][r].bytes, ncclInt8, r, comm, s));
        if (recvs[lane][r].bytes)
          CM_NCCL_TRY(g_nccl.Recv(recvs[lane][r].ptr, recvs[lane][r].bytes, ncclInt8, r, comm, s));
      }
    }
    CM_NCCL_TRY(g_nccl.GroupEnd());
    return CM_OK;
  }
  int stream_barrier(cudaStream_t s) override {
    CM_TRY(ensure_scratch(256));
    CM_NCCL_TRY(g_nccl.AllReduce(d_scratch, d_scratch, 1, ncclInt32, ncclSum, comm, s));
    return CM_OK;
  }
  int barrier(cudaStream_t s) override {
    CM_TRY(stream_barrier(s));
    CM_CUDA_TRY(cudaStreamSynchronize(s));
    return CM_OK;
  }
  int host_allgather(const void *send, void *recv, size_t bytes, cudaStream_t s) override {
    CM_TRY(ensure_scratch(bytes * (nranks + 1)));
    char *d = static_cast<char *>(d_scratch);
    CM_CUDA_TRY(cudaMemcpyAsync(d, send, bytes, cudaMemcpyHostToDevice, s));
    CM_NCCL_TRY(g_nccl.AllGather(d, d + bytes, bytes, ncclInt8, comm, s));
    CM_CUDA_TRY(cudaMemcpyAsync(recv, d + bytes, bytes * nranks, cudaMemcpyDeviceToHost, s));
    CM_CUDA_TRY(cudaStreamSynchronize(s));
    return CM_OK;
  }
};

// ------------------------------------------------------------------------------------
// Loopback: all ranks are threads of this process on one device.
// ------------------------------------------------------------------------------------
struct LoopWorld {
  int n;
  std::atomic<int> bar_count{0};
  std::atomic<long> bar_gen{0};
  std::vector<const void *> board;                             // allgather sources
  std::vector<const std::vector<std::vector<XferSpec>> *> sends;  // alltoallv sources
  explicit LoopWorld(int nranks) : n(nranks), board(nranks, nullptr), sends(nranks, nullptr) {}
  void barrier() {
    const long gen = bar_gen.load(std::memory_order_acquire);
    if (bar_count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
      bar_count.store(0, std::memory_order_relaxed);
      bar_gen.fetch_add(1, std::memory_order_acq_rel);
    } else {
      while (bar_gen.load(std::memory_order_acquire) == gen) std::this_thread::yield();
    }
  }
};

class LoopbackTransport final : public Transport {
 public:
  LoopWorld *w;
  int rank;
  LoopbackTransport(LoopWorld *world, int r) : w(world), rank(r) {}
  const char *name() const override { return "loopback"; }
  bool same_process() const override { return true; }

  int allgather(const void *d_send, void *d_recv, size_t bytes, cudaStream_t s) override {
    CM_CUDA_TRY(cudaStreamSynchronize(s));
    w->board[rank] = d_send;
    w->barrier();
    for (int r = 0; r < w->n; ++r)
      CM_CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(d_recv) + (size_t)r * bytes, w->board[r], bytes,
                                  cudaMemcpyDeviceToDevice, s));
    CM_CUDA_TRY(cudaStreamSynchronize(s));
    w->barrier();
    return CM_OK;
  }
  int alltoallv(const std::vector<std::vector<XferSpec>> &sends, const std::vector<std::vector<XferSpec>> &recvs,
                cudaStream_t s) override {
    CM_CUDA_TRY(cudaStreamSynchronize(s));  // my packed segments are complete
    w->sends[rank] = &sends;
    w->barrier();
    for (size_t lane = 0; lane < recvs.size(); ++lane)
      for (int r = 0; r < w->n; ++r) {
        if (r == rank) continue;
        const XferSpec &src = (*w->sends[r])[lane][rank];
        if (src.bytes != recvs[lane][r].bytes)
          return fail(CM_ERR_STATE, "loopback: rank %d sends %zu bytes to %d which expects %zu", r, src.bytes, rank,
                      recvs[lane][r].bytes);
        if (src.bytes)
          CM_CUDA_TRY(cudaMemcpyAsync(recvs[lane][r].ptr, src.ptr, src.bytes, cudaMemcpyDeviceToDevice, s));
      }
    CM_CUDA_TRY(cudaStreamSynchronize(s));
    w->barrier();  // senders may reuse their arenas
    return CM_OK;
  }
  int barrier(cudaStream_t s) override {
    CM_CUDA_TRY(cudaStreamSynchronize(s));
    w->barrier();
    return CM_OK;
  }
  int stream_barrier(cudaStream_t s) override { return barrier(s); }
  int host_allgather(const void *send, void *recv, size_t bytes, cudaStream_t) override {
    w->board[rank] = send;
    w->barrier();
    for (int r = 0; r < w->n; ++r) memcpy(static_cast<char *>(recv) + (size_t)r * bytes, w->board[r], bytes);
    w->barrier();
    return CM_OK;
  }
};

class SingleTransport final : public Transport {
 public:
  const char *name() const override { return "single"; }
  bool same_process() const override { return true; }
  int allgather(const void *d_send, void *d_recv, size_t bytes, cudaStream_t s) override {
    CM_CUDA_TRY(cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, s));
    return CM_OK;
  }
  int alltoallv(const std::vector<std::vector<XferSpec>> &, const std::vector<std::vector<XferSpec>> &,
                cudaStream_t) override {
    return CM_OK;
  }
  int barrier(cudaStream_t s) override {
    CM_CUDA_TRY(cudaStreamSynchronize(s));
    return CM_OK;
  }
  int stream_barrier(cudaStream_t) override { return CM_OK; }
  int host_allgather(const void *send, void *recv, size_t bytes, cudaStream_t) override {
    memcpy(recv, send, bytes);
    return CM_OK;
  }
};

// ------------------------------------------------------------------------------------
// runtime state
// ------------------------------------------------------------------------------------
enum Mode { kNone = 0, kProcess = 1, kLoopback = 2 };

struct Runtime {
  std::mutex mu;
  Mode mode = kNone;
  RankCtx process_ctx;               // kProcess
  std::vector<RankCtx> loop_ctx;     // kLoopback
  LoopWorld *loop_world = nullptr;
  cudaStream_t rt_stream = nullptr;  // for runtime-level barriers (kProcess)
};
static Runtime g_rt;
static thread_local RankCtx *tl_ctx = nullptr;

RankCtx *current_ctx() {
  if (g_rt.mode == kProcess) return &g_rt.process_ctx;
  if (g_rt.mode == kLoopback) return tl_ctx;
  return nullptr;
}

static int check_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(CM_ERR_CUDA, "no usable CUDA device (%s); libcm_b200 has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= count) return fail(CM_ERR_INVALID, "device %d out of range [0,%d)", device, count);
  CM_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CM_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(CM_ERR_CUDA, "device %d is sm_%d%d; libcm_b200 is built for sm_100a only", device, prop.major,
                prop.minor);
  return CM_OK;
}

}  // namespace cmb

using namespace cmb;

extern "C" {

const char *cm_version(void) { return "cm_b200 0.1 (sm_100a)"; }
const char *cm_last_error(void) { return last_error(); }

int cm_rt_unique_id(void *out, size_t len) {
  if (!out || len < sizeof(ncclUniqueId)) return fail(CM_ERR_INVALID, "unique id buffer must hold 128 bytes");
  CM_TRY(load_nccl());
  ncclUniqueId id;
  CM_NCCL_TRY(g_nccl.GetUniqueId(&id));
  memcpy(out, &id, sizeof(id));
  return CM_OK;
}

int cm_rt_init(int rank, int nranks, int device, const void *unique_id) {
  std::lock_guard<std::mutex> lk(g_rt.mu);
  if (g_rt.mode != kNone) return fail(CM_ERR_STATE, "runtime already initialised");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(CM_ERR_INVALID, "bad rank %d of %d", rank, nranks);
  CM_TRY(check_device(device));
  RankCtx &c = g_rt.process_ctx;
  c.rank = rank;
  c.nranks = nranks;
  c.device = device;
  if (nranks == 1) {
    c.tp = new SingleTransport();
  } else {
    if (!unique_id) return fail(CM_ERR_INVALID, "unique_id is required when nranks > 1");
    CM_TRY(load_nccl());
    NcclTransport *t = new NcclTransport(rank, nranks);
    int rc = t->init(unique_id);
    if (rc != CM_OK) {
      delete t;
      return rc;
    }
    c.tp = t;
  }
  CM_CUDA_TRY(cudaStreamCreateWithFlags(&g_rt.rt_stream, cudaStreamNonBlocking));
  g_rt.mode = kProcess;
  return CM_OK;
}

// RANK / WORLD_SIZE / LOCAL_RANK (torchrun convention) + a file rendezvous for the NCCL id.
int cm_rt_init_from_env(void) {
  const char *r = getenv("RANK"), *w = getenv("WORLD_SIZE"), *l = getenv("LOCAL_RANK");
  const int rank = r ? atoi(r) : 0, nranks = w ? atoi(w) : 1, local = l ? atoi(l) : rank;
  if (nranks == 1) return cm_rt_init(0, 1, local, nullptr);
  char path[512];
  const char *f = getenv("CM_B200_UID_FILE");
  const char *port = getenv("MASTER_PORT");
  if (f)
    snprintf(path, sizeof(path), "%s", f);
  else
    snprintf(path, sizeof(path), "/tmp/cm_b200_uid.%s.%d", port ? port : "0", (int)getppid());
  unsigned char uid[128];
  if (rank == 0) {
    CM_TRY(cm_rt_unique_id(uid, sizeof(uid)));
    char tmp[600];
    snprintf(tmp, sizeof(tmp), "%s.tmp", path);
    FILE *fp = fopen(tmp, "wb");
    if (!fp) return fail(CM_ERR_STATE, "cannot write %s", tmp);
    fwrite(uid, 1, sizeof(uid), fp);
    fclose(fp);
    if (rename(tmp, path) != 0) return fail(CM_ERR_STATE, "cannot rename %s", tmp);
  } else {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
      FILE *fp = fopen(path, "rb");
      if (fp) {
        size_t n = fread(uid, 1, sizeof(uid), fp);
        fclose(fp);
        if (n == sizeof(uid)) break;
      }
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 120.0)
        return fail(CM_ERR_STATE, "timed out waiting for %s", path);
      std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
  }
  int rc = cm_rt_init(rank, nranks, local, uid);
  if (rc == CM_OK) {
    rc = cm_rt_barrier();
    if (rank == 0) unlink(path);
  }
  return rc;
}

int cm_rt_init_loopback(int nranks, int device) {
  std::lock_guard<std::mutex> lk(g_rt.mu);
  if (g_rt.mode != kNone) return fail(CM_ERR_STATE, "runtime already initialised");
  if (nranks < 1) return fail(CM_ERR_INVALID, "bad nranks %d", nranks);
  CM_TRY(check_device(device));
  g_rt.loop_world = new LoopWorld(nranks);
  g_rt.loop_ctx.resize(nranks);
  for (int r = 0; r < nranks; ++r) {
    g_rt.loop_ctx[r].rank = r;
    g_rt.loop_ctx[r].nranks = nranks;
    g_rt.loop_ctx[r].device = device;
    g_rt.loop_ctx[r].tp = new LoopbackTransport(g_rt.loop_world, r);
  }
  g_rt.mode = kLoopback;
  return CM_OK;
}

int cm_rt_bind_rank(int rank) {
  if (g_rt.mode != kLoopback) return fail(CM_ERR_STATE, "cm_rt_bind_rank needs cm_rt_init_loopback");
  if (rank < 0 || rank >= (int)g_rt.loop_ctx.size()) return fail(CM_ERR_INVALID, "bad rank %d", rank);
  tl_ctx = &g_rt.loop_ctx[rank];
  CM_CUDA_TRY(cudaSetDevice(tl_ctx->device));
  return CM_OK;
}

int cm_rt_finalize(void) {
  std::lock_guard<std::mutex> lk(g_rt.mu);
  if (g_rt.mode == kProcess) {
    delete g_rt.process_ctx.tp;
    g_rt.process_ctx = RankCtx();
    if (g_rt.rt_stream) cudaStreamDestroy(g_rt.rt_stream);
    g_rt.rt_stream = nullptr;
  } else if (g_rt.mode == kLoopback) {
    for (auto &c : g_rt.loop_ctx) delete c.tp;
    g_rt.loop_ctx.clear();
    delete g_rt.loop_world;
    g_rt.loop_world = nullptr;
  }
  tl_ctx = nullptr;
  g_rt.mode = kNone;
  return CM_OK;
}

int cm_rt_rank_me(void) {
  RankCtx *c = current_ctx();
  return c ? c->rank : -1;
}
int cm_rt_rank_n(void) {
  RankCtx *c = current_ctx();
  if (c) return c->nranks;
  if (g_rt.mode == kLoopback) return (int)g_rt.loop_ctx.size();
  return -1;
}
int cm_rt_device(void) {
  RankCtx *c = current_ctx();
  return c ? c->device : -1;
}
int cm_rt_barrier(void) {
  RankCtx *c = current_ctx();
  if (!c) return fail(CM_ERR_STATE, "runtime not initialised (or rank thread not bound)");
  return c->tp->barrier(g_rt.mode == kProcess ? g_rt.rt_stream : (cudaStream_t)0);
}

}  // extern "C"
