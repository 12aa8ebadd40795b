// cm_matrix.cu — host side of ConvergentMatrix on one rank + the C ABI (include/cm_b200.h).
//
// Replaces, per reference include/convergent_matrix.hpp:
//   ctor/roc/zero-fill (:379-448)      Matrix::create
//   update()  (:604-615)               stage_slice(): append an UpdDesc; payload either stays
//                                      where the caller put it in HBM (device API, borrowed)
//                                      or is copied H2D into the staging arena (host API)
//   flush(thresh) (:330-375)           flush_local(): single-rank map + sorted scatter-add
//   commit() (:723-742)                commit(): map -> scan -> pack -> counts all-gather ->
//                                      all-to-all-v -> build -> sort -> scatter-add -> barrier
//   Bin<T> for elemental updates       host COO bins per owner (:192-208, :623-629)
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <vector>

#include "../../include/cm_b200.h"
#include "cm_common.h"
#include "cm_kernels.h"
#include "cm_transport.h"

namespace cmb {

// ------------------------------------------------------------------------------------
// small memory helpers
// ------------------------------------------------------------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return CM_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      want = bytes;
      e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(CM_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    cap = want;
    return CM_OK;
  }
  // exactly `bytes` (message buffers of threshold flushes: no growth margin)
  int alloc_exact(size_t bytes) {
    release();
    bytes = std::max(bytes, (size_t)256);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      return fail(CM_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    cap = bytes;
    return CM_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T *as() const {
    return static_cast<T *>(p);
  }
};

struct PinBuf {
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return CM_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(CM_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    cap = want;
    return CM_OK;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

// device bump arena with stable addresses (block list); reset() consolidates
struct DevArena {
  struct Block {
    char *d;
    size_t cap, used;
  };
  std::vector<Block> blocks;
  size_t min_block;
  explicit DevArena(size_t min_block_bytes) : min_block(min_block_bytes) {}
  int alloc(size_t bytes, void **out) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (blocks.empty() || blocks.back().used + bytes > blocks.back().cap) {
      Block b;
      b.cap = std::max(bytes, min_block);
      b.used = 0;
      cudaError_t e = cudaMalloc((void **)&b.d, b.cap);
      if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(CM_ERR_NOMEM, "staging arena: cudaMalloc(%zu) failed: %s", b.cap, cudaGetErrorString(e));
      }
      blocks.push_back(b);
    }
    Block &b = blocks.back();
    *out = b.d + b.used;
    b.used += bytes;
    return CM_OK;
  }
  size_t used() const {
    size_t u = 0;
    for (auto &b : blocks) u += b.used;
    return u;
  }
  // caller guarantees no kernel still reads the arena
  void reset() {
    if (blocks.size() > 1) {
      size_t total = 0;
      for (auto &b : blocks) {
        total += b.cap;
        cudaFree(b.d);
      }
      blocks.clear();
      Block nb;
      nb.cap = total;
      nb.used = 0;
      if (cudaMalloc((void **)&nb.d, nb.cap) == cudaSuccess)
        blocks.push_back(nb);
      else
        cudaGetLastError();
    } else if (!blocks.empty()) {
      blocks[0].used = 0;
    }
  }
  void release() {
    for (auto &b : blocks) cudaFree(b.d);
    blocks.clear();
  }
};

// pinned host arena with a device twin per block (raw ix/jx of the host API)
struct MirrorArena {
  struct Block {
    char *h, *d;
    size_t cap, used;
  };
  std::vector<Block> blocks;
  size_t min_block;
  explicit MirrorArena(size_t min_block_bytes) : min_block(min_block_bytes) {}
  int alloc(size_t bytes, void **h_out, void **d_out) {
    bytes = (bytes + 15) & ~(size_t)15;
    if (blocks.empty() || blocks.back().used + bytes > blocks.back().cap) {
      Block b;
      b.cap = std::max(bytes, min_block);
      b.used = 0;
      if (cudaMallocHost((void **)&b.h, b.cap) != cudaSuccess) {
        cudaGetLastError();
        return fail(CM_ERR_NOMEM, "index arena: cudaMallocHost(%zu) failed", b.cap);
      }
      if (cudaMalloc((void **)&b.d, b.cap) != cudaSuccess) {
        cudaGetLastError();
        cudaFreeHost(b.h);
        return fail(CM_ERR_NOMEM, "index arena: cudaMalloc(%zu) failed", b.cap);
      }
      blocks.push_back(b);
    }
    Block &b = blocks.back();
    *h_out = b.h + b.used;
    *d_out = b.d + b.used;
    b.used += bytes;
    return CM_OK;
  }
  size_t used() const {
    size_t u = 0;
    for (auto &b : blocks) u += b.used;
    return u;
  }
  int upload(cudaStream_t s, long *bytes_moved) {
    for (auto &b : blocks)
      if (b.used) {
        CM_CUDA_TRY(cudaMemcpyAsync(b.d, b.h, b.used, cudaMemcpyHostToDevice, s));
        *bytes_moved += (long)b.used;
      }
    return CM_OK;
  }
  void reset() {
    if (blocks.size() > 1) {
      size_t total = 0;
      for (auto &b : blocks) {
        total += b.cap;
        cudaFreeHost(b.h);
        cudaFree(b.d);
      }
      blocks.clear();
      Block nb;
      nb.cap = total;
      nb.used = 0;
      if (cudaMallocHost((void **)&nb.h, nb.cap) == cudaSuccess) {
        if (cudaMalloc((void **)&nb.d, nb.cap) == cudaSuccess)
          blocks.push_back(nb);
        else
          cudaFreeHost(nb.h);
      }
      cudaGetLastError();
    } else if (!blocks.empty()) {
      blocks[0].used = 0;
    }
  }
  void release() {
    for (auto &b : blocks) {
      cudaFreeHost(b.h);
      cudaFree(b.d);
    }
    blocks.clear();
  }
};

// host COO bin for elemental updates (reference Bin<T>, :192-208)
struct CooBin {
  std::vector<long> inds;
  std::vector<unsigned char> vals;
};

// device buffers of one scatter-add batch (one per pipelined panel)
struct ApplyWork {
  DevBuf msgs, segs, keys0, keys1, items0, items1, sort_tmp, recs, col_start, cuts;
  void release() {
    DevBuf *b[] = {&msgs, &segs, &keys0, &keys1, &items0, &items1, &sort_tmp, &recs, &col_start, &cuts};
    for (DevBuf *x : b) x->release();
  }
};

struct PendingTimer {
  int kind;  // 0 apply, 1 pack, 2 exchange
  cudaEvent_t a, b;
  long elements;
};

// One message packed by a threshold flush with several ranks (the XBuff of a Bin::flush inside update(),
// reference :220-245, :346, :614), waiting for the next commit. `remote`: the pack kernel stored it straight
// into the owner's flush region of this source over NVLink (the reference's rput, :168-171); otherwise it sits
// in buffers of the source and is delivered by the commit. Exchanged between the ranks at commit as plain bytes.
struct FlushDesc {
  int src, dst;
  int remote;
  int seq;            // order among this source's flushed messages
  long off_v, off_m;  // remote: byte offsets inside the (owner, source) flush region
  MsgCounts c;
};
struct FlushedMsg {
  FlushDesc d;
  DevBuf vals, meta;  // !remote
};

}  // namespace cmb

using namespace cmb;

// ------------------------------------------------------------------------------------
// the matrix object (opaque to C callers)
// ------------------------------------------------------------------------------------
struct cm_matrix {
  RankCtx *ctx = nullptr;
  int dtype = 0;
  size_t esz = 8;
  Geom g{};
  int P = 1, me = 0;
  long m_local = 0, n_local = 0;
  size_t local_elems = 0;
  void *d_local = nullptr;
  void *h_mirror = nullptr;
  bool mirror_pinned = false;  // page-locked (D2H at the link rate) unless that allocation failed
  bool mirror_valid = false;
  std::vector<void *> peer_local;  // remote read (operator(), :585-596): IPC / same-process
  std::vector<bool> peer_ipc_open;
  // receive arenas of every rank (mine included), see ensure_recv_arenas()
  bool p2p_possible = false;
  bool p2p = false;       // every rank maps every rank's receive arenas: the pack kernel stores into them (mode 2)
  int exchange_mode = 0;  // 0 auto, 1 transport all-to-all-v (grouped ncclSend/ncclRecv), 2 peer-to-peer stores
  std::vector<void *> peer_vals, peer_meta;
  std::vector<size_t> cap_vals, cap_meta;
  std::vector<bool> peer_arena_open;

  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // H2D payload copies of the host API (overlap with kernels)
  cudaEvent_t copy_done = nullptr;
  bool copies_pending = false;
  bool defer_copy_wait = false;  // inside a batch call: one wait for all payload DMAs at its end
  // payload copy being merged: consecutive updates of a batch whose host payloads and arena
  // slots are both contiguous travel as one cudaMemcpyAsync (fewer, larger DMAs)
  const char *merge_src = nullptr;
  char *merge_dst = nullptr;
  size_t merge_bytes = 0;
  // Deferred scatter-add (opt-in, cm_set_apply_policy): commit() delivers the messages and returns;
  // they stay in the receive arenas — an append-only log, laid out identically on every rank from the
  // all-gathered counts — until the block is read, the log exceeds its budget or cm_materialize().
  // One scatter-add over several commits' messages sweeps the local block once instead of once per commit.
  struct PendMsg {
    const char *meta;
    const void *vals;
    long nseg, nitems, nelems, nadj;
  };
  int apply_policy = 0;               // CM_APPLY_EAGER / CM_APPLY_DEFERRED / CM_APPLY_BACKGROUND
  bool bg_running = false;            // a background scatter-add of the log is (possibly) still on the apply stream
  long bg_min_elems = 16L << 20;      // background policy: launch one when this many logged elements wait
  size_t defer_budget = 8ull << 30;   // bytes of logged values per owner before everybody materialises
  std::vector<PendMsg> pend;          // messages logged in MY arenas since the last collective materialise
  size_t pend_applied = 0;            // prefix of `pend` a local materialise has already applied
  std::vector<long> pend_v, pend_m;   // [P] append offsets of EVERY owner's arenas (elements / bytes), same on all ranks
  long flush_threshold = CM_DEFAULT_FLUSH_THRESHOLD;
  int progress_interval = 1;
  int deterministic = 0;
  int sorted_sweep = 1;
  int profiling = 0;
  int keep_wire = 0;
  int apply_kernel = 0;        // 0 auto, 1 RED, 2 column tiles
  // auto: tile when (elements / local block elements) >= this. B200, NS-owner shape (4 GiB block) at 0.125:
  // tile 1.61 ms, RED 1.93 ms per commit; K3's cold owners at 0.056: RED
  double tile_density = 0.125;
  double apply_share = 1.0;  // share of the local block's columns the scatter-add being issued can reach (one panel: 1 / panels)

  // staged updates
  std::vector<UpdDesc> descs;
  std::vector<long> item_base;  // single-rank path: first item of each update
  long staged_elems = 0, staged_rows = 0, staged_cols = 0, staged_items = 0;
  bool borrowed_pending = false;  // device-API payloads staged since the last flush
  bool reads_local_block = false;  // a staged payload is the local block itself (fill_lower)
  DevArena payload_arena{64ull << 20};
  MirrorArena index_arena{4ull << 20};
  std::vector<CooBin> coo;  // [owner]
  long coo_total = 0;
  // threshold flushes with several ranks (flush_shipped): messages already packed since the last commit
  std::vector<FlushedMsg> flushed;
  // flush regions: every owner holds one region per source (flush_region_v / _m bytes each, the same on all
  // ranks), mapped by every rank; a source fills its region at an owner front to back between two commits
  DevBuf d_flush_vals, d_flush_meta;
  std::vector<void *> peer_flush_vals, peer_flush_meta;  // [owner] base of that owner's regions
  std::vector<bool> peer_flush_open;
  size_t flush_region_v = 0, flush_region_m = 0;
  std::vector<size_t> flush_used_v, flush_used_m;  // [owner] bytes this rank has used of its region there
  long max_staged_updates = (1L << kItemSegBits) - 1;  // segments one scatter-add can address (cm_debug_set_staging_limit)

  // fill_lower scratch (global indices of local columns / rows)
  DevBuf d_fill_ix, d_fill_jx;

  // device work buffers
  DevBuf d_descs, d_item_base, d_rl, d_rpos, d_roff, d_cl, d_cpos, d_coff, d_uflags, d_radj, d_coo_tmp;
  DevBuf d_valoff, d_idxoff, d_itemoff, d_totals;
  DevBuf d_send_vals, d_send_meta, d_recv_vals, d_recv_meta, d_send_coo, d_recv_coo;
  DevBuf d_counts_all, d_dst_ptrs, d_coo_counts;
  ApplyWork work[kMaxChunks];
  cudaStream_t apply_stream = nullptr;  // scatter-add of panel k while panel k+1 is packed / exchanged
  cudaEvent_t ev_panel[kMaxChunks] = {};
  cudaEvent_t ev_applied = nullptr;
  cudaEvent_t ev_bg = nullptr;  // main stream -> apply stream dependency of a background scatter-add
  // Panel-pipelined commits (peer-to-peer stores): on by default, CM_B200_PIPELINE=0 disables them.
  // B200, 8 GPUs: K1-weak 10.78 -> 9.39 ms per step, K3 23.4 -> 14.1 ms.
  int pipeline = 1;
  long pipeline_min_elems = 32L << 20;  // ... when the ranks stage at least this many elements each on average
  long pipeline_owner_min_elems = 128L << 20;  // ... and the busiest owner receives at least this many (CM_B200_PIPELINE_MIN_ELEMS sets both)
  int npanel_cfg = 4;                   // column panels of a pipelined commit (CM_B200_PANELS, at most kMaxChunks)
  int pipeline_apply_blocks = 2;        // resident scatter-add blocks per SM while the next panel is being packed
                                        // (CM_B200_PIPELINE_APPLY_BLOCKS; 0 = no cap)
  PinBuf h_counts;

  cm_stats stats{};
  std::vector<PendingTimer> timers;
  std::vector<cudaEvent_t> event_pool;

  // CM_B200_TRACE=n: per-phase timeline (CUDA events + host clock) of the first n commits of rank 0, on stderr
  struct TraceMark {
    const char *name;
    cudaEvent_t ev;
    double host_us;
  };
  int trace_left = 0;
  long trace_skip = 0;  // CM_B200_TRACE_SKIP: commits to let pass first (warm-up, arena growth)
  long commit_seq = 0;
  std::vector<TraceMark> marks;

  // wire capture (parity tests)
  struct DbgMsg {
    int dst;
    int group;  // threshold flushes before the commit: 0, 1, ... in order; the commit's own messages come last
    std::vector<char> meta, vals;
  };
  std::vector<DbgMsg> dbg_msgs;  // the messages of the last commit, owner-major, panel order
  std::vector<CooBin> dbg_coo;
  bool dbg_valid = false;
};

namespace {

int roc_ok(long v) { return v > 0; }

int flush_direct(cm_matrix *M, bool sync);

cudaEvent_t get_event(cm_matrix *M) {
  if (!M->event_pool.empty()) {
    cudaEvent_t e = M->event_pool.back();
    M->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

double host_now_us() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

// timeline mark: an event on `st` now, plus the host clock
void trace_mark(cm_matrix *M, const char *name, cudaStream_t st = nullptr) {
  if (M->trace_left <= 0 || M->commit_seq < M->trace_skip) return;
  cudaEvent_t e = get_event(M);
  cudaEventRecord(e, st ? st : M->stream);
  M->marks.push_back(cm_matrix::TraceMark{name, e, host_now_us()});
}

// print and drop the marks whose events have completed (call with the main stream synchronised)
void trace_dump(cm_matrix *M, bool end_of_commit) {
  if (M->marks.empty()) return;
  cudaEventSynchronize(M->marks.back().ev);
  const cm_matrix::TraceMark &base = M->marks.front();
  for (const auto &mk : M->marks) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, base.ev, mk.ev) != cudaSuccess) {
      cudaGetLastError();
      ms = -1;
    }
    fprintf(stderr, "[cm trace r%d c%ld] %-28s gpu %9.3f us   host %9.3f us\n", M->me, M->commit_seq, mk.name, ms * 1e3,
            mk.host_us - base.host_us);
    M->event_pool.push_back(mk.ev);
  }
  M->marks.clear();
  if (end_of_commit) M->trace_left -= 1;
}

struct ScopedTimer {
  cm_matrix *M;
  int kind;
  long elements;
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st;
  ScopedTimer(cm_matrix *m, int k, long el, cudaStream_t stream = nullptr)
      : M(m), kind(k), elements(el), st(stream ? stream : m->stream) {
    if (M->profiling) {
      a = get_event(M);
      b = get_event(M);
      cudaEventRecord(a, st);
    }
  }
  ~ScopedTimer() {
    if (M->profiling) {
      cudaEventRecord(b, st);
      M->timers.push_back(PendingTimer{kind, a, b, elements});
    }
  }
};

// stream must be synchronised
void resolve_timers(cm_matrix *M) {
  std::vector<PendingTimer> keep;
  for (auto &t : M->timers) {
    if (cudaEventQuery(t.b) == cudaErrorNotReady) {  // a background scatter-add still running on another stream
      keep.push_back(t);
      continue;
    }
    cudaGetLastError();
    float ms = 0;
    if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
      if (t.kind == 0) {
        M->stats.apply_ms += ms;
        M->stats.apply_launches += 1;
        M->stats.apply_elements += t.elements;
      } else if (t.kind == 1) {
        M->stats.pack_ms += ms;
      } else {
        M->stats.exchange_ms += ms;
      }
    } else {
      cudaGetLastError();
    }
    M->event_pool.push_back(t.a);
    M->event_pool.push_back(t.b);
  }
  M->timers.swap(keep);
}

int ceil_log2(long v) {
  int b = 0;
  while ((1L << b) < v) ++b;
  return b;
}

void clear_staged(cm_matrix *M, bool keep_coo = false) {
  M->descs.clear();
  M->item_base.clear();
  M->staged_elems = M->staged_rows = M->staged_cols = M->staged_items = 0;
  M->borrowed_pending = false;
  M->reads_local_block = false;
  M->payload_arena.reset();
  M->index_arena.reset();
  if (keep_coo) return;  // a threshold flush with several ranks: elemental updates stay in their host bins
  for (auto &b : M->coo) {
    b.inds.clear();
    b.vals.clear();
  }
  M->coo_total = 0;
}

int issue_merged_copy(cm_matrix *M);

// staged descriptors, raw index sets and the mapped-index arenas of a flush, on the device
int upload_staged(cm_matrix *M, int npanel = 1) {
  const int nupd = (int)M->descs.size();
  CM_TRY(issue_merged_copy(M));
  if (M->copies_pending) {  // host-API payloads travel on their own stream
    CM_CUDA_TRY(cudaEventRecord(M->copy_done, M->copy_stream));
    CM_CUDA_TRY(cudaStreamWaitEvent(M->stream, M->copy_done, 0));
    M->copies_pending = false;
  }
  CM_TRY(M->index_arena.upload(M->stream, &M->stats.h2d_bytes));
  CM_TRY(M->d_descs.ensure(sizeof(UpdDesc) * (size_t)std::max(nupd, 1)));
  if (nupd) CM_CUDA_TRY(cudaMemcpyAsync(M->d_descs.p, M->descs.data(), sizeof(UpdDesc) * (size_t)nupd, cudaMemcpyHostToDevice, M->stream));
  const int npr = (int)M->g.nprow + 1, npc = (int)M->g.npcol * npanel + 1;
  CM_TRY(M->d_rl.ensure(4 * (size_t)std::max(M->staged_rows, 1L)));
  CM_TRY(M->d_rpos.ensure(4 * (size_t)std::max(M->staged_rows, 1L)));
  CM_TRY(M->d_cl.ensure(4 * (size_t)std::max(M->staged_cols, 1L)));
  CM_TRY(M->d_cpos.ensure(4 * (size_t)std::max(M->staged_cols, 1L)));
  CM_TRY(M->d_roff.ensure(4 * (size_t)npr * std::max(nupd, 1)));
  CM_TRY(M->d_coff.ensure(4 * (size_t)npc * std::max(nupd, 1)));
  CM_TRY(M->d_uflags.ensure(4 * (size_t)std::max(nupd, 1)));
  CM_TRY(M->d_radj.ensure(4 * (size_t)npr * std::max(nupd, 1)));
  CM_CUDA_TRY(cudaMemsetAsync(M->d_uflags.p, 0, 4 * (size_t)std::max(nupd, 1), M->stream));
  return CM_OK;
}

MapArenas map_arenas(cm_matrix *M) {
  MapArenas ma;
  ma.rl = M->d_rl.as<int>();
  ma.rpos = M->d_rpos.as<int>();
  ma.roff = M->d_roff.as<int>();
  ma.cl = M->d_cl.as<int>();
  ma.cpos = M->d_cpos.as<int>();
  ma.coff = M->d_coff.as<int>();
  ma.uflags = M->d_uflags.as<int>();
  ma.radj = M->d_radj.as<int>();
  return ma;
}

// sort (optional) + scatter-add over nitems (key,item) pairs sitting in w.keys0 / w.items0
int sort_and_apply(cm_matrix *M, ApplyWork &w, cudaStream_t st, long nitems, long elements, long adjacent, int max_blocks_per_sm = 0) {
  if (nitems <= 0) return CM_OK;
  const uint64_t *items = w.items0.as<uint64_t>();
  const uint32_t *keys = w.keys0.as<uint32_t>();
  // which scatter-add: the column-tile kernel pays 2*s bytes per row of a touched column and no
  // atomics; RED pays one sector round trip per element. Dense flushes -> tile.
  // (a panel of a pipelined commit only reaches its share of the block's columns)
  const double density = (double)elements / ((double)M->m_local * (double)M->n_local * M->apply_share);
  // RED pays per 32-byte sector it touches, the tile kernel per element plus one sweep of the columns it reaches:
  // with c = the share of elements whose row directly follows the previous one (runs of consecutive indices, K3),
  // RED costs about 28.8 - 25 c ps per element, tiles 2 ps per element + 3.2 ps per element of the block (B200:
  // NS-owner shape, c = 0: 28.8 / break-even at 0.12 hits per element; K3's hot owner, c = 0.98: RED 3.8 ps)
  const double c = elements > 0 ? std::min(1.0, (double)adjacent / (double)elements) : 0.0;
  const double tile_from = M->tile_density * 26.8 / (26.8 - 25.0 * c);
  bool tile = M->apply_kernel == 2 || (M->apply_kernel == 0 && density >= tile_from);
  if (M->deterministic) tile = true;
  // the column sort pays when the flush is big enough for L2 locality to matter; a small sparse
  // flush (latency-bound commits of tiny updates) goes straight to RED
  const bool sorted = tile || (M->sorted_sweep && elements >= (4L << 20));
  if (sorted) {
    const int end_bit = std::max(1, ceil_log2(std::max(M->n_local, 2L)));
    const size_t tmp = sort_temp_bytes(nitems, end_bit);
    CM_TRY(w.sort_tmp.ensure(tmp));
    CM_TRY(w.keys1.ensure(4 * (size_t)nitems));
    CM_TRY(w.items1.ensure(8 * (size_t)nitems));
    launch_sort(w.sort_tmp.p, tmp, w.keys0.as<uint32_t>(), w.keys1.as<uint32_t>(), w.items0.as<uint64_t>(),
                w.items1.as<uint64_t>(), nitems, end_bit, st);
    M->stats.kernel_launches += 3;  // CUB radix sort passes (library metadata sort, not the hot op)
    items = w.items1.as<uint64_t>();
    keys = w.keys1.as<uint32_t>();
  }
  CM_TRY(w.recs.ensure(sizeof(ItemRec) * (size_t)nitems));
  if (tile) CM_TRY(w.col_start.ensure(8 * (size_t)(M->n_local + 1)));
  const size_t cb = tile ? cuts_bytes(M->dtype, M->m_local, nitems) : 0;
  if (cb) CM_TRY(w.cuts.ensure(cb));
  unsigned short *cuts = cb ? w.cuts.as<unsigned short>() : nullptr;
  launch_build_recs(M->dtype, w.segs.as<Seg>(), keys, items, nitems, w.recs.as<ItemRec>(), M->n_local,
                    tile ? w.col_start.as<long>() : nullptr, cuts, M->m_local, st);
  M->stats.kernel_launches += 1;
  {
    ScopedTimer t(M, 0, elements, st);
    if (tile)
      launch_apply_tile(M->dtype, w.recs.as<ItemRec>(), w.col_start.as<long>(), nitems, M->d_local, M->n_local,
                        M->m_local, M->g, max_blocks_per_sm, M->deterministic, cuts, st);
    else
      launch_apply_red(M->dtype, w.recs.as<ItemRec>(), nitems, elements, M->d_local, M->g, st);
    M->stats.kernel_launches += 1;
    if (tile) M->stats.tile_launches += 1;
  }
  CM_CUDA_TRY(cudaGetLastError());
  return CM_OK;
}

// local[inds[k]] += vals[k] for a COO blob already in HBM ([inds x n][vals x n])
int apply_coo_blob(cm_matrix *M, const char *blob, long n) {
  const long *inds = reinterpret_cast<const long *>(blob);
  const void *vals = blob + (size_t)n * 8;
  if (M->deterministic) {
    const size_t tmp = coo_det_temp_bytes(n);
    CM_TRY(M->d_coo_tmp.ensure(tmp));
    launch_apply_coo_det(M->dtype, inds, vals, n, M->d_local, M->d_coo_tmp.p, tmp, M->stream);
    M->stats.kernel_launches += 4;
  } else {
    launch_apply_coo(M->dtype, inds, vals, n, M->d_local, M->stream);
    M->stats.kernel_launches += 1;
  }
  M->stats.elements_applied += n;
  M->mirror_valid = false;  // the block changed, whatever the apply policy
  return CM_OK;
}

int apply_coo_local(cm_matrix *M, const CooBin &bin) {
  const long n = (long)bin.inds.size();
  if (n == 0) return CM_OK;
  CM_TRY(M->d_send_coo.ensure((size_t)n * (8 + M->esz)));
  char *d = M->d_send_coo.as<char>();
  CM_CUDA_TRY(cudaMemcpyAsync(d, bin.inds.data(), (size_t)n * 8, cudaMemcpyHostToDevice, M->stream));
  CM_CUDA_TRY(cudaMemcpyAsync(d + (size_t)n * 8, bin.vals.data(), (size_t)n * M->esz, cudaMemcpyHostToDevice, M->stream));
  M->stats.h2d_bytes += n * (long)(8 + M->esz);
  return apply_coo_blob(M, d, n);
}

// ---- single-rank flush: map -> (sort) -> scatter-add straight from the staged payloads ----
// sync = false (threshold flushes of device-resident updates only): the kernels are left running
// on the stream while the host stages the next batch; nothing host-side is reused by them
int flush_direct(cm_matrix *M, bool sync = true) {
  const int nupd = (int)M->descs.size();
  if (nupd == 0 && M->coo_total == 0) {
    if (sync) {
      CM_CUDA_TRY(cudaStreamSynchronize(M->stream));
      resolve_timers(M);
    }
    return CM_OK;
  }
  if (M->payload_arena.used() || M->index_arena.used() || M->coo_total) sync = true;
  if (nupd) {
    trace_mark(M, "flush: begin");
    CM_TRY(upload_staged(M));
    const long nitems = M->staged_items;
    CM_TRY(M->d_item_base.ensure(8 * (size_t)nupd));
    CM_CUDA_TRY(cudaMemcpyAsync(M->d_item_base.p, M->item_base.data(), 8 * (size_t)nupd, cudaMemcpyHostToDevice, M->stream));
    ApplyWork &w = M->work[0];
    CM_TRY(w.segs.ensure(sizeof(Seg) * (size_t)nupd));
    CM_TRY(w.keys0.ensure(4 * (size_t)std::max(nitems, 1L)));
    CM_TRY(w.items0.ensure(8 * (size_t)std::max(nitems, 1L)));
    const MapArenas ma = map_arenas(M);
    {
      ScopedTimer t(M, 1, 0);
      launch_map_group(M->d_descs.as<UpdDesc>(), nupd, M->g, PanelPlan{1, 1}, ma, M->stream);
      launch_build_direct(M->dtype, M->d_descs.as<UpdDesc>(), M->d_item_base.as<long>(), nupd, M->g, ma,
                          w.segs.as<Seg>(), w.keys0.as<uint32_t>(), w.items0.as<uint64_t>(), M->stream);
      M->stats.kernel_launches += 2;
    }
    trace_mark(M, "flush: mapped, items built");
    CM_TRY(sort_and_apply(M, w, M->stream, nitems, M->staged_elems, 0, 0));
    trace_mark(M, "flush: scatter-add done");
    M->stats.elements_applied += M->staged_elems;
  }
  if (M->coo_total) CM_TRY(apply_coo_local(M, M->coo[0]));
  if (sync) {
    CM_CUDA_TRY(cudaStreamSynchronize(M->stream));
    CM_CUDA_TRY(cudaGetLastError());
    resolve_timers(M);
    trace_dump(M, true);
  }
  M->stats.flushes += 1;
  M->mirror_valid = false;
  clear_staged(M);
  return CM_OK;
}

// ---- barriers of a commit ------------------------------------------------------------------
// stream-ordered: work enqueued on `st` afterwards starts only when every rank's stream has reached
// its own call (a 1-element ncclAllReduce; flags over peer memory were measured no faster at 2 and 8 GPUs)
int commit_stream_barrier(cm_matrix *M, cudaStream_t st) { return M->ctx->tp->stream_barrier(st); }
// returns when every rank's stream work up to its own call is complete
int commit_barrier(cm_matrix *M, cudaStream_t st) {
  CM_TRY(commit_stream_barrier(M, st));
  CM_CUDA_TRY(cudaStreamSynchronize(st));
  return CM_OK;
}

// ---- receive arenas ---------------------------------------------------------------------
// Every rank owns a values arena and a metadata arena that hold the messages of ONE commit from
// all sources (its own included). With peer-to-peer exchange the sources' pack kernels store
// straight into them over NVLink, so every rank keeps them mapped (CUDA IPC; plain pointers when
// the ranks share a process). Capacities only grow; because every rank sees the full counts
// matrix, every rank takes the same growth decision for every arena without extra messages.
struct ArenaHandles {
  cudaIpcMemHandle_t vals, meta;
  void *pvals, *pmeta;  // same-process transports
};

size_t grown(size_t need) { return need + need / 4 + (1u << 20); }

int ensure_recv_arenas(cm_matrix *M, const std::vector<size_t> &need_vals, const std::vector<size_t> &need_meta) {
  const int P = M->P, me = M->me;
  Transport *tp = M->ctx->tp;
  if (!M->p2p) {  // only my own arenas matter
    CM_TRY(M->d_recv_vals.ensure(std::max(need_vals[me], (size_t)256)));
    CM_TRY(M->d_recv_meta.ensure(std::max(need_meta[me], (size_t)256)));
    M->peer_vals[me] = M->d_recv_vals.p;
    M->peer_meta[me] = M->d_recv_meta.p;
    return CM_OK;
  }
  bool any = false;
  std::vector<char> grow(P, 0);
  for (int d = 0; d < P; ++d)
    if (need_vals[d] > M->cap_vals[d] || need_meta[d] > M->cap_meta[d]) grow[d] = 1, any = true;
  if (!any) return CM_OK;
  if (getenv("CM_B200_TRACE") && me == 0)
    for (int d = 0; d < P; ++d)
      if (grow[d])
        fprintf(stderr, "[cm trace r0 c%ld] arena of owner %d grows: values %zu -> %zu bytes, metadata %zu -> %zu\n", M->commit_seq, d,
                M->cap_vals[d], need_vals[d], M->cap_meta[d], need_meta[d]);
  // 1. nobody keeps a mapping of an arena that is about to be freed
  for (int d = 0; d < P; ++d)
    if (grow[d] && d != me && M->peer_arena_open[d]) {
      cudaIpcCloseMemHandle(M->peer_vals[d]);
      cudaIpcCloseMemHandle(M->peer_meta[d]);
      M->peer_arena_open[d] = false;
    }
  CM_TRY(tp->barrier(M->stream));
  // 2. owners reallocate
  if (grow[me]) {
    M->d_recv_vals.release();
    M->d_recv_meta.release();
    CM_TRY(M->d_recv_vals.ensure(grown(need_vals[me])));
    CM_TRY(M->d_recv_meta.ensure(grown(need_meta[me])));
  }
  // 3. publish + map
  ArenaHandles mine;
  memset(&mine, 0, sizeof(mine));
  mine.pvals = M->d_recv_vals.p;
  mine.pmeta = M->d_recv_meta.p;
  if (!tp->same_process()) {
    CM_CUDA_TRY(cudaIpcGetMemHandle(&mine.vals, M->d_recv_vals.p));
    CM_CUDA_TRY(cudaIpcGetMemHandle(&mine.meta, M->d_recv_meta.p));
  }
  std::vector<ArenaHandles> all(P);
  CM_TRY(tp->host_allgather(&mine, all.data(), sizeof(mine), M->stream));
  for (int d = 0; d < P; ++d) {
    if (!grow[d]) continue;
    if (d == me || tp->same_process()) {
      M->peer_vals[d] = all[d].pvals;
      M->peer_meta[d] = all[d].pmeta;
    } else {
      CM_CUDA_TRY(cudaIpcOpenMemHandle(&M->peer_vals[d], all[d].vals, cudaIpcMemLazyEnablePeerAccess));
      CM_CUDA_TRY(cudaIpcOpenMemHandle(&M->peer_meta[d], all[d].meta, cudaIpcMemLazyEnablePeerAccess));
      M->peer_arena_open[d] = true;
    }
    // every rank records the same capacity the owner allocated at least
    M->cap_vals[d] = grown(need_vals[d]);
    M->cap_meta[d] = grown(need_meta[d]);
  }
  return CM_OK;
}

// owner side: messages -> Seg table + (key, item) pairs -> sort -> scatter-add
int apply_messages(cm_matrix *M, std::vector<MsgRef> &msgs, long seg_total, long item_total, long elem_total, long adj_total,
                   long max_seg, ApplyWork &w, cudaStream_t st, int max_blocks_per_sm) {
  if (msgs.empty()) return CM_OK;
  CM_TRY(w.msgs.ensure(sizeof(MsgRef) * msgs.size()));
  CM_CUDA_TRY(cudaMemcpyAsync(w.msgs.p, msgs.data(), sizeof(MsgRef) * msgs.size(), cudaMemcpyHostToDevice, st));
  CM_TRY(w.segs.ensure(sizeof(Seg) * (size_t)seg_total));
  CM_TRY(w.keys0.ensure(4 * (size_t)std::max(item_total, 1L)));
  CM_TRY(w.items0.ensure(8 * (size_t)std::max(item_total, 1L)));
  launch_build_recv(M->dtype, w.msgs.as<MsgRef>(), (int)msgs.size(), max_seg, w.segs.as<Seg>(), w.keys0.as<uint32_t>(),
                    w.items0.as<uint64_t>(), st);
  M->stats.kernel_launches += 1;
  CM_TRY(sort_and_apply(M, w, st, item_total, elem_total, adj_total, max_blocks_per_sm));
  M->stats.elements_applied += elem_total;
  return CM_OK;
}

// a list of delivered messages, in order, as ONE scatter-add on stream `st` — or several, when the list holds more
// segments than one scatter-add can address (many threshold flushes between two commits)
int apply_list(cm_matrix *M, const cm_matrix::PendMsg *first, const cm_matrix::PendMsg *last, ApplyWork &w, cudaStream_t st,
               int max_blocks_per_sm) {
  std::vector<MsgRef> msgs;
  long seg_total = 0, item_total = 0, elem_total = 0, adj_total = 0, max_seg = 0;
  for (const cm_matrix::PendMsg *pm = first; pm != last; ++pm) {
    if (pm->nseg > M->max_staged_updates)
      return fail(CM_ERR_INVALID, "a message holds %ld segments; the limit per scatter-add is %ld", pm->nseg, M->max_staged_updates);
    if (seg_total + pm->nseg > M->max_staged_updates) {
      CM_TRY(apply_messages(M, msgs, seg_total, item_total, elem_total, adj_total, max_seg, w, st, max_blocks_per_sm));
      msgs.clear();
      seg_total = item_total = elem_total = adj_total = max_seg = 0;
    }
    MsgRef m;
    m.meta = pm->meta;
    m.vals = pm->vals;
    m.seg_base = seg_total;
    m.item_base = item_total;
    seg_total += pm->nseg;
    item_total += pm->nitems;
    elem_total += pm->nelems;
    adj_total += pm->nadj;
    max_seg = std::max(max_seg, pm->nseg);
    msgs.push_back(m);
  }
  return apply_messages(M, msgs, seg_total, item_total, elem_total, adj_total, max_seg, w, st, max_blocks_per_sm);
}

// a background scatter-add (apply stream) must have finished before anything else touches the local
// block or the work buffers it uses
int wait_background(cm_matrix *M) {
  if (!M->bg_running) return CM_OK;
  CM_CUDA_TRY(cudaStreamWaitEvent(M->stream, M->ev_applied, 0));
  M->bg_running = false;
  return CM_OK;
}

// the un-applied part of my log as one scatter-add on stream `st`
int apply_log(cm_matrix *M, ApplyWork &w, cudaStream_t st, int max_blocks_per_sm) {
  CM_TRY(apply_list(M, M->pend.data() + M->pend_applied, M->pend.data() + M->pend.size(), w, st, max_blocks_per_sm));
  M->pend_applied = M->pend.size();
  M->mirror_valid = false;
  return CM_OK;
}

// deferred policy: apply what has been logged in my arenas and not applied yet (local, not collective;
// the log itself is only truncated by materialize_all, because every rank tracks its length)
int materialize_local(cm_matrix *M) {
  if (M->pend_applied >= M->pend.size() && !M->bg_running) return CM_OK;
  CM_TRY(wait_background(M));
  if (M->pend_applied < M->pend.size()) CM_TRY(apply_log(M, M->work[0], M->stream, 0));
  CM_CUDA_TRY(cudaStreamSynchronize(M->stream));
  resolve_timers(M);
  return CM_OK;
}

// background policy: start the scatter-add of what has been logged on the apply stream and return; the
// pack / exchange of the following commits overlaps it (the log is append-only, nothing it reads changes)
int start_background_apply(cm_matrix *M) {
  long waiting = 0;
  for (size_t i = M->pend_applied; i < M->pend.size(); ++i) waiting += M->pend[i].nelems;
  if (waiting < M->bg_min_elems) return CM_OK;
  // everything this stream has done so far (the delivery barrier included) precedes the scatter-add; an
  // earlier background scatter-add precedes it in apply-stream order
  CM_CUDA_TRY(cudaEventRecord(M->ev_bg, M->stream));
  CM_CUDA_TRY(cudaStreamWaitEvent(M->apply_stream, M->ev_bg, 0));
  trace_mark(M, "background apply: start", M->apply_stream);
  CM_TRY(apply_log(M, M->work[1], M->apply_stream, M->pipeline_apply_blocks));
  trace_mark(M, "background apply: end", M->apply_stream);
  CM_CUDA_TRY(cudaEventRecord(M->ev_applied, M->apply_stream));
  M->bg_running = true;
  return CM_OK;
}

// COLLECTIVE: everybody applies its log, then the logs are truncated everywhere
int materialize_all(cm_matrix *M) {
  if (M->P <= 1 || M->pend_v.empty()) return CM_OK;
  bool any = false;  // from the global offsets only: every rank must take the same branch
  for (int d = 0; d < M->P; ++d) any = any || M->pend_v[d] != 0 || M->pend_m[d] != 0;
  if (!any) return CM_OK;  // same answer on every rank: the offsets are global knowledge
  CM_TRY(materialize_local(M));
  CM_TRY(commit_barrier(M, M->stream));  // nobody overwrites an arena that is still being read
  M->pend.clear();
  M->pend_applied = 0;
  std::fill(M->pend_v.begin(), M->pend_v.end(), 0L);
  std::fill(M->pend_m.begin(), M->pend_m.end(), 0L);
  return CM_OK;
}

// Column panels of this commit (cm_common.h, PanelPlan). Decided by each source on its own from what
// it has staged; the wire format carries whatever it decides.
PanelPlan plan_panels(cm_matrix *M) {
  PanelPlan pp{1, 1};
  const bool wanted = M->P > 1 && M->p2p && M->pipeline == 1;
  // a staged payload that IS the local block (fill_lower) must be packed before any scatter-add of this
  // commit writes the block: one panel, one batch
  if (!wanted || M->apply_policy != 0 || M->reads_local_block || M->staged_elems < M->pipeline_min_elems) return pp;
  long k = std::min<long>(M->npanel_cfg, kMaxChunks);
  k = std::min<long>(k, kMaxGridDim / M->g.npcol);
  k = std::min<long>(k, 1024 / M->P);
  const long ncols_max = cm_numroc(M->g.n, M->g.npcol, 0, M->g.nb);  // owner column 0 holds the most columns
  k = std::min(k, ncols_max);
  if (k <= 1) return pp;
  // whole NB-blocks per panel: then a sorted jx is non-decreasing in panel number across ALL owners
  // (global block B -> local block B / NPCOL -> panel), which is what keeps the wire order the reference's
  const long nb = M->g.nb;
  pp.panel_cols = (((ncols_max + k - 1) / k + nb - 1) / nb) * nb;
  pp.npanel = (int)std::min<long>(k, (ncols_max + pp.panel_cols - 1) / pp.panel_cols);
  if (pp.npanel <= 1) return PanelPlan{1, 1};
  return pp;
}

// ---- general commit path (also used for P == 1 when the wire is being captured, so the pack
// kernels are testable on one rank), in four stages:
//   commit_counts   map -> scan -> counts all-gather -> read-back (the one host sync)
//   commit_layout   where every message of every rank goes; arenas grown; elemental updates staged
//   commit_deliver  pack (+ all-to-all-v) and, panel by panel, the owners' scatter-adds
//   commit_finish   wire capture, log or scatter-add, elemental updates, quiescence barrier
// The pack kernel writes every message where the owner will read it: with peer-to-peer exchange
// that is the owner's receive arena itself (pack and exchange are ONE kernel whose stores cross
// NVLink); otherwise a local send arena that a grouped ncclSend/ncclRecv all-to-all-v moves.
struct CommitPlan {
  int P = 1, me = 0, K = kMaxChunks, nupd = 0;
  PanelPlan pp{1, 1};
  long max_rows = 0;
  MapArenas ma{};
  ScanArenas sa{};
  MsgCounts *counts = nullptr;  // host copy of the all-gathered counts [src][panel][dst]
  long *coo_counts = nullptr;   // [src][dst]
  bool empty = false;           // nothing staged on any rank
  bool deferred = false, pipelined = false;
  int Kp = 0;                   // panel slots that hold data on any rank (same value everywhere)
  std::vector<long> rbase_v, rbase_m;  // [owner][panel * P + src]: where each message starts (elements / bytes)
  std::vector<void *> dst_ptrs;        // host copy of sa.dst_vals / sa.dst_meta
  std::vector<long> r_coo_base, s_coo_base;
  cudaEvent_t ev_pack0 = nullptr, ev_pack1 = nullptr;
  // messages packed by threshold flushes since the last commit, of EVERY rank (source-major, in flush order)
  std::vector<FlushDesc> fl;
  std::vector<long> fl_pos_v, fl_pos_m;  // !remote: where the message goes in its owner's receive arena (elements / bytes)
  std::vector<long> end_v, end_m;        // [owner] end of everything this commit puts into that owner's receive arenas
  int fl_lanes = 0;                      // without peer-to-peer stores: transport lanes that carry them (2 per message of a pair)

  size_t cidx(int src, int k, int dst) const { return ((size_t)src * K + k) * P + dst; }
  size_t ridx(int d, int k, int src) const { return (size_t)d * (K * P + 1) + (size_t)k * P + src; }
  long coo_blob(size_t esz, long n) const { return n * 8 + (long)(((size_t)n * esz + 7) & ~(size_t)7); }
  bool panel_has_data(int k) const {  // does this rank send anything in panel k?
    if (k >= pp.npanel || nupd == 0) return false;
    for (int d = 0; d < P; ++d)
      if (counts[cidx(me, k, d)].nvals) return true;
    return false;
  }
};

// owner map + grouping + per-(panel, owner) sizes of the staged updates, on the main stream
int map_and_scan(cm_matrix *M, const PanelPlan &pp, int nupd, MapArenas &ma, ScanArenas &sa) {
  const int P = M->P, K = kMaxChunks;
  CM_TRY(upload_staged(M, pp.npanel));
  const size_t pu = (size_t)pp.npanel * P * std::max(nupd, 1);
  CM_TRY(M->d_valoff.ensure(8 * pu));
  CM_TRY(M->d_idxoff.ensure(8 * pu));
  CM_TRY(M->d_itemoff.ensure(8 * pu));
  CM_TRY(M->d_totals.ensure(sizeof(MsgCounts) * (size_t)P * K));
  CM_TRY(M->d_dst_ptrs.ensure(2 * sizeof(void *) * (size_t)P * K));
  ma = map_arenas(M);
  sa.valoff = M->d_valoff.as<long>();
  sa.idxoff = M->d_idxoff.as<long>();
  sa.itemoff = M->d_itemoff.as<long>();
  sa.totals = M->d_totals.as<MsgCounts>();
  sa.dst_vals = M->d_dst_ptrs.as<void *>();
  sa.dst_meta = reinterpret_cast<char **>(M->d_dst_ptrs.as<void *>() + (size_t)P * K);
  // host-known per-owner counts ride in the same totals (one tiny H2D): elemental updates and the
  // messages threshold flushes have packed since the last commit
  std::vector<long> extra(2 * (size_t)P, 0);
  for (int d = 0; d < P; ++d) extra[d] = (long)M->coo[d].inds.size();
  for (const FlushedMsg &f : M->flushed) extra[(size_t)P + f.d.dst] += 1;
  CM_TRY(M->d_coo_counts.ensure(16 * (size_t)P));
  CM_CUDA_TRY(cudaMemcpyAsync(M->d_coo_counts.p, extra.data(), 16 * (size_t)P, cudaMemcpyHostToDevice, M->stream));
  launch_map_group(M->d_descs.as<UpdDesc>(), nupd, M->g, pp, ma, M->stream);
  launch_scan(nupd, pp, M->g, ma, sa, M->d_coo_counts.as<long>(), M->stream);
  M->stats.kernel_launches += (nupd ? 1 : 0) + 1;
  CM_CUDA_TRY(cudaGetLastError());
  return CM_OK;
}

long pad32(long nvals) { return (nvals + 31) & ~31L; }  // every message starts on a 256-byte boundary

// -- source side: owner map + per-(panel, owner) sizes; counts all-gathered and read back ----------
int commit_counts(cm_matrix *M, CommitPlan &cp) {
  const int P = cp.P, K = cp.K, nupd = cp.nupd;
  Transport *tp = M->ctx->tp;
  // panel plan (a LOCAL decision of this source; owners cope with any mix): a source that stages
  // a large commit cuts its updates' columns into panels so that the owners can pipeline it
  cp.pp = plan_panels(M);
  for (const UpdDesc &d : M->descs) cp.max_rows = std::max<long>(cp.max_rows, d.m);
  if (M->profiling) {
    cp.ev_pack0 = get_event(M);
    cp.ev_pack1 = get_event(M);
    cudaEventRecord(cp.ev_pack0, M->stream);
  }
  CM_TRY(map_and_scan(M, cp.pp, nupd, cp.ma, cp.sa));
  trace_mark(M, "counts: mapped + scanned");

  const size_t per_rank = sizeof(MsgCounts) * (size_t)P * K;
  CM_TRY(M->d_counts_all.ensure(per_rank * P));
  CM_TRY(tp->allgather(M->d_totals.p, M->d_counts_all.p, per_rank, M->stream));
  CM_TRY(M->h_counts.ensure(per_rank * P + 8 * (size_t)P * P));
  cp.counts = static_cast<MsgCounts *>(M->h_counts.p);
  CM_CUDA_TRY(cudaMemcpyAsync(cp.counts, M->d_counts_all.p, per_rank * P, cudaMemcpyDeviceToHost, M->stream));
  M->stats.d2h_bytes += (long)(per_rank * P);
  trace_mark(M, "counts: all-gathered");
  CM_CUDA_TRY(cudaStreamSynchronize(M->stream));
  cp.coo_counts = reinterpret_cast<long *>(static_cast<char *>(M->h_counts.p) + per_rank * P);
  for (int src = 0; src < P; ++src)
    for (int d = 0; d < P; ++d) cp.coo_counts[(size_t)src * P + d] = cp.counts[cp.cidx(src, 0, d)].ncoo;
  bool any = false;
  for (size_t i = 0; i < (size_t)P * K * P && !any; ++i)
    any = cp.counts[i].nvals != 0 || cp.counts[i].ncoo != 0 || cp.counts[i].nflushed != 0;
  cp.empty = !any;
  // messages of threshold flushes (rare): everybody learns every message's owner, size and place
  long nmax = 0;
  for (int src = 0; src < P; ++src) {
    long n = 0;
    for (int d = 0; d < P; ++d) n += cp.counts[cp.cidx(src, 0, d)].nflushed;
    nmax = std::max(nmax, n);
  }
  if (nmax > 0) {
    std::vector<FlushDesc> mine((size_t)nmax), all((size_t)nmax * P);
    for (auto &f : mine) f.src = -1;
    for (size_t i = 0; i < M->flushed.size(); ++i) mine[i] = M->flushed[i].d;
    CM_TRY(tp->host_allgather(mine.data(), all.data(), sizeof(FlushDesc) * (size_t)nmax, M->stream));
    for (const FlushDesc &f : all)
      if (f.src >= 0) cp.fl.push_back(f);
  }
  return CM_OK;
}

// -- where everything goes (identical computation on every rank) ------------------------------------
int commit_layout(cm_matrix *M, CommitPlan &cp) {
  const int P = cp.P, me = cp.me, K = cp.K;
  const size_t esz = M->esz;
  const MsgCounts *counts = cp.counts;
  // pipeline or not: same decision on every rank, taken from the global counts. Pipelining pays
  // when the commit is large and most of it arrived cut into panels (panel k > 0 non-empty);
  // everything else goes as one batch.
  cp.deferred = M->apply_policy != 0 && P > 1 && M->p2p;
  if (!cp.deferred && P > 1 && M->p2p && M->pipeline == 1) {
    long total = 0, later = 0;
    for (int src = 0; src < P; ++src)
      for (int k = 0; k < K; ++k)
        for (int d = 0; d < P; ++d) {
          total += counts[cp.cidx(src, k, d)].nelems;
          if (k > 0) later += counts[cp.cidx(src, k, d)].nelems;
        }
    // ... and the busiest owner receives enough for its scatter-add to be worth hiding: every panel costs a
    // stream-ordered barrier and a round of small kernels (B200, 8 GPUs: K2's commits of 67 Mi elements per owner
    // 1.98 ms pipelined against 1.55 ms as one batch; NS, 268 Mi per owner, 6.05 against 6.26 ms; K1-weak, 655 Mi,
    // 9.7 against 10.8 ms; K3's hot owner receives 243 Mi per commit)
    long busiest = 0;
    for (int d = 0; d < P; ++d) {
      long got = 0;
      for (int src = 0; src < P; ++src)
        for (int k = 0; k < K; ++k) got += counts[cp.cidx(src, k, d)].nelems;
      busiest = std::max(busiest, got);
    }
    cp.pipelined = total / P >= M->pipeline_min_elems && 2 * later >= total && busiest >= M->pipeline_owner_min_elems;
  }
  if (!cp.fl.empty()) cp.pipelined = false;  // flushed messages are applied with the commit's own, as one batch
  for (int src = 0; src < P; ++src)
    for (int k = 0; k < K; ++k)
      for (int d = 0; d < P; ++d)
        if (counts[cp.cidx(src, k, d)].nvals) cp.Kp = std::max(cp.Kp, k + 1);

  // layout of every rank's receive arenas: owner d holds, panel after panel, the messages of every source
  cp.rbase_v.assign((size_t)P * (K * P + 1), 0);
  cp.rbase_m.assign((size_t)P * (K * P + 1), 0);
  cp.fl_pos_v.assign(cp.fl.size(), 0);
  cp.fl_pos_m.assign(cp.fl.size(), 0);
  cp.end_v.assign(P, 0);
  cp.end_m.assign(P, 0);
  std::vector<size_t> need_vals(P), need_meta(P);
  auto lay_out = [&]() {  // eager: from the start of the arenas; deferred: behind what is logged there
    for (int d = 0; d < P; ++d) {
      cp.rbase_v[cp.ridx(d, 0, 0)] = cp.deferred ? M->pend_v[d] : 0;
      cp.rbase_m[cp.ridx(d, 0, 0)] = cp.deferred ? M->pend_m[d] : 0;
      for (int k = 0; k < K; ++k)
        for (int src = 0; src < P; ++src) {
          const MsgCounts &c = counts[cp.cidx(src, k, d)];
          const size_t i = cp.ridx(d, k, src);
          cp.rbase_v[i + 1] = cp.rbase_v[i] + ((c.nvals + 31) & ~31L);  // every message on a 256-byte boundary
          cp.rbase_m[i + 1] = cp.rbase_m[i] + meta_bytes(c.nseg, c.nidx);
        }
      // behind them: the flushed messages that still sit at their sources
      long fv = cp.rbase_v[cp.ridx(d, K - 1, P - 1) + 1], fm = cp.rbase_m[cp.ridx(d, K - 1, P - 1) + 1];
      for (size_t i = 0; i < cp.fl.size(); ++i) {
        const FlushDesc &f = cp.fl[i];
        if (f.dst != d || f.remote) continue;
        cp.fl_pos_v[i] = fv;
        cp.fl_pos_m[i] = fm;
        fv += pad32(f.c.nvals);
        fm += meta_bytes(f.c.nseg, f.c.nidx);
      }
      cp.end_v[d] = fv;
      cp.end_m[d] = fm;
      need_vals[d] = (size_t)fv * esz;
      need_meta[d] = (size_t)fm;
    }
  };
  lay_out();
  if (cp.deferred) {
    // an arena that has to grow is reallocated: everybody applies and truncates its log first (the
    // capacities are tracked identically on every rank, so this is the same decision everywhere)
    bool grow = false;
    for (int d = 0; d < P; ++d) grow = grow || need_vals[d] > M->cap_vals[d] || need_meta[d] > M->cap_meta[d];
    if (grow) {
      // the new arenas hold what the log WOULD have needed, not just this commit: otherwise a log that is
      // truncated at every growth never reaches the size a step needs and keeps growing a little each time
      const std::vector<size_t> log_vals = need_vals, log_meta = need_meta;
      CM_TRY(materialize_all(M));
      lay_out();
      // ... and twice that, for every owner at once: a growth costs a collective reallocation and re-mapping of
      // gigabytes (hundreds of milliseconds); the log of a step must fit after a couple of them
      size_t top_v = 0, top_m = 0;
      for (int d = 0; d < P; ++d) {
        top_v = std::max(top_v, log_vals[d]);
        top_m = std::max(top_m, log_meta[d]);
      }
      for (int d = 0; d < P; ++d) {
        need_vals[d] = std::max(need_vals[d], 2 * top_v);
        need_meta[d] = std::max(need_meta[d], 2 * top_m);
      }
    }
  }
  CM_TRY(ensure_recv_arenas(M, need_vals, need_meta));

  // where my messages go: peer-to-peer = straight into the owners' arenas; otherwise my own message
  // into my arena and the others into a local send arena (moved by the all-to-all-v)
  std::vector<long> s_val_base((size_t)K * P + 1, 0), s_meta_base((size_t)K * P + 1, 0);
  for (int k = 0; k < K; ++k)
    for (int d = 0; d < P; ++d) {
      const MsgCounts &c = counts[cp.cidx(me, k, d)];
      const bool staged = !M->p2p && d != me;
      const size_t i = (size_t)k * P + d;
      s_val_base[i + 1] = s_val_base[i] + (staged ? ((c.nvals + 31) & ~31L) : 0);
      s_meta_base[i + 1] = s_meta_base[i] + (staged ? meta_bytes(c.nseg, c.nidx) : 0);
    }
  if (!M->p2p) {
    CM_TRY(M->d_send_vals.ensure(std::max((size_t)s_val_base[(size_t)K * P] * esz, (size_t)256)));
    CM_TRY(M->d_send_meta.ensure(std::max((size_t)s_meta_base[(size_t)K * P], (size_t)256)));
  }
  cp.dst_ptrs.resize(2 * (size_t)K * P);
  for (int k = 0; k < K; ++k)
    for (int d = 0; d < P; ++d) {
      const size_t i = (size_t)k * P + d;
      if (M->p2p || d == me) {
        cp.dst_ptrs[i] = static_cast<char *>(M->peer_vals[d]) + (size_t)cp.rbase_v[cp.ridx(d, k, me)] * esz;
        cp.dst_ptrs[(size_t)K * P + i] = static_cast<char *>(M->peer_meta[d]) + cp.rbase_m[cp.ridx(d, k, me)];
      } else {
        cp.dst_ptrs[i] = M->d_send_vals.as<char>() + (size_t)s_val_base[i] * esz;
        cp.dst_ptrs[(size_t)K * P + i] = M->d_send_meta.as<char>() + s_meta_base[i];
      }
    }
  CM_CUDA_TRY(cudaMemcpyAsync(M->d_dst_ptrs.p, cp.dst_ptrs.data(), cp.dst_ptrs.size() * sizeof(void *), cudaMemcpyHostToDevice,
                              M->stream));

  // COO blobs (elemental updates): [inds x n][vals x n] per owner, always via the transport
  cp.r_coo_base.assign(P + 1, 0);
  cp.s_coo_base.assign(P + 1, 0);
  for (int r = 0; r < P; ++r) {
    cp.r_coo_base[r + 1] = cp.r_coo_base[r] + (r != me ? cp.coo_blob(esz, cp.coo_counts[(size_t)r * P + me]) : 0);
    cp.s_coo_base[r + 1] = cp.s_coo_base[r] + cp.coo_blob(esz, (long)M->coo[r].inds.size());
  }
  CM_TRY(M->d_recv_coo.ensure(std::max((size_t)cp.r_coo_base[P], (size_t)256)));
  CM_TRY(M->d_send_coo.ensure(std::max((size_t)cp.s_coo_base[P], (size_t)256)));
  for (int d = 0; d < P; ++d) {
    const long n = (long)M->coo[d].inds.size();
    if (!n) continue;
    char *dst = M->d_send_coo.as<char>() + cp.s_coo_base[d];
    CM_CUDA_TRY(cudaMemcpyAsync(dst, M->coo[d].inds.data(), (size_t)n * 8, cudaMemcpyHostToDevice, M->stream));
    CM_CUDA_TRY(cudaMemcpyAsync(dst + (size_t)n * 8, M->coo[d].vals.data(), (size_t)n * esz, cudaMemcpyHostToDevice,
                                M->stream));
    M->stats.h2d_bytes += n * (long)(8 + esz);
  }
  for (int r = 0; r < P; ++r) {
    if (r == me) continue;
    for (int k = 0; k < K; ++k) {
      const MsgCounts &cs = counts[cp.cidx(me, k, r)], &cr = counts[cp.cidx(r, k, me)];
      M->stats.bytes_sent += (long)(cs.nvals * esz + (cs.nvals ? meta_bytes(cs.nseg, cs.nidx) : 0));
      M->stats.bytes_received += (long)(cr.nvals * esz + (cr.nvals ? meta_bytes(cr.nseg, cr.nidx) : 0));
    }
    M->stats.bytes_sent += cp.coo_blob(esz, (long)M->coo[r].inds.size()) * (M->coo[r].inds.empty() ? 0 : 1);
    M->stats.bytes_received += cp.coo_blob(esz, cp.coo_counts[(size_t)r * P + me]) * (cp.coo_counts[(size_t)r * P + me] ? 1 : 0);
  }
  // flushed messages: what went out at flush time (remote) was counted then
  std::vector<int> per_pair((size_t)P * P, 0);
  for (const FlushDesc &f : cp.fl) {
    const long bytes = (long)(f.c.nvals * esz) + meta_bytes(f.c.nseg, f.c.nidx);
    if (f.src == me && f.dst != me && !f.remote) M->stats.bytes_sent += bytes;
    if (f.dst == me && f.src != me) M->stats.bytes_received += bytes;
    if (!f.remote && f.src != f.dst) cp.fl_lanes = std::max(cp.fl_lanes, 2 * ++per_pair[(size_t)f.src * P + f.dst]);
  }
  return CM_OK;
}

// where a flushed message for THIS rank sits: in my flush region of its source, or in my receive arena
cm_matrix::PendMsg flushed_ref(cm_matrix *M, const CommitPlan &cp, size_t i) {
  const FlushDesc &f = cp.fl[i];
  const char *meta, *vals;
  if (f.remote) {
    meta = M->d_flush_meta.as<char>() + (size_t)f.src * M->flush_region_m + f.off_m;
    vals = M->d_flush_vals.as<char>() + (size_t)f.src * M->flush_region_v + f.off_v;
  } else {
    meta = static_cast<char *>(M->peer_meta[cp.me]) + cp.fl_pos_m[i];
    vals = static_cast<char *>(M->peer_vals[cp.me]) + (size_t)cp.fl_pos_v[i] * M->esz;
  }
  return cm_matrix::PendMsg{meta, vals, f.c.nseg, f.c.nitems, f.c.nelems, f.c.nadj};
}

// what this commit delivered to THIS rank, in stream order per source: the messages of threshold flushes
// (with_flushed), then the commit's own messages of the panels [k0, k1)
void delivered_messages(cm_matrix *M, const CommitPlan &cp, int k0, int k1, bool with_flushed, std::vector<cm_matrix::PendMsg> &out) {
  if (with_flushed)
    for (size_t i = 0; i < cp.fl.size(); ++i)
      if (cp.fl[i].dst == cp.me) out.push_back(flushed_ref(M, cp, i));
  for (int k = k0; k < k1; ++k)
    for (int src = 0; src < cp.P; ++src) {
      const MsgCounts &c = cp.counts[cp.cidx(src, k, cp.me)];
      if (c.nvals == 0) continue;
      out.push_back(cm_matrix::PendMsg{static_cast<char *>(M->peer_meta[cp.me]) + cp.rbase_m[cp.ridx(cp.me, k, src)],
                                       static_cast<char *>(M->peer_vals[cp.me]) + (size_t)cp.rbase_v[cp.ridx(cp.me, k, src)] * M->esz,
                                       c.nseg, c.nitems, c.nelems, c.nadj});
    }
}

// owner side of the panels [k0, k1): their messages -> one scatter-add on stream `st`
int apply_panels(cm_matrix *M, const CommitPlan &cp, int k0, int k1, bool with_flushed, ApplyWork &w, cudaStream_t st,
                 int max_blocks_per_sm) {
  std::vector<cm_matrix::PendMsg> msgs;
  delivered_messages(M, cp, k0, k1, with_flushed, msgs);
  return apply_list(M, msgs.data(), msgs.data() + msgs.size(), w, st, max_blocks_per_sm);
}

// exchange lanes of the transport: values + metadata of the panels [k0, k1) without peer-to-peer stores,
// elemental updates (COO) always
int transport_lanes(cm_matrix *M, const CommitPlan &cp, int k0, int k1, bool with_coo) {
  const int P = cp.P, me = cp.me, K = cp.K;
  const size_t esz = M->esz;
  const MsgCounts *counts = cp.counts;
  bool global_lane = false;  // every rank must take the same branch: decide from the global counts
  for (int src = 0; src < P && !global_lane; ++src)
    for (int d = 0; d < P && !global_lane; ++d) {
      if (src == d) continue;
      if (with_coo && cp.coo_counts[(size_t)src * P + d]) global_lane = true;
      if (!M->p2p)
        for (int k = k0; k < k1; ++k)
          if (counts[cp.cidx(src, k, d)].nvals) global_lane = true;
    }
  const int nfl = (with_coo && !M->p2p) ? cp.fl_lanes : 0;  // flushed messages still at their sources ride along
  if (nfl) global_lane = true;
  if (!global_lane) return CM_OK;  // peer-to-peer stores without elemental updates: nothing rides on the transport
  const int nl = 2 * (k1 - k0) + 1 + nfl;
  std::vector<std::vector<XferSpec>> sends(nl, std::vector<XferSpec>(P, XferSpec{nullptr, 0}));
  std::vector<std::vector<XferSpec>> recvs(nl, std::vector<XferSpec>(P, XferSpec{nullptr, 0}));
  for (int r = 0; r < P; ++r) {
    if (r == me) continue;
    if (!M->p2p)
      for (int k = k0; k < k1; ++k) {
        const MsgCounts &cs = counts[cp.cidx(me, k, r)];
        const MsgCounts &cr = counts[cp.cidx(r, k, me)];
        const size_t i = (size_t)k * P + r;
        const int l = 2 * (k - k0);
        sends[l][r] = XferSpec{cp.dst_ptrs[i], (size_t)cs.nvals * esz};
        sends[l + 1][r] = XferSpec{cp.dst_ptrs[(size_t)K * P + i], cs.nvals ? (size_t)meta_bytes(cs.nseg, cs.nidx) : 0};
        recvs[l][r] = XferSpec{static_cast<char *>(M->peer_vals[me]) + (size_t)cp.rbase_v[cp.ridx(me, k, r)] * esz,
                               (size_t)cr.nvals * esz};
        recvs[l + 1][r] = XferSpec{static_cast<char *>(M->peer_meta[me]) + cp.rbase_m[cp.ridx(me, k, r)],
                                   cr.nvals ? (size_t)meta_bytes(cr.nseg, cr.nidx) : 0};
      }
    if (with_coo) {
      if (M->coo[r].inds.size())
        sends[nl - nfl - 1][r] = XferSpec{M->d_send_coo.as<char>() + cp.s_coo_base[r], (size_t)cp.coo_blob(esz, (long)M->coo[r].inds.size())};
      if (cp.coo_counts[(size_t)r * P + me])
        recvs[nl - nfl - 1][r] = XferSpec{M->d_recv_coo.as<char>() + cp.r_coo_base[r], (size_t)cp.coo_blob(esz, cp.coo_counts[(size_t)r * P + me])};
    }
  }
  if (nfl) {  // the j-th message a pair (source, owner) still holds at the source: lanes 2j (values), 2j + 1 (metadata)
    std::vector<int> nth((size_t)P * P, 0);
    const int base = nl - nfl;
    for (size_t i = 0; i < cp.fl.size(); ++i) {
      const FlushDesc &f = cp.fl[i];
      if (f.remote || f.src == f.dst) continue;
      const int j = nth[(size_t)f.src * P + f.dst]++;
      const size_t vb = (size_t)f.c.nvals * esz, mb = (size_t)meta_bytes(f.c.nseg, f.c.nidx);
      if (f.src == me) {
        const FlushedMsg &fm = M->flushed[f.seq];
        sends[base + 2 * j][f.dst] = XferSpec{fm.vals.p, vb};
        sends[base + 2 * j + 1][f.dst] = XferSpec{fm.meta.p, mb};
      } else if (f.dst == me) {
        recvs[base + 2 * j][f.src] = XferSpec{static_cast<char *>(M->peer_vals[me]) + (size_t)cp.fl_pos_v[i] * esz, vb};
        recvs[base + 2 * j + 1][f.src] = XferSpec{static_cast<char *>(M->peer_meta[me]) + cp.fl_pos_m[i], mb};
      }
    }
  }
  ScopedTimer t(M, 2, 0);
  return M->ctx->tp->alltoallv(sends, recvs, M->stream);
}

// -- pack / exchange, and for a pipelined commit the owners' scatter-adds panel by panel ------------
int commit_deliver(cm_matrix *M, CommitPlan &cp) {
  const int P = cp.P, K = cp.K, nupd = cp.nupd;
  if (!cp.pipelined) {
    // one batch: pack everything, (all-to-all-v), barrier; the scatter-add follows in commit_finish
    launch_pack(M->dtype, M->d_descs.as<UpdDesc>(), nupd, 0, cp.pp.npanel, cp.pp, M->g, cp.ma, cp.sa, cp.max_rows, M->stream);
    M->stats.kernel_launches += nupd ? 1 : 0;
    if (M->profiling) {
      cudaEventRecord(cp.ev_pack1, M->stream);
      M->timers.push_back(PendingTimer{1, cp.ev_pack0, cp.ev_pack1, 0});
    }
    CM_CUDA_TRY(cudaGetLastError());
    trace_mark(M, "deliver: packed");
    // flushed messages that still sit here: into their owners' arenas (peer copies), or over the transport below
    for (size_t i = 0; i < cp.fl.size(); ++i) {
      const FlushDesc &f = cp.fl[i];
      if (f.src != cp.me || f.remote || !(M->p2p || f.dst == cp.me)) continue;
      const FlushedMsg &fm = M->flushed[f.seq];
      CM_CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(M->peer_vals[f.dst]) + (size_t)cp.fl_pos_v[i] * M->esz, fm.vals.p,
                                  (size_t)f.c.nvals * M->esz, cudaMemcpyDefault, M->stream));
      CM_CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(M->peer_meta[f.dst]) + cp.fl_pos_m[i], fm.meta.p,
                                  (size_t)meta_bytes(f.c.nseg, f.c.nidx), cudaMemcpyDefault, M->stream));
    }
    CM_TRY(transport_lanes(M, cp, 0, K, true));
    // peer-to-peer: every source's pack kernel has finished storing into my arenas
    if (M->p2p && P > 1) CM_TRY(commit_stream_barrier(M, M->stream));
    trace_mark(M, "deliver: everybody packed");
    return CM_OK;
  }
  // pipelined (peer-to-peer stores): panel k is packed straight into the owners' arenas, a stream-ordered
  // barrier says "everybody's panel k has landed", and its scatter-add (which touches only panel
  // k's columns of the local block) runs on the apply stream while panel k+1 is being packed and
  // stored across NVLink on the main stream
  CM_TRY(transport_lanes(M, cp, 0, 0, true));  // COO lane, if any
  for (int k = 0; k < cp.Kp; ++k) {
    if (cp.panel_has_data(k)) {
      launch_pack(M->dtype, M->d_descs.as<UpdDesc>(), nupd, k, k + 1, cp.pp, M->g, cp.ma, cp.sa, cp.max_rows, M->stream);
      M->stats.kernel_launches += 1;
    }
    trace_mark(M, "deliver: panel packed");
    // "everybody's panel k has landed" is awaited on the apply stream: the main stream goes straight
    // on to the pack of panel k+1
    CM_CUDA_TRY(cudaEventRecord(M->ev_panel[k], M->stream));
    CM_CUDA_TRY(cudaStreamWaitEvent(M->apply_stream, M->ev_panel[k], 0));
    CM_TRY(commit_stream_barrier(M, M->apply_stream));
    trace_mark(M, "deliver: panel landed", M->apply_stream);
    // all but the last panel's scatter-add share the SMs with the next panel's pack kernel
    M->apply_share = 1.0 / cp.Kp;
    const int rc = apply_panels(M, cp, k, k + 1, false, M->work[k], M->apply_stream, k + 1 < cp.Kp ? M->pipeline_apply_blocks : 0);
    M->apply_share = 1.0;
    CM_TRY(rc);
    trace_mark(M, "deliver: panel applied", M->apply_stream);
  }
  if (M->profiling) {
    cudaEventRecord(cp.ev_pack1, M->stream);
    M->timers.push_back(PendingTimer{1, cp.ev_pack0, cp.ev_pack1, 0});
  }
  CM_CUDA_TRY(cudaEventRecord(M->ev_applied, M->apply_stream));
  CM_CUDA_TRY(cudaStreamWaitEvent(M->stream, M->ev_applied, 0));
  return CM_OK;
}

// what THIS rank sent to every owner — flushed messages first, then the commit's own, panel after panel —
// read back from wherever pack put it (parity tests)
int capture_wire(cm_matrix *M, const CommitPlan &cp) {
  CM_CUDA_TRY(cudaStreamSynchronize(M->stream));
  M->dbg_msgs.clear();
  int group = 0;
  for (const FlushedMsg &fm : M->flushed) {
    const FlushDesc &f = fm.d;
    cm_matrix::DbgMsg dm;
    dm.dst = f.dst;
    dm.group = group++;
    dm.meta.resize((size_t)meta_bytes(f.c.nseg, f.c.nidx));
    dm.vals.resize((size_t)f.c.nvals * M->esz);
    const char *meta = f.remote ? static_cast<char *>(M->peer_flush_meta[f.dst]) + (size_t)cp.me * M->flush_region_m + f.off_m
                                : fm.meta.as<char>();
    const char *vals = f.remote ? static_cast<char *>(M->peer_flush_vals[f.dst]) + (size_t)cp.me * M->flush_region_v + f.off_v
                                : fm.vals.as<char>();
    CM_CUDA_TRY(cudaMemcpy(dm.meta.data(), meta, dm.meta.size(), cudaMemcpyDefault));
    CM_CUDA_TRY(cudaMemcpy(dm.vals.data(), vals, dm.vals.size(), cudaMemcpyDefault));
    M->dbg_msgs.push_back(std::move(dm));
  }
  for (int d = 0; d < cp.P; ++d)
    for (int k = 0; k < cp.K; ++k) {
      const MsgCounts &c = cp.counts[cp.cidx(cp.me, k, d)];
      cm_matrix::DbgMsg dm;
      dm.dst = d;
      dm.group = group;
      if (c.nvals) {
        const size_t i = (size_t)k * cp.P + d;
        dm.meta.resize((size_t)meta_bytes(c.nseg, c.nidx));
        dm.vals.resize((size_t)c.nvals * M->esz);
        CM_CUDA_TRY(cudaMemcpy(dm.meta.data(), cp.dst_ptrs[(size_t)cp.K * cp.P + i], dm.meta.size(), cudaMemcpyDefault));
        CM_CUDA_TRY(cudaMemcpy(dm.vals.data(), cp.dst_ptrs[i], dm.vals.size(), cudaMemcpyDefault));
      }
      M->dbg_msgs.push_back(std::move(dm));
    }
  M->dbg_coo = M->coo;
  M->dbg_valid = true;
  return CM_OK;
}

// ---- flush regions -------------------------------------------------------------------------------
// Every owner holds, for every source, a region its threshold flushes store into over NVLink between two
// commits without any rendezvous (the reference's rput into a buffer the owner allocated by RPC, :163-173,
// :226-228; here the buffers exist beforehand). Allocated and mapped COLLECTIVELY, at the end of the first
// commit that carried flushed messages (until then, and whenever a region is full, flushed messages wait in
// buffers of the source): room for four flushes of the largest size seen per (owner, source) pair, at most
// 8 GiB per owner. Same decision on every rank: it is taken from the all-gathered descriptors.
int ensure_flush_regions(cm_matrix *M, const CommitPlan &cp) {
  if (cp.fl.empty() || !M->p2p || M->P <= 1 || M->apply_policy != 0) return CM_OK;
  const int P = M->P, me = M->me;
  Transport *tp = M->ctx->tp;
  size_t top_v = 0, top_m = 0;
  for (const FlushDesc &f : cp.fl) {
    top_v = std::max(top_v, (size_t)pad32(f.c.nvals) * M->esz);
    top_m = std::max(top_m, (size_t)meta_bytes(f.c.nseg, f.c.nidx));
  }
  const size_t cap_v = ((size_t)8 << 30) / P;
  size_t want_v = std::min(cap_v, (4 * top_v + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1));
  size_t want_m = (4 * top_m + 65535) & ~(size_t)65535;
  if (want_v <= M->flush_region_v && want_m <= M->flush_region_m) return CM_OK;
  want_v = std::max(want_v, M->flush_region_v);
  want_m = std::max(want_m, M->flush_region_m);
  // nobody keeps a mapping of regions that are about to be freed
  for (int d = 0; d < P; ++d)
    if (M->peer_flush_open[d]) {
      cudaIpcCloseMemHandle(M->peer_flush_vals[d]);
      cudaIpcCloseMemHandle(M->peer_flush_meta[d]);
      M->peer_flush_open[d] = false;
    }
  CM_TRY(tp->barrier(M->stream));
  CM_TRY(M->d_flush_vals.alloc_exact(want_v * P));
  CM_TRY(M->d_flush_meta.alloc_exact(want_m * P));
  ArenaHandles mine;
  memset(&mine, 0, sizeof(mine));
  mine.pvals = M->d_flush_vals.p;
  mine.pmeta = M->d_flush_meta.p;
  if (!tp->same_process()) {
    CM_CUDA_TRY(cudaIpcGetMemHandle(&mine.vals, M->d_flush_vals.p));
    CM_CUDA_TRY(cudaIpcGetMemHandle(&mine.meta, M->d_flush_meta.p));
  }
  std::vector<ArenaHandles> all(P);
  CM_TRY(tp->host_allgather(&mine, all.data(), sizeof(mine), M->stream));
  for (int d = 0; d < P; ++d) {
    if (d == me || tp->same_process()) {
      M->peer_flush_vals[d] = all[d].pvals;
      M->peer_flush_meta[d] = all[d].pmeta;
    } else {
      CM_CUDA_TRY(cudaIpcOpenMemHandle(&M->peer_flush_vals[d], all[d].vals, cudaIpcMemLazyEnablePeerAccess));
      CM_CUDA_TRY(cudaIpcOpenMemHandle(&M->peer_flush_meta[d], all[d].meta, cudaIpcMemLazyEnablePeerAccess));
      M->peer_flush_open[d] = true;
    }
  }
  M->flush_region_v = want_v;
  M->flush_region_m = want_m;
  return tp->barrier(M->stream);
}

// collective: nobody maps or uses the flush regions any more (exchange mode change, destruction)
void release_flush_regions(cm_matrix *M) {
  for (int d = 0; d < M->P; ++d)
    if (M->peer_flush_open[d]) {
      cudaIpcCloseMemHandle(M->peer_flush_vals[d]);
      cudaIpcCloseMemHandle(M->peer_flush_meta[d]);
      M->peer_flush_open[d] = false;
    }
  M->ctx->tp->barrier(M->stream);
  M->d_flush_vals.release();
  M->d_flush_meta.release();
  M->flush_region_v = M->flush_region_m = 0;
  std::fill(M->peer_flush_vals.begin(), M->peer_flush_vals.end(), nullptr);
  std::fill(M->peer_flush_meta.begin(), M->peer_flush_meta.end(), nullptr);
}

// the flushed messages of this commit have been applied (or logged) by their owners: drop them
void forget_flushed(cm_matrix *M) {
  for (FlushedMsg &f : M->flushed) {
    f.vals.release();
    f.meta.release();
  }
  M->flushed.clear();
  std::fill(M->flush_used_v.begin(), M->flush_used_v.end(), (size_t)0);
  std::fill(M->flush_used_m.begin(), M->flush_used_m.end(), (size_t)0);
}

// -- log or scatter-add what was delivered, elemental updates, quiescence ---------------------------
int commit_finish(cm_matrix *M, CommitPlan &cp) {
  const int P = cp.P, me = cp.me, K = cp.K;
  if (M->keep_wire) CM_TRY(capture_wire(M, cp));
  if (cp.deferred) {
    // log instead of apply: the messages of this commit stay where they landed
    std::vector<cm_matrix::PendMsg> msgs;
    delivered_messages(M, cp, 0, K, true, msgs);
    M->pend.insert(M->pend.end(), msgs.begin(), msgs.end());
    for (int d = 0; d < P; ++d) {
      M->pend_v[d] = cp.end_v[d];
      M->pend_m[d] = cp.end_m[d];
    }
  } else if (!cp.pipelined) {
    CM_TRY(apply_panels(M, cp, 0, K, true, M->work[0], M->stream, 0));
    trace_mark(M, "finish: scatter-add done");
  }
  for (int src = 0; src < P; ++src) {
    const long n = cp.coo_counts[(size_t)src * P + me];
    if (!n) continue;
    CM_TRY(wait_background(M));  // the tile kernel's write-back is not atomic: no RED into the block beside it
    const char *blob = src == me ? M->d_send_coo.as<char>() + cp.s_coo_base[me] : M->d_recv_coo.as<char>() + cp.r_coo_base[src];
    CM_TRY(apply_coo_blob(M, blob, n));
  }
  CM_CUDA_TRY(cudaGetLastError());
  if (cp.deferred) {
    // delivery is complete (the barrier behind the pack kernels) once this stream has drained; the
    // log is append-only, so no second barrier is needed before the next commit may store
    CM_CUDA_TRY(cudaStreamSynchronize(M->stream));
    CM_CUDA_TRY(cudaGetLastError());
    resolve_timers(M);
    clear_staged(M);
    forget_flushed(M);  // under this policy they all waited at their sources and have been copied into the log
    bool over = false;  // same decision on every rank
    for (int d = 0; d < P; ++d) over = over || (size_t)M->pend_v[d] * M->esz > M->defer_budget;
    if (over) CM_TRY(materialize_all(M));
    else if (M->apply_policy == CM_APPLY_BACKGROUND) CM_TRY(start_background_apply(M));
    trace_dump(M, true);
    return CM_OK;
  }
  // quiescence: everybody has applied everything (reference :737-741)
  CM_TRY(commit_barrier(M, M->stream));
  CM_CUDA_TRY(cudaGetLastError());
  trace_mark(M, "finish: quiescent");
  resolve_timers(M);
  trace_dump(M, true);
  M->mirror_valid = false;
  clear_staged(M);
  forget_flushed(M);  // every owner has applied them: the flush regions are free again
  return ensure_flush_regions(M, cp);
}

// ---- threshold flush with several ranks (reference flush(thresh) inside update(), :330-375, :614) ----
// NOT collective. Everything staged is mapped and packed NOW, one message per owner: straight into the
// owner's flush region of this source over NVLink when there is one with room (the reference's rput,
// :168-171), otherwise into buffers of this rank that the next commit delivers. Either way the staged
// payloads, index sets and descriptors are released on return (borrowed device payloads included), and
// the owner scatter-adds the message inside its next commit, in front of that commit's own messages, so
// the per-(source, owner) stream stays the reference's. Elemental updates stay in their host bins.
int flush_shipped(cm_matrix *M) {
  const int nupd = (int)M->descs.size();
  if (nupd == 0) return CM_OK;
  const int P = M->P, me = M->me, K = kMaxChunks;
  const size_t esz = M->esz;
  if (M->reads_local_block) CM_TRY(wait_background(M));  // fill_lower's payload is the local block itself
  const PanelPlan pp{1, 1};
  long max_rows = 0;
  for (const UpdDesc &d : M->descs) max_rows = std::max<long>(max_rows, d.m);
  trace_mark(M, "flush: begin");
  {
    ScopedTimer t(M, 1, 0);
    MapArenas ma{};
    ScanArenas sa{};
    CM_TRY(map_and_scan(M, pp, nupd, ma, sa));
    // sizes of MY messages only: a local read-back, no rendezvous
    std::vector<MsgCounts> mine((size_t)P * K);
    CM_CUDA_TRY(cudaMemcpyAsync(mine.data(), M->d_totals.p, sizeof(MsgCounts) * mine.size(), cudaMemcpyDeviceToHost, M->stream));
    CM_CUDA_TRY(cudaStreamSynchronize(M->stream));
    M->stats.d2h_bytes += (long)(sizeof(MsgCounts) * mine.size());
    const bool regions = M->p2p && M->apply_policy == 0 && M->flush_region_v > 0;
    std::vector<void *> dst_ptrs(2 * (size_t)K * P, nullptr);
    for (int d = 0; d < P; ++d) {
      MsgCounts c = mine[d];  // panel 0
      if (c.nvals == 0) continue;
      c.ncoo = c.nflushed = 0;
      const size_t vb = (size_t)pad32(c.nvals) * esz, mb = (size_t)meta_bytes(c.nseg, c.nidx);
      FlushedMsg fm;
      fm.d = FlushDesc{me, d, 0, (int)M->flushed.size(), 0, 0, c};
      if (regions && M->flush_used_v[d] + vb <= M->flush_region_v && M->flush_used_m[d] + mb <= M->flush_region_m) {
        fm.d.remote = 1;
        fm.d.off_v = (long)M->flush_used_v[d];
        fm.d.off_m = (long)M->flush_used_m[d];
        M->flush_used_v[d] += vb;
        M->flush_used_m[d] += (mb + 255) & ~(size_t)255;
        dst_ptrs[d] = static_cast<char *>(M->peer_flush_vals[d]) + (size_t)me * M->flush_region_v + fm.d.off_v;
        dst_ptrs[(size_t)K * P + d] = static_cast<char *>(M->peer_flush_meta[d]) + (size_t)me * M->flush_region_m + fm.d.off_m;
        if (d != me) M->stats.bytes_sent += (long)(c.nvals * esz + mb);
        M->stats.flushed_shipped += 1;
      } else {
        M->stats.flushed_held += 1;
        CM_TRY(fm.vals.alloc_exact(vb));
        CM_TRY(fm.meta.alloc_exact(mb));
        dst_ptrs[d] = fm.vals.p;
        dst_ptrs[(size_t)K * P + d] = fm.meta.p;
      }
      M->flushed.push_back(fm);
    }
    CM_CUDA_TRY(cudaMemcpyAsync(M->d_dst_ptrs.p, dst_ptrs.data(), dst_ptrs.size() * sizeof(void *), cudaMemcpyHostToDevice, M->stream));
    launch_pack(M->dtype, M->d_descs.as<UpdDesc>(), nupd, 0, 1, pp, M->g, ma, sa, max_rows, M->stream);
    M->stats.kernel_launches += 1;
    CM_CUDA_TRY(cudaGetLastError());
  }
  // the staged payloads (borrowed ones included) and the host-side arenas are released on return
  CM_CUDA_TRY(cudaStreamSynchronize(M->stream));
  CM_CUDA_TRY(cudaGetLastError());
  trace_mark(M, "flush: packed");
  resolve_timers(M);
  trace_dump(M, false);
  M->stats.flushes += 1;
  clear_staged(M, true);
  return CM_OK;
}

// everything staged on this rank leaves the staging area: scatter-added at once with one rank, packed
// (and, with a flush region, shipped) with several
int flush_any(cm_matrix *M, bool sync) {
  if (M->P == 1 && !M->keep_wire) return flush_direct(M, sync);
  return flush_shipped(M);
}

int commit_exchange(cm_matrix *M) {
  CommitPlan cp;
  cp.P = M->P;
  cp.me = M->me;
  cp.nupd = (int)M->descs.size();
  trace_mark(M, "commit: begin");
  // a staged payload that is the local block itself (fill_lower) is read by the pack kernel: no scatter-add
  // of an earlier commit may still be writing the block
  if (M->reads_local_block) CM_TRY(wait_background(M));
  CM_TRY(commit_counts(M, cp));
  if (cp.empty) {
    // nothing staged on any rank: the counts all-gather already was the rendezvous of this commit
    // (nobody leaves it before everybody has entered it); there is nothing to exchange, to apply or
    // to wait for, so the barriers are skipped (empty-commit latency)
    if (M->profiling) {
      M->event_pool.push_back(cp.ev_pack0);
      M->event_pool.push_back(cp.ev_pack1);
    }
    if (M->keep_wire) {
      M->dbg_msgs.clear();
      M->dbg_coo = M->coo;
      M->dbg_valid = true;
    }
    resolve_timers(M);
    trace_dump(M, true);
    clear_staged(M);
    return CM_OK;
  }
  CM_TRY(commit_layout(M, cp));
  CM_TRY(commit_deliver(M, cp));
  return commit_finish(M, cp);
}

int do_commit(cm_matrix *M) {
  int rc;
  if (M->P == 1 && !M->keep_wire) {
    rc = flush_direct(M);
    if (rc == CM_OK) rc = M->ctx->tp->barrier(M->stream);
  } else {
    rc = commit_exchange(M);
  }
  if (rc == CM_OK) M->stats.commits += 1;
  M->commit_seq += 1;
  return rc;
}

int check_dims(long m_u, long n_u) {
  if (m_u < 0 || n_u < 0) return fail(CM_ERR_INVALID, "negative slice dims %ld x %ld", m_u, n_u);
  if (n_u >= (1L << kItemColBits))
    return fail(CM_ERR_INVALID, "slice has %ld columns; the limit per update is %ld", n_u, (1L << kItemColBits) - 1);
  if (m_u > (long)kRowChunk << kItemChunkBits)
    return fail(CM_ERR_INVALID, "slice has %ld rows; the limit per update is %ld", m_u, (long)kRowChunk << kItemChunkBits);
  return CM_OK;
}

// append one staged update whose payload / indices already have device addresses
int stage_desc(cm_matrix *M, const void *d_data, long m_u, long n_u, long ld, int trans, const long *d_ix,
               const long *d_jx, int mask) {
  if (m_u == 0 || n_u == 0) return CM_OK;
  if ((long)M->descs.size() + 1 > M->max_staged_updates)  // make_room() flushes before this can happen
    return fail(CM_ERR_INVALID, "more than %ld updates staged", M->max_staged_updates);
  UpdDesc d;
  d.data = d_data;
  d.ix = d_ix;
  d.jx = d_jx;
  d.ld = ld;
  d.m = (int)m_u;
  d.n = (int)n_u;
  d.trans = trans ? 1 : 0;
  d.mask = mask;
  d.rbase = M->staged_rows;
  d.cbase = M->staged_cols;
  M->descs.push_back(d);
  M->item_base.push_back(M->staged_items);
  M->staged_rows += m_u;
  M->staged_cols += n_u;
  M->staged_elems += m_u * n_u;
  M->staged_items += n_u * ((m_u + kRowChunk - 1) / kRowChunk);
  M->stats.updates += 1;
  M->stats.elements_staged += m_u * n_u;
  return CM_OK;
}

// reference flush(thresh) trigger, :346 / :614: strict '>' on the number of staged elements
int maybe_flush(cm_matrix *M) {
  if (M->staged_elems + M->coo_total > M->flush_threshold) return flush_any(M, false);
  return CM_OK;
}

// called before an update is staged: one scatter-add addresses at most max_staged_updates segments, so a
// rank that has staged that many flushes first (instead of refusing the update)
int make_room(cm_matrix *M) {
  if ((long)M->descs.size() + 1 > M->max_staged_updates) return flush_any(M, true);
  return CM_OK;
}

int issue_merged_copy(cm_matrix *M) {
  if (M->merge_bytes) {
    CM_CUDA_TRY(cudaMemcpyAsync(M->merge_dst, M->merge_src, M->merge_bytes, cudaMemcpyHostToDevice, M->copy_stream));
    M->merge_bytes = 0;
    M->copies_pending = true;
  }
  return CM_OK;
}

int wait_payload_copies(cm_matrix *M) {
  CM_TRY(issue_merged_copy(M));
  cudaError_t q;
  while ((q = cudaStreamQuery(M->copy_stream)) == cudaErrorNotReady) {
  }
  if (q != cudaSuccess) return fail(CM_ERR_CUDA, "payload copy failed: %s", cudaGetErrorString(q));
  return CM_OK;
}

int stage_host_slice(cm_matrix *M, const void *data, long m_u, long n_u, long ld, int trans, const long *ix,
                     const long *jx, int mask) {
  CM_TRY(check_dims(m_u, n_u));
  if (m_u == 0 || n_u == 0) return CM_OK;
  if (!data || !ix || !jx) return fail(CM_ERR_INVALID, "null payload or index pointer");
  // storage footprint: contiguous dim x other dim with leading dimension ld
  const long rows_s = trans ? n_u : m_u, cols_s = trans ? m_u : n_u;
  if (ld < rows_s) return fail(CM_ERR_INVALID, "ld (%ld) smaller than the stored leading dimension (%ld)", ld, rows_s);
  CM_TRY(make_room(M));
  const size_t esz = M->esz;
  void *d_payload = nullptr;
  CM_TRY(M->payload_arena.alloc((size_t)rows_s * cols_s * esz, &d_payload));
  if (ld == rows_s) {
    const size_t bytes = (size_t)rows_s * cols_s * esz;
    const char *src = static_cast<const char *>(data);
    if (M->defer_copy_wait && M->merge_bytes && src == M->merge_src + M->merge_bytes &&
        static_cast<char *>(d_payload) == M->merge_dst + M->merge_bytes && M->merge_bytes + bytes <= (64u << 20)) {
      M->merge_bytes += bytes;  // extends the pending DMA
    } else {
      CM_TRY(issue_merged_copy(M));
      if (M->defer_copy_wait) {
        M->merge_src = src;
        M->merge_dst = static_cast<char *>(d_payload);
        M->merge_bytes = bytes;
      } else {
        CM_CUDA_TRY(cudaMemcpyAsync(d_payload, data, bytes, cudaMemcpyHostToDevice, M->copy_stream));
      }
    }
  } else {
    CM_TRY(issue_merged_copy(M));
    CM_CUDA_TRY(cudaMemcpy2DAsync(d_payload, (size_t)rows_s * esz, data, (size_t)ld * esz, (size_t)rows_s * esz,
                                  (size_t)cols_s, cudaMemcpyHostToDevice, M->copy_stream));
  }
  M->copies_pending = true;
  M->stats.h2d_bytes += (long)((size_t)rows_s * cols_s * esz);
  // index sets: host memcpy into the pinned arena while the payload DMA is in flight
  void *h_ix = nullptr, *dv_ix = nullptr, *h_jx = nullptr, *dv_jx = nullptr;
  CM_TRY(M->index_arena.alloc((size_t)m_u * 8, &h_ix, &dv_ix));
  memcpy(h_ix, ix, (size_t)m_u * 8);
  if (jx == ix && m_u == n_u) {
    dv_jx = dv_ix;
  } else {
    CM_TRY(M->index_arena.alloc((size_t)n_u * 8, &h_jx, &dv_jx));
    memcpy(h_jx, jx, (size_t)n_u * 8);
  }
  // reference semantics (:611): the caller may reuse `data` on return. Pageable sources are
  // consumed before cudaMemcpyAsync returns; pinned / managed sources are DMA'd asynchronously,
  // so wait for the DMA — polling, because a blocking synchronise costs more than the copy.
  if (!M->defer_copy_wait) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, data) == cudaSuccess) {
      if (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged) CM_TRY(wait_payload_copies(M));
    } else {
      cudaGetLastError();
    }
  }
  M->stats.staging_peak_bytes = std::max(M->stats.staging_peak_bytes, (long)(M->payload_arena.used() + M->index_arena.used()));
  CM_TRY(stage_desc(M, d_payload, m_u, n_u, rows_s, trans, static_cast<const long *>(dv_ix),
                    static_cast<const long *>(dv_jx), mask));
  return maybe_flush(M);
}

}  // namespace

// ====================================================================================
// C ABI
// ====================================================================================
extern "C" {

long cm_numroc(long m, long np, long ip, long mb) {
  const long nblocks = m / mb;
  long loc = (nblocks / np) * mb;
  if (nblocks % np > ip)
    loc += mb;
  else if (nblocks % np == ip)
    loc += m % mb;
  return loc;
}

void cm_owner_of(long i, long j, long nprow, long npcol, long mb, long nb, long lld, int *rank, long *ij) {
  if (rank) *rank = owner_coord(j, nb, npcol) + (int)npcol * owner_coord(i, mb, nprow);
  if (ij) *ij = lld * local_index(j, nb, npcol) + local_index(i, mb, nprow);
}

void cm_descinit_raw(int ictxt, long m, long n, long mb, long nb, long lld, int desc[9]) {
  desc[0] = 1;  // DTYPE_ = dense
  desc[1] = ictxt;
  desc[2] = (int)m;
  desc[3] = (int)n;
  desc[4] = (int)mb;
  desc[5] = (int)nb;
  desc[6] = 0;  // RSRC_
  desc[7] = 0;  // CSRC_
  desc[8] = (int)lld;
}

int cm_create(cm_matrix **out, int dtype, long m, long n, long nprow, long npcol, long mb, long nb, long lld) {
  if (!out) return fail(CM_ERR_INVALID, "out is null");
  *out = nullptr;
  RankCtx *ctx = current_ctx();
  if (!ctx) return fail(CM_ERR_STATE, "runtime not initialised (cm_rt_init) or rank thread not bound");
  if (dtype != CM_F64 && dtype != CM_F32 && dtype != CM_I32) return fail(CM_ERR_INVALID, "bad dtype %d", dtype);
  // reference asserts, :417-428
  if (m <= 0 || n <= 0) return fail(CM_ERR_INVALID, "matrix dims must be positive (%ld x %ld)", m, n);
  if (mb <= 0 || nb <= 0 || nprow <= 0 || npcol <= 0 || lld <= 0) return fail(CM_ERR_INVALID, "bad distribution parameters");
  if (nprow * npcol != ctx->nranks)
    return fail(CM_ERR_INVALID, "NPROW*NPCOL = %ld but rank_n() = %d", nprow * npcol, ctx->nranks);
  if (nprow > kMaxGridDim || npcol > kMaxGridDim || nprow * npcol > 1024)
    return fail(CM_ERR_INVALID, "grid dims above %d (or more than 1024 ranks) unsupported", kMaxGridDim);
  cm_matrix *M = new cm_matrix();
  M->ctx = ctx;
  M->dtype = dtype;
  M->esz = dtype_size(dtype);
  M->P = (int)(nprow * npcol);
  M->me = ctx->rank;
  M->g.m = m;
  M->g.n = n;
  M->g.mb = mb;
  M->g.nb = nb;
  M->g.nprow = nprow;
  M->g.npcol = npcol;
  M->g.lld = lld;
  M->g.myprow = (int)(ctx->rank / npcol);  // row-major grid (see header)
  M->g.mypcol = (int)(ctx->rank % npcol);
  M->m_local = cm_numroc(m, nprow, M->g.myprow, mb);
  M->n_local = cm_numroc(n, npcol, M->g.mypcol, nb);
  if (!roc_ok(M->m_local) || !roc_ok(M->n_local) || M->m_local > lld) {
    const long ml = M->m_local, nl = M->n_local;
    delete M;
    return fail(CM_ERR_INVALID, "local block %ld x %ld incompatible with LLD %ld", ml, nl, lld);
  }
  if (lld >= (1L << 31) || M->n_local >= (1L << 31)) {
    delete M;
    return fail(CM_ERR_INVALID, "local dims must fit in 31 bits");
  }
  M->local_elems = (size_t)lld * M->n_local;
  M->coo.resize(M->P);
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&M->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&M->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) {
    // the scatter-add of panel k must get SM slots while the pack kernel of panel k+1 (launched
    // earlier, thousands of short blocks) is running: give its stream the higher priority, so that
    // freed slots go to its pending blocks first
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    e = cudaStreamCreateWithPriority(&M->apply_stream, cudaStreamNonBlocking, prio_hi);
  }
  for (int k = 0; k < kMaxChunks && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&M->ev_panel[k], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&M->ev_applied, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&M->ev_bg, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&M->copy_done, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&M->d_local, M->local_elems * M->esz);
  if (e == cudaSuccess) e = cudaMemsetAsync(M->d_local, 0, M->local_elems * M->esz, M->stream);  // :431-432
  if (e == cudaSuccess) e = cudaStreamSynchronize(M->stream);
  if (e != cudaSuccess) {
    int rc = fail(e == cudaErrorMemoryAllocation ? CM_ERR_NOMEM : CM_ERR_CUDA, "cm_create: %s", cudaGetErrorString(e));
    cudaGetLastError();
    if (M->d_local) cudaFree(M->d_local);
    if (M->stream) cudaStreamDestroy(M->stream);
    if (M->copy_stream) cudaStreamDestroy(M->copy_stream);
    if (M->copy_done) cudaEventDestroy(M->copy_done);
    delete M;
    return rc;
  }
  // publish local blocks for one-sided reads (dist_object fetch + rget, :589-595)
  M->peer_local.assign(M->P, nullptr);
  M->peer_ipc_open.assign(M->P, false);
  M->peer_local[M->me] = M->d_local;
  if (M->P > 1) {
    if (ctx->tp->same_process()) {
      std::vector<void *> all(M->P);
      int rc = ctx->tp->host_allgather(&M->d_local, all.data(), sizeof(void *), M->stream);
      if (rc != CM_OK) return rc;
      M->peer_local = all;
    } else {
      cudaIpcMemHandle_t mine;
      std::vector<cudaIpcMemHandle_t> all(M->P);
      memset(&mine, 0, sizeof(mine));
      const bool have = cudaIpcGetMemHandle(&mine, M->d_local) == cudaSuccess;
      if (!have) cudaGetLastError();
      int rc = ctx->tp->host_allgather(&mine, all.data(), sizeof(mine), M->stream);
      if (rc != CM_OK) return rc;
      for (int r = 0; r < M->P && have; ++r) {
        if (r == M->me) continue;
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) {
          M->peer_local[r] = p;
          M->peer_ipc_open[r] = true;
        } else {
          cudaGetLastError();
        }
      }
    }
  }
  M->peer_vals.assign(M->P, nullptr);
  M->peer_meta.assign(M->P, nullptr);
  M->cap_vals.assign(M->P, 0);
  M->cap_meta.assign(M->P, 0);
  M->peer_arena_open.assign(M->P, false);
  M->pend_v.assign(M->P, 0);
  M->pend_m.assign(M->P, 0);
  M->peer_flush_vals.assign(M->P, nullptr);
  M->peer_flush_meta.assign(M->P, nullptr);
  M->peer_flush_open.assign(M->P, false);
  M->flush_used_v.assign(M->P, 0);
  M->flush_used_m.assign(M->P, 0);
  {
    // peer-to-peer stores need every rank to be able to map every other rank's memory
    int mine_ok = 1;
    for (int r = 0; r < M->P; ++r) mine_ok &= M->peer_local[r] != nullptr;
    std::vector<int> all_ok(M->P, 0);
    int rc0 = ctx->tp->host_allgather(&mine_ok, all_ok.data(), sizeof(int), M->stream);
    if (rc0 != CM_OK) return rc0;
    M->p2p_possible = true;
    for (int r = 0; r < M->P; ++r) M->p2p_possible &= all_ok[r] != 0;
    if (getenv("CM_B200_PIPELINE")) M->pipeline = atoi(getenv("CM_B200_PIPELINE")) != 0 ? 1 : 0;
    if (getenv("CM_B200_PIPELINE_APPLY_BLOCKS")) M->pipeline_apply_blocks = std::max(0, atoi(getenv("CM_B200_PIPELINE_APPLY_BLOCKS")));
    if (getenv("CM_B200_PANELS")) M->npanel_cfg = std::max(1, atoi(getenv("CM_B200_PANELS")));
    if (getenv("CM_B200_PIPELINE_MIN_ELEMS")) M->pipeline_min_elems = M->pipeline_owner_min_elems = atol(getenv("CM_B200_PIPELINE_MIN_ELEMS"));
    const char *env = getenv("CM_B200_EXCHANGE");  // "nccl" / "p2p" override, must agree on all ranks
    if (env && !strcmp(env, "nccl")) M->exchange_mode = 1;
    if (env && !strcmp(env, "p2p")) M->exchange_mode = 2;
    M->p2p = M->P > 1 && M->p2p_possible && M->exchange_mode != 1;
    if (const char *pol = getenv("CM_B200_APPLY_POLICY")) {  // default policy of new matrices: eager | deferred | background
      if (!strcmp(pol, "deferred")) M->apply_policy = CM_APPLY_DEFERRED;
      if (!strcmp(pol, "background")) M->apply_policy = CM_APPLY_BACKGROUND;
    }
    if (getenv("CM_B200_BG_MIN_ELEMS")) M->bg_min_elems = atol(getenv("CM_B200_BG_MIN_ELEMS"));
    if (getenv("CM_B200_TRACE") && (M->me == 0 || getenv("CM_B200_TRACE_ALL"))) M->trace_left = atoi(getenv("CM_B200_TRACE"));
    if (getenv("CM_B200_TRACE_SKIP")) M->trace_skip = atol(getenv("CM_B200_TRACE_SKIP"));
  }
  int rc = ctx->tp->barrier(M->stream);
  if (rc != CM_OK) return rc;
  *out = M;
  return CM_OK;
}

int cm_destroy(cm_matrix *M) {
  if (!M) return CM_OK;
  int rc = do_commit(M);  // :458
  if (rc == CM_OK) rc = materialize_all(M);
  cudaStreamSynchronize(M->stream);
  for (int r = 0; r < M->P; ++r) {
    if (M->peer_ipc_open[r]) cudaIpcCloseMemHandle(M->peer_local[r]);
    if (M->peer_arena_open[r]) {
      cudaIpcCloseMemHandle(M->peer_vals[r]);
      cudaIpcCloseMemHandle(M->peer_meta[r]);
    }
  }
  forget_flushed(M);
  release_flush_regions(M);        // (contains a barrier)
  M->ctx->tp->barrier(M->stream);  // nobody still reads my block
  cudaFree(M->d_local);
  if (M->mirror_pinned) cudaFreeHost(M->h_mirror); else free(M->h_mirror);
  M->payload_arena.release();
  M->index_arena.release();
  DevBuf *bufs[] = {&M->d_fill_ix, &M->d_fill_jx, &M->d_descs, &M->d_item_base, &M->d_rl, &M->d_rpos, &M->d_roff,
                    &M->d_cl, &M->d_cpos, &M->d_coff, &M->d_uflags, &M->d_radj, &M->d_coo_tmp, &M->d_valoff, &M->d_idxoff, &M->d_itemoff, &M->d_totals,
                    &M->d_send_vals, &M->d_send_meta, &M->d_recv_vals,
                    &M->d_recv_meta, &M->d_send_coo, &M->d_recv_coo, &M->d_counts_all, &M->d_dst_ptrs, &M->d_coo_counts};
  for (DevBuf *b : bufs) b->release();
  for (auto &w : M->work) w.release();
  cudaStreamDestroy(M->apply_stream);
  for (auto e : M->ev_panel) cudaEventDestroy(e);
  cudaEventDestroy(M->ev_applied);
  cudaEventDestroy(M->ev_bg);
  M->h_counts.release();
  for (auto e : M->event_pool) cudaEventDestroy(e);
  cudaStreamDestroy(M->stream);
  cudaStreamDestroy(M->copy_stream);
  cudaEventDestroy(M->copy_done);
  delete M;
  return rc;
}

int cm_update_slice(cm_matrix *M, const void *data, long m_u, long n_u, long ld, int trans, const long *ix,
                    const long *jx) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  return stage_host_slice(M, data, m_u, n_u, ld, trans, ix, jx, CM_MASK_NONE);
}

int cm_update_sym(cm_matrix *M, const void *data, long n_u, long ld, int trans, const long *ix) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  return stage_host_slice(M, data, n_u, n_u, ld, trans, ix, ix, CM_MASK_UPPER);
}

int cm_update_elems(cm_matrix *M, const void *vals, const long *is, const long *js, long n) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  if (n < 0 || (n > 0 && (!vals || !is || !js))) return fail(CM_ERR_INVALID, "bad elemental update arguments");
  const unsigned char *v = static_cast<const unsigned char *>(vals);
  const Geom &g = M->g;
  for (long k = 0; k < n; ++k) {
    // :624-627
    const int rank = owner_coord(js[k], g.nb, g.npcol) + (int)g.npcol * owner_coord(is[k], g.mb, g.nprow);
    const long ij = g.lld * local_index(js[k], g.nb, g.npcol) + local_index(is[k], g.mb, g.nprow);
    CooBin &b = M->coo[rank];
    b.inds.push_back(ij);
    b.vals.insert(b.vals.end(), v + (size_t)k * M->esz, v + (size_t)(k + 1) * M->esz);
  }
  M->coo_total += n;
  M->stats.elements_staged += n;
  return maybe_flush(M);
}

int cm_update_elem(cm_matrix *M, const void *val, long i, long j) { return cm_update_elems(M, val, &i, &j, 1); }

int cm_update_slice_device(cm_matrix *M, const void *d_data, long m_u, long n_u, long ld, int trans, const long *d_ix,
                           const long *d_jx, int mask) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  CM_TRY(check_dims(m_u, n_u));
  if (m_u == 0 || n_u == 0) return CM_OK;
  if (!d_data || !d_ix || !d_jx) return fail(CM_ERR_INVALID, "null device pointer");
  if (mask < 0 || mask > 2) return fail(CM_ERR_INVALID, "bad mask %d", mask);
  if (ld < (trans ? n_u : m_u)) return fail(CM_ERR_INVALID, "ld too small");
  CM_TRY(make_room(M));
  CM_TRY(stage_desc(M, d_data, m_u, n_u, ld, trans, d_ix, d_jx, mask));
  M->borrowed_pending = true;
  return maybe_flush(M);
}

// batch forms: identical to `count` consecutive calls; they exist so that harnesses with a
// slow call path (ctypes) do not measure their own per-call overhead
int cm_update_slices_device(cm_matrix *M, long count, const void *const *d_data, const long *m_u, const long *n_u,
                            const long *ld, const int *trans, const long *const *d_ix, const long *const *d_jx,
                            int mask) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  if (count < 0 || (count > 0 && (!d_data || !m_u || !n_u || !ld || !d_ix || !d_jx)))
    return fail(CM_ERR_INVALID, "bad batch arguments");
  for (long k = 0; k < count; ++k)
    CM_TRY(cm_update_slice_device(M, d_data[k], m_u[k], n_u[k], ld[k], trans ? trans[k] : 0, d_ix[k], d_jx[k], mask));
  return CM_OK;
}

int cm_update_slices(cm_matrix *M, long count, const void *const *data, const long *m_u, const long *n_u,
                     const long *ld, const int *trans, const long *const *ix, const long *const *jx) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  if (count < 0 || (count > 0 && (!data || !m_u || !n_u || !ld || !ix || !jx)))
    return fail(CM_ERR_INVALID, "bad batch arguments");
  // the caller gets its buffers back when THIS call returns, so the payload DMAs of the whole
  // batch may overlap each other; one wait at the end (a single update() pays a full DMA round
  // trip: ~26 GB/s for 512 KiB payloads against ~55 GB/s for a streamed pinned copy, B200 measured)
  M->defer_copy_wait = true;
  int rc = CM_OK;
  for (long k = 0; k < count && rc == CM_OK; ++k)
    rc = cm_update_slice(M, data[k], m_u[k], n_u[k], ld[k], trans ? trans[k] : 0, ix[k], jx[k]);
  M->defer_copy_wait = false;
  if (rc == CM_OK) rc = wait_payload_copies(M);
  return rc;
}

int cm_fill_lower(cm_matrix *M) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  // One masked, transposed slice update whose payload is the local block itself: view element
  // (k,l) = local[l + k*LLD] goes to global (row = gcol(k), col = grow(l)) iff row > col, which
  // is the loop at :696-709 (ix = global column of local column j, jx = global row of local row i).
  const Geom &g = M->g;
  CM_TRY(materialize_local(M));  // the payload is the local block itself: it must be current
  CM_TRY(M->d_fill_ix.ensure(8 * (size_t)M->n_local));
  CM_TRY(M->d_fill_jx.ensure(8 * (size_t)M->m_local));
  launch_iota_global(M->d_fill_ix.as<long>(), M->n_local, g.nb, g.npcol, g.mypcol, M->stream);
  launch_iota_global(M->d_fill_jx.as<long>(), M->m_local, g.mb, g.nprow, g.myprow, M->stream);
  M->stats.kernel_launches += 2;
  CM_TRY(check_dims(M->n_local, M->m_local));
  CM_TRY(make_room(M));
  CM_TRY(stage_desc(M, M->d_local, M->n_local, M->m_local, g.lld, 1, M->d_fill_ix.as<long>(), M->d_fill_jx.as<long>(),
                    CM_MASK_STRICT_LOWER));
  M->reads_local_block = true;  // the commit that ships this packs it before any scatter-add touches the block
  // single rank: apply now, so later updates cannot change what fill_lower reads (the
  // reference reads the block at call time)
  if (M->P == 1 && !M->keep_wire) return flush_direct(M);
  return CM_OK;
}

int cm_flush(cm_matrix *M) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  return flush_any(M, true);
}

int cm_commit(cm_matrix *M) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  return do_commit(M);
}

int cm_reset(cm_matrix *M) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  CM_TRY(do_commit(M));  // :493
  CM_TRY(materialize_all(M));
  CM_CUDA_TRY(cudaMemsetAsync(M->d_local, 0, M->local_elems * M->esz, M->stream));  // :496-497
  M->mirror_valid = false;
  return M->ctx->tp->barrier(M->stream);  // :500
}

int cm_local_data_device(cm_matrix *M, const void **out) {
  if (!M || !out) return fail(CM_ERR_INVALID, "null argument");
  CM_TRY(materialize_local(M));
  *out = M->d_local;
  return CM_OK;
}

int cm_local_data_copy(cm_matrix *M, void *dst_host) {
  if (!M || !dst_host) return fail(CM_ERR_INVALID, "null argument");
  CM_TRY(materialize_local(M));
  CM_CUDA_TRY(cudaMemcpyAsync(dst_host, M->d_local, M->local_elems * M->esz, cudaMemcpyDeviceToHost, M->stream));
  CM_CUDA_TRY(cudaStreamSynchronize(M->stream));
  M->stats.d2h_bytes += (long)(M->local_elems * M->esz);
  return CM_OK;
}

int cm_local_data_host(cm_matrix *M, const void **out) {
  if (!M || !out) return fail(CM_ERR_INVALID, "null argument");
  if (!M->h_mirror) {
    // page-locked: the copy of a 512 MiB block runs at the PCIe rate instead of through the driver's staging buffer
    if (cudaMallocHost(&M->h_mirror, M->local_elems * M->esz) == cudaSuccess) {
      M->mirror_pinned = true;
    } else {
      cudaGetLastError();
      M->h_mirror = malloc(M->local_elems * M->esz);
      M->mirror_pinned = false;
    }
    if (!M->h_mirror) return fail(CM_ERR_NOMEM, "host mirror allocation failed");
    M->mirror_valid = false;
  }
  CM_TRY(materialize_local(M));  // may invalidate the mirror
  if (!M->mirror_valid) {
    CM_TRY(cm_local_data_copy(M, M->h_mirror));
    M->mirror_valid = true;
  }
  *out = M->h_mirror;
  return CM_OK;
}

// ---- save() / load(), reference :788-873 / :888-967 ------------------------------------------------
// Same file as the reference's MPI-IO darray view produces: the dense GLOBAL matrix, column-major
// (element (i,j) at byte (i + j*M) * sizeof(T)), native endianness, no header. No MPI: every rank
// writes / reads the runs of MB consecutive global rows it owns with pwrite / pread. As in the
// reference there is no implicit commit (:778-779) and both calls are collective.
static int for_each_owned_run(cm_matrix *M, int fd, unsigned char *host, bool write) {
  const Geom &g = M->g;
  const size_t esz = M->esz;
  for (long lj = 0; lj < M->n_local; ++lj) {
    const long gj = global_index(lj, g.nb, g.npcol, g.mypcol);
    if (gj >= g.n) continue;
    for (long li = 0; li < M->m_local; li += g.mb) {
      const long gi = global_index(li, g.mb, g.nprow, g.myprow);
      if (gi >= g.m) continue;
      const long len = std::min(std::min(g.mb, M->m_local - li), g.m - gi);
      unsigned char *p = host + ((size_t)lj * g.lld + li) * esz;
      const off_t off = (off_t)(((size_t)gj * g.m + gi) * esz);
      size_t done = 0;
      while (done < (size_t)len * esz) {
        const ssize_t k = write ? pwrite(fd, p + done, (size_t)len * esz - done, off + (off_t)done)
                                : pread(fd, p + done, (size_t)len * esz - done, off + (off_t)done);
        if (k <= 0) return fail(CM_ERR_STATE, "%s failed at offset %lld: %s", write ? "pwrite" : "pread",
                                (long long)off, k == 0 ? "short file" : strerror(errno));
        done += (size_t)k;
      }
    }
  }
  return CM_OK;
}

int cm_save(cm_matrix *M, const char *fname) {
  if (!M || !fname) return fail(CM_ERR_INVALID, "null argument");
  Transport *tp = M->ctx->tp;
  CM_TRY(tp->barrier(M->stream));  // :798
  std::vector<unsigned char> host(M->local_elems * M->esz);
  CM_TRY(cm_local_data_copy(M, host.data()));
  if (M->me == 0) {  // one rank creates the file at its full size
    const int fd0 = open(fname, O_CREAT | O_WRONLY | O_TRUNC, 0644);
    if (fd0 < 0 || ftruncate(fd0, (off_t)((size_t)M->g.m * M->g.n * M->esz)) != 0) {
      const int e = errno;
      if (fd0 >= 0) close(fd0);
      tp->barrier(M->stream);
      tp->barrier(M->stream);
      return fail(CM_ERR_STATE, "cannot create %s: %s", fname, strerror(e));
    }
    close(fd0);
  }
  CM_TRY(tp->barrier(M->stream));
  int rc = CM_OK;
  const int fd = open(fname, O_WRONLY);
  if (fd < 0)
    rc = fail(CM_ERR_STATE, "cannot open %s: %s", fname, strerror(errno));
  else {
    rc = for_each_owned_run(M, fd, host.data(), true);
    if (fsync(fd) != 0 && rc == CM_OK) rc = fail(CM_ERR_STATE, "fsync(%s): %s", fname, strerror(errno));
    close(fd);
  }
  const int rcb = tp->barrier(M->stream);  // :869
  return rc != CM_OK ? rc : rcb;
}

int cm_load(cm_matrix *M, const char *fname) {
  if (!M || !fname) return fail(CM_ERR_INVALID, "null argument");
  Transport *tp = M->ctx->tp;
  CM_TRY(materialize_all(M));      // what was committed before the load is applied before it
  CM_TRY(tp->barrier(M->stream));  // :898
  std::vector<unsigned char> host(M->local_elems * M->esz, 0);  // padding rows stay zero
  int rc = CM_OK;
  const int fd = open(fname, O_RDONLY);
  if (fd < 0)
    rc = fail(CM_ERR_STATE, "cannot open %s: %s", fname, strerror(errno));
  else {
    rc = for_each_owned_run(M, fd, host.data(), false);
    close(fd);
  }
  if (rc == CM_OK) {
    // replaces the current contents of the local block (:879)
    if (cudaMemcpyAsync(M->d_local, host.data(), host.size(), cudaMemcpyHostToDevice, M->stream) != cudaSuccess ||
        cudaStreamSynchronize(M->stream) != cudaSuccess)
      rc = fail(CM_ERR_CUDA, "cm_load: %s", cudaGetErrorString(cudaGetLastError()));
    M->stats.h2d_bytes += (long)host.size();
    M->mirror_valid = false;
  }
  const int rcb = tp->barrier(M->stream);  // :963
  return rc != CM_OK ? rc : rcb;
}

int cm_get(cm_matrix *M, long i, long j, void *out) {
  if (!M || !out) return fail(CM_ERR_INVALID, "null argument");
  int rank;
  long ij;
  cm_owner_of(i, j, M->g.nprow, M->g.npcol, M->g.mb, M->g.nb, M->g.lld, &rank, &ij);
  if (rank < 0 || rank >= M->P || !M->peer_local[rank])
    return fail(CM_ERR_STATE, "local block of rank %d is not mapped on rank %d (CUDA IPC unavailable)", rank, M->me);
  if (rank == M->me) CM_TRY(materialize_local(M));
  else if (!M->pend_v.empty() && (M->pend_v[rank] || M->pend_m[rank]))
    return fail(CM_ERR_STATE, "rank %d holds committed updates it has not applied yet (deferred scatter-add): "
                "call cm_materialize (collective) before one-sided reads", rank);
  CM_CUDA_TRY(cudaStreamSynchronize(M->stream));
  CM_CUDA_TRY(cudaMemcpy(out, static_cast<const char *>(M->peer_local[rank]) + (size_t)ij * M->esz, M->esz,
                         cudaMemcpyDefault));
  M->stats.d2h_bytes += (long)M->esz;
  return CM_OK;
}

long cm_m(const cm_matrix *M) { return M ? M->g.m : -1; }
long cm_n(const cm_matrix *M) { return M ? M->g.n : -1; }
long cm_m_local(const cm_matrix *M) { return M ? M->m_local : -1; }
long cm_n_local(const cm_matrix *M) { return M ? M->n_local : -1; }
long cm_lld(const cm_matrix *M) { return M ? M->g.lld : -1; }
long cm_pgrid_row(const cm_matrix *M) { return M ? M->g.myprow : -1; }
long cm_pgrid_col(const cm_matrix *M) { return M ? M->g.mypcol : -1; }
int cm_dtype(const cm_matrix *M) { return M ? M->dtype : -1; }

long cm_get_flush_threshold(const cm_matrix *M) { return M ? M->flush_threshold : -1; }
int cm_set_flush_threshold(cm_matrix *M, long elements) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  M->flush_threshold = elements;
  return CM_OK;
}
int cm_get_progress_interval(const cm_matrix *M) { return M ? M->progress_interval : -1; }
int cm_set_progress_interval(cm_matrix *M, int interval) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  M->progress_interval = interval;
  return CM_OK;
}
int cm_set_deterministic(cm_matrix *M, int on) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  M->deterministic = on ? 1 : 0;
  return CM_OK;
}
int cm_set_sorted_sweep(cm_matrix *M, int on) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  M->sorted_sweep = on ? 1 : 0;
  return CM_OK;
}
int cm_set_apply_kernel(cm_matrix *M, int kernel, double tile_density) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  if (kernel < 0 || kernel > 2) return fail(CM_ERR_INVALID, "apply kernel must be 0 (auto), 1 (RED) or 2 (tile)");
  M->apply_kernel = kernel;
  if (tile_density > 0) M->tile_density = tile_density;
  return CM_OK;
}
int cm_set_apply_policy(cm_matrix *M, int policy, long budget_bytes) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  if (policy != CM_APPLY_EAGER && policy != CM_APPLY_DEFERRED && policy != CM_APPLY_BACKGROUND)
    return fail(CM_ERR_INVALID, "apply policy must be CM_APPLY_EAGER (0), CM_APPLY_DEFERRED (1) or CM_APPLY_BACKGROUND (2)");
  if (!M->descs.empty() || M->coo_total || !M->flushed.empty())
    return fail(CM_ERR_STATE, "change the apply policy only with nothing staged (right after a commit)");
  CM_TRY(materialize_all(M));
  M->apply_policy = policy;
  if (budget_bytes > 0) M->defer_budget = (size_t)budget_bytes;
  return M->ctx->tp->barrier(M->stream);
}
int cm_get_apply_policy(const cm_matrix *M) { return M ? M->apply_policy : -1; }
int cm_materialize(cm_matrix *M) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  return materialize_all(M);
}

int cm_set_exchange_mode(cm_matrix *M, int mode) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  if (mode < 0 || mode > 2)
    return fail(CM_ERR_INVALID, "exchange mode must be 0 (auto), 1 (all-to-all-v) or 2 (peer-to-peer stores)");
  if (!M->descs.empty() || M->coo_total || !M->flushed.empty())
    return fail(CM_ERR_STATE, "change the exchange mode only with nothing staged (right after a commit)");
  if (mode >= 2 && M->P > 1 && !M->p2p_possible)
    return fail(CM_ERR_STATE, "peer-to-peer exchange needs every rank's memory mapped (CUDA IPC unavailable)");
  CM_TRY(materialize_all(M));  // the arenas are about to be released
  M->exchange_mode = mode;
  M->p2p = M->P > 1 && M->p2p_possible && mode != 1;
  // arenas allocated under the other mode are not mapped by the peers: start over
  for (int r = 0; r < M->P; ++r) {
    if (M->peer_arena_open[r]) {
      cudaIpcCloseMemHandle(M->peer_vals[r]);
      cudaIpcCloseMemHandle(M->peer_meta[r]);
      M->peer_arena_open[r] = false;
    }
    M->peer_vals[r] = M->peer_meta[r] = nullptr;
    M->cap_vals[r] = M->cap_meta[r] = 0;
  }
  CM_TRY(M->ctx->tp->barrier(M->stream));
  M->d_recv_vals.release();
  M->d_recv_meta.release();
  release_flush_regions(M);  // mapped for peer-to-peer stores only
  return CM_OK;
}
int cm_get_exchange_mode(const cm_matrix *M) { return M ? (M->p2p ? 2 : 1) : -1; }
int cm_set_profiling(cm_matrix *M, int on) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  M->profiling = on ? 1 : 0;
  return CM_OK;
}
int cm_stream(cm_matrix *M, void **out) {
  if (!M || !out) return fail(CM_ERR_INVALID, "null argument");
  *out = M->stream;
  return CM_OK;
}

int cm_descinit(const cm_matrix *M, int ictxt, int desc[9]) {
  if (!M || !desc) return fail(CM_ERR_INVALID, "null argument");
  cm_descinit_raw(ictxt, M->g.m, M->g.n, M->g.mb, M->g.nb, M->g.lld, desc);
  return CM_OK;
}

int cm_get_stats(cm_matrix *M, cm_stats *out) {
  if (!M || !out) return fail(CM_ERR_INVALID, "null argument");
  *out = M->stats;
  return CM_OK;
}
int cm_reset_stats(cm_matrix *M) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  memset(&M->stats, 0, sizeof(M->stats));
  return CM_OK;
}

// ---- parity hooks -------------------------------------------------------------------
int cm_debug_keep_wire(cm_matrix *M, int on) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  if (!M->descs.empty() || M->coo_total || !M->flushed.empty())
    return fail(CM_ERR_STATE, "toggle wire capture only with nothing staged (right after a commit)");
  M->keep_wire = on ? 1 : 0;
  M->dbg_valid = false;
  return CM_OK;
}

// expands the compact messages to owner `dst` into the reference's (inds, data) stream; pass null
// outputs to count only. The reference appends update after update to the owner's bin (:611), so the
// panels of a commit are walked update-major: all panels of update 0, then of update 1, ... — the
// reference's order, because the grouping kernel only cuts an update into panels when its column
// indices are non-decreasing in panel number.
static long expand_wire(cm_matrix *M, int dst, long *inds_out, unsigned char *data_out) {
  const Geom &g = M->g;
  const size_t esz = M->esz;
  const int dp = dst / (int)g.npcol, dq = dst % (int)g.npcol;
  long n = 0;
  int ngroups = 0;
  for (const auto &dm : M->dbg_msgs) ngroups = std::max(ngroups, dm.group + 1);
  for (int grp = 0; grp < ngroups; ++grp) {  // threshold flushes in order, then the commit's own messages
    std::vector<const cm_matrix::DbgMsg *> msgs;
    long nseg_max = 0;
    for (const auto &dm : M->dbg_msgs) {
      if (dm.dst != dst || dm.group != grp || dm.meta.size() < sizeof(WireHdr)) continue;
      msgs.push_back(&dm);
      nseg_max = std::max(nseg_max, reinterpret_cast<const WireHdr *>(dm.meta.data())->nseg);
    }
    for (long u = 0; u < nseg_max; ++u)
      for (const cm_matrix::DbgMsg *dm : msgs) {
        const char *meta = dm->meta.data();
        const WireHdr *h = reinterpret_cast<const WireHdr *>(meta);
        if (u >= h->nseg) continue;
        const WireSeg *ws = reinterpret_cast<const WireSeg *>(meta + sizeof(WireHdr));
        const int *idx = reinterpret_cast<const int *>(meta + sizeof(WireHdr) + h->nseg * sizeof(WireSeg));
        const unsigned char *vals = reinterpret_cast<const unsigned char *>(dm->vals.data());
        const WireSeg &s = ws[u];
        const int *lrow = idx + s.idx_off, *lcol = lrow + s.nr;
        for (int b = 0; b < s.nc; ++b)
          for (int a = 0; a < s.nr; ++a) {
            if (s.mask) {
              const long gr = global_index(lrow[a], g.mb, g.nprow, dp), gc = global_index(lcol[b], g.nb, g.npcol, dq);
              if (s.mask == CM_MASK_UPPER ? !(gr <= gc) : !(gr > gc)) continue;
            }
            if (inds_out) {
              inds_out[n] = g.lld * (long)lcol[b] + lrow[a];
              memcpy(data_out + (size_t)n * esz,
                     vals + ((size_t)s.val_off + (size_t)a + (size_t)b * seg_col_stride(s.nr)) * esz, esz);
            }
            ++n;
          }
      }
  }
  const CooBin &cb = M->dbg_coo[dst];
  for (size_t k = 0; k < cb.inds.size(); ++k) {
    if (inds_out) {
      inds_out[n] = cb.inds[k];
      memcpy(data_out + (size_t)n * esz, cb.vals.data() + k * esz, esz);
    }
    ++n;
  }
  return n;
}

int cm_debug_set_staging_limit(cm_matrix *M, long max_updates) {
  if (!M) return fail(CM_ERR_INVALID, "null matrix");
  if (max_updates < 1 || max_updates > (1L << kItemSegBits) - 1) return fail(CM_ERR_INVALID, "limit out of range");
  M->max_staged_updates = max_updates;
  return CM_OK;
}

int cm_debug_wire_count(cm_matrix *M, int dst, long *n_out) {
  if (!M || !n_out) return fail(CM_ERR_INVALID, "null argument");
  if (!M->dbg_valid) return fail(CM_ERR_STATE, "no captured wire (cm_debug_keep_wire + cm_commit first)");
  if (dst < 0 || dst >= M->P) return fail(CM_ERR_INVALID, "bad owner %d", dst);
  *n_out = expand_wire(M, dst, nullptr, nullptr);
  return CM_OK;
}

int cm_debug_wire_expand(cm_matrix *M, int dst, long *inds_out, void *data_out) {
  if (!M || !inds_out || !data_out) return fail(CM_ERR_INVALID, "null argument");
  if (!M->dbg_valid) return fail(CM_ERR_STATE, "no captured wire (cm_debug_keep_wire + cm_commit first)");
  if (dst < 0 || dst >= M->P) return fail(CM_ERR_INVALID, "bad owner %d", dst);
  expand_wire(M, dst, inds_out, static_cast<unsigned char *>(data_out));
  return CM_OK;
}

}  // extern "C"
