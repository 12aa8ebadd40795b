// cm_kernels.cu — sm_100a kernels of the ConvergentMatrix assembly path.
//
// Reference (swfrench/convergent-matrix, include/convergent_matrix.hpp) -> kernel:
//   owner map + local index (:606-610)            k_map_group
//   Bin::append per owner, column-major (:605-613) k_scan_dest + k_pack (per-owner segments)
//   XBuff wire format (:128-133)                   WireHdr / WireSeg written by k_pack
//   apply<T>: elts[inds[i]] += data[i] (:178-187)  k_apply_tile (dense flushes: column tiles in
//                                                  shared memory) / k_apply_red (sparse: RED.ADD)
//
// All of this is HBM-bound integer / gather-scatter work: no tensor cores. The design
// levers are (a) O(m+n) index work per update instead of O(m*n) (owner and local index are
// separable in i and j), (b) values-only wire format, (c) a column-sorted sweep of the local
// block so that repeated hits on a sector meet in L2 instead of DRAM.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>

#include "cm_kernels.h"

// Kernel launches and dynamic shared memory are spelled through two macros so that the very same
// kernel sources can also be compiled by the host compiler against tests/emu/ (a test double of the
// CUDA runtime that runs a block's threads as fibers; it backs the CPU-side kernel-logic tests and
// is never part of libcm_b200.so). Under nvcc they are the plain CUDA spellings.
#ifndef CM_EMU
#define CM_LAUNCH(grid, block, smem, stream, ...) __VA_ARGS__<<<(grid), (block), (smem), (stream)>>>
#define CM_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

namespace cmb {

size_t dtype_size(int dtype) { return dtype == 0 ? 8 : 4; }

int kernel_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

__device__ __forceinline__ uint64_t encode_item(long seg, int col, int chunk) {
  return ((uint64_t)seg << (kItemColBits + kItemChunkBits)) | ((uint64_t)(uint32_t)col << kItemChunkBits) |
         (uint64_t)(uint32_t)chunk;
}

// ------------------------------------------------------------------------------------
// (1) owner map + stable grouping.  grid (nupd, 2): y = 0 rows (ix), y = 1 columns (jx).
// For every index: owner coordinate p (:606/:609) and local index l (:607/:610). Output is
// grouped by p, stable within a group (== the order the reference appends to bin (p,*)).
// Columns of a commit cut into panels are grouped by gq = panel(l) * NPCOL + p instead (PanelPlan).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_map_group(const UpdDesc *__restrict__ descs, Geom g, PanelPlan pp, MapArenas ma) {
  const int u = blockIdx.x;
  const int axis = blockIdx.y;
  const UpdDesc d = descs[u];
  const long *__restrict__ idx = axis == 0 ? d.ix : d.jx;
  const int len = axis == 0 ? d.m : d.n;
  const long blk = axis == 0 ? g.mb : g.nb;
  const int npo = (int)(axis == 0 ? g.nprow : g.npcol);      // owners along this axis
  const int np = axis == 0 ? npo : npo * pp.npanel;           // groups along this axis
  const long base = axis == 0 ? d.rbase : d.cbase;
  int *__restrict__ lout = (axis == 0 ? ma.rl : ma.cl) + base;
  int *__restrict__ pout = (axis == 0 ? ma.rpos : ma.cpos) + base;
  int *__restrict__ off_out = (axis == 0 ? ma.roff : ma.coff) + (long)u * (np + 1);

  const int tid = threadIdx.x;
  // radj (rows only): per owner row, how many of its rows directly follow the previous index of the update (same
  // MB block, so the same owner and the next local row). The owner's choice between the two scatter-adds needs it:
  // RED pays per 32-byte sector, and a run of consecutive rows shares its sectors.
  __shared__ int adjc[kMaxGridDim];
  if (np == 1) {  // single group along this axis: identity grouping
    bool rep = false;
    int adj = 0;
    if (tid == 0) adjc[0] = 0;
    __syncthreads();
    for (int k = tid; k < len; k += blockDim.x) {
      lout[k] = (int)local_index(idx[k], blk, 1);
      pout[k] = k;
      if (k > 0) rep |= idx[k] <= idx[k - 1];
      if (k > 0) adj += idx[k] == idx[k - 1] + 1;
    }
    if (axis == 0 && rep) atomicOr(&ma.uflags[u], kSegRowsMayRepeat);
    if (axis == 0 && adj) atomicAdd(&adjc[0], adj);
    __syncthreads();
    if (tid == 0) {
      off_out[0] = 0;
      off_out[1] = len;
      if (axis == 0) ma.radj[(long)u * 2] = adjc[0];
    }
    return;
  }
  // Panels keep the reference's per-owner order (update after update, column after column, :605-613)
  // only if this update's columns are non-decreasing in panel number — true for sorted jx, because
  // panel_cols is a multiple of NB. An update whose columns are not goes into panel 0 as a whole.
  bool panels = axis == 1 && pp.npanel > 1;
  if (panels) {
    int bad = 0;
    for (int k = tid + 1; k < len; k += blockDim.x)
      bad |= panel_of(pp, local_index(idx[k], blk, npo)) < panel_of(pp, local_index(idx[k - 1], blk, npo));
    panels = __syncthreads_or(bad) == 0;
  }
  auto group_of = [&](long gi) {
    int p = owner_coord(gi, blk, npo);
    if (panels) p += npo * panel_of(pp, local_index(gi, blk, npo));
    return p;
  };

  __shared__ int cnt[kMaxGridDim];
  __shared__ int gbase[kMaxGridDim + 1];
  __shared__ int run[kMaxGridDim];
  __shared__ int wcnt[8][kMaxGridDim];
  const int lane = tid & 31, warp = tid >> 5;

  if (tid < np) {
    cnt[tid] = 0;
    run[tid] = 0;
    adjc[tid] = 0;
  }
  for (int k = tid; k < 8 * kMaxGridDim; k += blockDim.x) (&wcnt[0][0])[k] = 0;
  __syncthreads();
  // pass 1: group sizes
  for (int k = tid; k < len; k += blockDim.x) atomicAdd(&cnt[group_of(idx[k])], 1);
  if (axis == 0) {
    bool rep = false;
    for (int k = tid + 1; k < len; k += blockDim.x) {
      const long a = idx[k - 1], b = idx[k];
      rep |= b <= a;
      if (b == a + 1 && b % blk != 0) atomicAdd(&adjc[owner_coord(b, blk, npo)], 1);
    }
    if (rep) atomicOr(&ma.uflags[u], kSegRowsMayRepeat);
  }
  __syncthreads();
  if (axis == 0 && tid < np) ma.radj[(long)u * (np + 1) + tid] = adjc[tid];
  if (tid == 0) {
    int acc = 0;
    for (int p = 0; p < np; ++p) {
      gbase[p] = acc;
      off_out[p] = acc;
      acc += cnt[p];
    }
    gbase[np] = acc;
    off_out[np] = acc;
  }
  __syncthreads();
  // pass 2: stable placement, 256 entries at a time
  for (int c0 = 0; c0 < len; c0 += blockDim.x) {
    const int k = c0 + tid;
    const bool valid = k < len;
    long gi = 0;
    int p = np;  // sentinel group for lanes past the end
    if (valid) {
      gi = idx[k];
      p = group_of(gi);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, p);
    const int r = __popc(peers & ((1u << lane) - 1u));
    if (valid && r == 0) wcnt[warp][p] = __popc(peers);
    __syncthreads();
    if (valid) {
      int off = gbase[p] + run[p] + r;
      for (int w = 0; w < warp; ++w) off += wcnt[w][p];
      lout[off] = (int)local_index(gi, blk, npo);
      pout[off] = k;
    }
    __syncthreads();
    if (tid < np) {
      int s = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        s += wcnt[w][tid];
        wcnt[w][tid] = 0;
      }
      run[tid] += s;
    }
    __syncthreads();
  }
}

void launch_map_group(const UpdDesc *d_descs, int nupd, const Geom &g, const PanelPlan &pp, const MapArenas &ma,
                      cudaStream_t s) {
  if (nupd <= 0) return;
  CM_LAUNCH(dim3(nupd, 2), 256, 0, s, k_map_group)(d_descs, g, pp, ma);
}

// ------------------------------------------------------------------------------------
// (2) per-owner exclusive scans over the staged updates. grid = P blocks (one per owner).
// ------------------------------------------------------------------------------------
struct Long3 {
  long a, b, c, e, f;
  __host__ __device__ Long3 operator+(const Long3 &o) const { return Long3{a + o.a, b + o.b, c + o.c, e + o.e, f + o.f}; }
};

__global__ void __launch_bounds__(256) k_scan_dest(int nupd, PanelPlan pp, Geom g, MapArenas ma, ScanArenas sa,
                                                   const long *__restrict__ coo_counts) {  // [2P]: ncoo, nflushed per owner
  using BlockScan = cub::BlockScan<Long3, 256>;
  __shared__ typename BlockScan::TempStorage tmp;
  const int dst = blockIdx.x;
  const int panel = blockIdx.y;
  const int P = gridDim.x;
  if (panel >= pp.npanel) {  // unused panel slot: an empty message
    if (threadIdx.x == 0) sa.totals[panel * P + dst] = MsgCounts{0, 0, 0, 0, 0, 0, 0, 0};
    return;
  }
  const int p = dst / (int)g.npcol, q = dst % (int)g.npcol;
  const int gq = panel * (int)g.npcol + q;  // column group of (panel, owner column)
  const int npr = (int)g.nprow + 1, npc = (int)g.npcol * pp.npanel + 1;
  const long obase = ((long)panel * P + dst) * nupd;
  Long3 running{0, 0, 0, 0, 0};
  for (int c0 = 0; c0 < nupd; c0 += 256) {
    const int u = c0 + threadIdx.x;
    Long3 v{0, 0, 0, 0, 0};
    if (u < nupd) {
      const int nr = ma.roff[(long)u * npr + p + 1] - ma.roff[(long)u * npr + p];
      const int nc = ma.coff[(long)u * npc + gq + 1] - ma.coff[(long)u * npc + gq];
      const long nv = (long)nr * nc;
      if (nv > 0) {
        v.a = seg_slots(nr, nc);
        v.b = nr + nc;
        v.c = (long)nc * ((nr + kRowChunk - 1) / kRowChunk);
        v.e = nv;
        v.f = (long)ma.radj[(long)u * npr + p] * nc;
      }
    }
    Long3 excl, total;
    BlockScan(tmp).ExclusiveScan(v, excl, Long3{0, 0, 0, 0, 0}, cub::Sum(), total);
    if (u < nupd) {
      sa.valoff[obase + u] = running.a + excl.a;
      sa.idxoff[obase + u] = running.b + excl.b;
      sa.itemoff[obase + u] = running.c + excl.c;
    }
    running = running + total;
    __syncthreads();
  }
  // a message without values ships nothing at all (no segment table either): nseg = 0
  if (threadIdx.x == 0)
    sa.totals[panel * P + dst] = MsgCounts{running.a, running.b, running.a ? (long)nupd : 0, running.c, running.e,
                                           panel == 0 ? coo_counts[dst] : 0, panel == 0 ? coo_counts[P + dst] : 0, running.f};
}

void launch_scan(int nupd, const PanelPlan &pp, const Geom &g, const MapArenas &ma, const ScanArenas &sa,
                 const long *d_coo_counts, cudaStream_t s) {
  const int P = (int)(g.nprow * g.npcol);
  CM_LAUNCH(dim3(P, kMaxChunks), 256, 0, s, k_scan_dest)(nupd, pp, g, ma, sa, d_coo_counts);
}

// ------------------------------------------------------------------------------------
// (3) pack. grid (nupd, ytiles): block (u, y) handles the grouped columns y, y+ytiles, ... of the
// panels [k0, k1) of update u: for every owner row p it gathers the rows R_p of that column into
// the (panel, p, q) segment — contiguous, coalesced stores; gathers hit every sector of the source
// column exactly once per block. Block y == 0 also writes the update's wire metadata.
// ------------------------------------------------------------------------------------
constexpr int kPackThreads = 384;
constexpr int kPackMaxDst = 1024;  // (k1 - k0) * P destinations per launch

// wire metadata of update u for the panels [k0, k1): header (first update only), WireSeg, index lists.
// Called by the y == 0 block of an update with the block's shared roff / coff tables.
__device__ __forceinline__ void pack_write_meta(const UpdDesc &d, int u, int nupd, int k0, int k1, int nprow, int npcol,
                                                const int *roff, const int *coff, const MapArenas &ma, const ScanArenas &sa) {
  const int tid = threadIdx.x, P = nprow * npcol;
  const int *__restrict__ rl = ma.rl + d.rbase;
  const int *__restrict__ cl = ma.cl + d.cbase;
  for (int k = k0; k < k1; ++k)
    for (int dst = 0; dst < P; ++dst) {
      const MsgCounts c = sa.totals[k * P + dst];
      if (c.nvals == 0) continue;  // empty message: nothing was reserved for it
      const int p = dst / npcol, q = dst % npcol, gq = k * npcol + q;
      const int nr = roff[p + 1] - roff[p], nc = coff[gq + 1] - coff[gq];
      char *mb = sa.dst_meta[k * P + dst];
      if (u == 0 && tid == 0) {  // message header
        WireHdr *h = reinterpret_cast<WireHdr *>(mb);
        h->nseg = c.nseg;
        h->nvals = c.nvals;
        h->nidx = c.nidx;
        h->nitems = c.nitems;
      }
      WireSeg *ws = reinterpret_cast<WireSeg *>(mb + sizeof(WireHdr)) + u;
      const bool empty = (long)nr * nc == 0;
      const long obase = ((long)k * P + dst) * nupd + u;
      const long ioff = sa.idxoff[obase];
      if (tid == 0) {
        ws->nr = empty ? 0 : nr;
        ws->nc = empty ? 0 : nc;
        ws->mask = d.mask;
        ws->flags = ma.uflags[u];
        ws->val_off = sa.valoff[obase];
        ws->idx_off = ioff;
        ws->item_off = sa.itemoff[obase];
      }
      if (!empty) {
        int *iw = reinterpret_cast<int *>(mb + sizeof(WireHdr) + (long)nupd * sizeof(WireSeg)) + ioff;
        for (int a = tid; a < nr; a += blockDim.x) iw[a] = rl[roff[p] + a];
        for (int cc = tid; cc < nc; cc += blockDim.x) iw[nr + cc] = cl[coff[gq] + cc];
      }
    }
}

constexpr int kPackCols = 4;  // columns a thread keeps in flight (8 measured no faster: NVLink-bound)
template <typename T>
__global__ void __launch_bounds__(kPackThreads) k_pack(const UpdDesc *__restrict__ descs, int nupd, int k0, int k1, PanelPlan pp,
                                                       Geom g, MapArenas ma, ScanArenas sa) {
  constexpr int NC = kPackCols;
  const int u = blockIdx.x;
  const UpdDesc d = descs[u];
  const int nprow = (int)g.nprow, npcol = (int)g.npcol, P = nprow * npcol;
  const int ncg = npcol * pp.npanel;  // column groups of an update
  __shared__ int roff[kMaxGridDim + 1];
  __shared__ int coff[kMaxGridDim + 1];
  __shared__ int tstart[kMaxGridDim + 1];  // first thread slot of each owner row's rows (16-aligned)
  __shared__ int gdst[kMaxGridDim];        // column group -> (panel - k0) * P + owner column
  __shared__ T *base[kPackMaxDst];         // per (panel, owner): where this update's segment starts
  const int tid = threadIdx.x;
  if (tid <= nprow) roff[tid] = ma.roff[(long)u * (nprow + 1) + tid];
  if (tid <= ncg) coff[tid] = ma.coff[(long)u * (ncg + 1) + tid];
  if (tid < ncg) gdst[tid] = (tid / npcol - k0) * P + tid % npcol;
  for (int i = tid; i < (k1 - k0) * P; i += blockDim.x) {
    const int k = k0 + i / P, dst = i % P;
    base[i] = static_cast<T *>(sa.dst_vals[k * P + dst]) + sa.valoff[((long)k * P + dst) * nupd + u];
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int p = 0; p < nprow; ++p) {
      tstart[p] = acc;
      acc += (roff[p + 1] - roff[p] + kSegAlign - 1) & ~(kSegAlign - 1);
    }
    tstart[nprow] = acc;
  }
  __syncthreads();
  const T *__restrict__ data = static_cast<const T *>(d.data);
  const int *__restrict__ rpos = ma.rpos + d.rbase;
  const int *__restrict__ cpos = ma.cpos + d.cbase;
  const long src_rs = d.trans ? d.ld : 1, src_cs = d.trans ? 1 : d.ld;
  const int cbeg = coff[k0 * npcol], cend = coff[k1 * npcol];  // grouped columns of the panels [k0, k1)
  const int gy = (int)gridDim.y;

  // Thread slot t -> (owner row p, slot a of that owner's segment rows). Slots of one owner row
  // start on a multiple of 16, and a tall segment's columns are 16-slot aligned in the message, so
  // every half-warp stores one whole 128-byte line (fp64) — locally or across NVLink.
  for (int t = tid; t < tstart[nprow]; t += blockDim.x) {
    int p = 0;
    while (t >= tstart[p + 1]) ++p;
    const int a = t - tstart[p];
    const int nr = roff[p + 1] - roff[p];
    const bool live = a < nr;
    const int stride = seg_col_stride(nr);
    const long srow = live ? (long)rpos[roff[p] + a] * src_rs : 0;
    T *const *brow = base + p * npcol;
    int kc = cbeg + (int)blockIdx.y;
    int gq = k0 * npcol;  // column group of the current column: only moves forward
    // NC columns in flight per thread
    for (; kc + (NC - 1) * gy < cend; kc += NC * gy) {
      T x[NC];
      T *out[NC];
      int cp[NC];
      // the source-column positions first (independent loads), then the gathers that depend on
      // them; the destination lookups (shared memory) fill the wait
#pragma unroll
      for (int i = 0; i < NC; ++i) cp[i] = cpos[kc + i * gy];
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const int c = kc + i * gy;
        while (c >= coff[gq + 1]) ++gq;
        out[i] = brow[gdst[gq]] + (long)(c - coff[gq]) * stride + a;
      }
#pragma unroll
      for (int i = 0; i < NC; ++i) x[i] = live ? data[srow + (long)cp[i] * src_cs] : T(0);
      if (live) {
#pragma unroll
        for (int i = 0; i < NC; ++i) *out[i] = x[i];
      }
    }
    for (; kc < cend; kc += gy) {
      while (kc >= coff[gq + 1]) ++gq;
      if (live) brow[gdst[gq]][(long)(kc - coff[gq]) * stride + a] = data[srow + (long)cpos[kc] * src_cs];
    }
  }

  if (blockIdx.y == 0) pack_write_meta(d, u, nupd, k0, k1, nprow, npcol, roff, coff, ma, sa);
}

// Blocks per update. The grid should hold several resident waves of blocks, so that the last, partly
// filled wave costs little (B200, K2 at 8 GPUs: 256 updates per commit were cut into 1280 blocks for 1036
// resident ones, and the pack kernel ran two block times for 1.24 waves of work).
static int pick_ytiles(int nupd, long max_cols, int waves = 4) {
  const int target = kernel_sm_count() * 8 * waves;
  long y = (target + nupd - 1) / nupd;
  if (y < 1) y = 1;
  if (y > max_cols) y = max_cols;
  if (y > 1024) y = 1024;
  if (y < 1) y = 1;
  return (int)y;
}

// Block size of the pack kernel. Thread slot t <-> one row of the update (slots of one owner row padded
// to 16), so an update of m_u rows has about m_u + 16 * NPROW slots: the block is sized to a whole number
// of passes over the slots with few idle threads (B200, 8 GPUs: K4's 32-row updates 36.0 -> 31.5 ms per
// step against a fixed 384-thread block; neutral on 256- and 512-row updates, which are NVLink-bound).
static int pack_threads(long max_rows, long nprow) {
  const long slots = std::max(32L, max_rows + 16 * nprow);
  const long passes = (slots + kPackThreads - 1) / kPackThreads;
  const long per = (slots + passes - 1) / passes;
  return (int)std::min<long>(std::max<long>((per + 31) & ~31L, 64), kPackThreads);
}

void launch_pack(int dtype, const UpdDesc *d_descs, int nupd, int k0, int k1, const PanelPlan &pp, const Geom &g,
                 const MapArenas &ma, const ScanArenas &sa, long max_rows, cudaStream_t s) {
  if (nupd <= 0 || k1 <= k0) return;
  const dim3 grid(nupd, pick_ytiles(nupd, 128));
  const int threads = pack_threads(max_rows, g.nprow);
  switch (dtype) {
    case 0: CM_LAUNCH(grid, threads, 0, s, k_pack<double>)(d_descs, nupd, k0, k1, pp, g, ma, sa); break;
    case 1: CM_LAUNCH(grid, threads, 0, s, k_pack<float>)(d_descs, nupd, k0, k1, pp, g, ma, sa); break;
    default: CM_LAUNCH(grid, threads, 0, s, k_pack<int>)(d_descs, nupd, k0, k1, pp, g, ma, sa); break;
  }
}

// ------------------------------------------------------------------------------------
// (4a) single-rank path: one Seg per staged update, pointing at the payload itself.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_build_direct(const UpdDesc *__restrict__ descs, const long *__restrict__ item_base,
                                                      int esz, MapArenas ma, Seg *__restrict__ segs,
                                                      uint32_t *__restrict__ keys, uint64_t *__restrict__ items) {
  const int u = blockIdx.x;
  const UpdDesc d = descs[u];
  const int *cl = ma.cl + d.cbase;
  if (blockIdx.y == 0 && threadIdx.x == 0) {
    Seg s;
    s.vals = d.data;
    s.lrow = ma.rl + d.rbase;
    s.lcol = cl;
    s.row_stride = d.trans ? d.ld : 1;
    s.col_stride = d.trans ? 1 : d.ld;
    s.nr = d.m;
    s.nc = d.n;
    s.mask = d.mask;
    s.flags = ma.uflags[u];
    segs[u] = s;
  }
  if (keys == nullptr) return;  // Seg table only
  const int nch = (d.m + kRowChunk - 1) / kRowChunk;
  const long total = (long)d.n * nch;
  const long ib = item_base[u];
  for (long t = (long)blockIdx.y * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.y * blockDim.x) {
    const int b = (int)(t / nch), ch = (int)(t % nch);
    keys[ib + t] = (uint32_t)cl[b];
    items[ib + t] = encode_item(u, b, ch);
  }
  (void)esz;
}

void launch_build_direct(int dtype, const UpdDesc *d_descs, const long *d_item_base, int nupd, const Geom &,
                         const MapArenas &ma, Seg *d_segs, uint32_t *d_keys, uint64_t *d_items, cudaStream_t s) {
  if (nupd <= 0) return;
  const dim3 grid(nupd, pick_ytiles(nupd, 64, 1));
  CM_LAUNCH(grid, 256, 0, s, k_build_direct)(d_descs, d_item_base, (int)dtype_size(dtype), ma, d_segs, d_keys, d_items);
}

// ------------------------------------------------------------------------------------
// (4b) owner side: received messages -> Seg table + (key, item) pairs. One warp per segment.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_build_recv(const MsgRef *__restrict__ msgs, int esz, Seg *__restrict__ segs,
                                                    uint32_t *__restrict__ keys, uint64_t *__restrict__ items) {
  const MsgRef m = msgs[blockIdx.y];
  const WireHdr *h = reinterpret_cast<const WireHdr *>(m.meta);
  const long nseg = h->nseg;
  const WireSeg *wsegs = reinterpret_cast<const WireSeg *>(m.meta + sizeof(WireHdr));
  const int *idx = reinterpret_cast<const int *>(m.meta + sizeof(WireHdr) + nseg * sizeof(WireSeg));
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long u = warp; u < nseg; u += nwarps) {
    const WireSeg ws = wsegs[u];
    const long sid = m.seg_base + u;
    const int *lrow = idx + ws.idx_off;
    const int *lcol = lrow + ws.nr;
    if (lane == 0) {
      Seg s;
      s.vals = static_cast<const char *>(m.vals) + ws.val_off * esz;
      s.lrow = lrow;
      s.lcol = lcol;
      s.row_stride = 1;
      s.col_stride = seg_col_stride(ws.nr);
      s.nr = ws.nr;
      s.nc = ws.nc;
      s.mask = ws.mask;
      s.flags = ws.flags;
      segs[sid] = s;
    }
    if (keys == nullptr) continue;  // Seg table only
    const int nch = (ws.nr + kRowChunk - 1) / kRowChunk;
    const long total = (long)ws.nc * nch;
    const long ib = m.item_base + ws.item_off;
    for (long t = lane; t < total; t += 32) {
      const int b = (int)(t / nch), ch = (int)(t % nch);
      keys[ib + t] = (uint32_t)lcol[b];
      items[ib + t] = encode_item(sid, b, ch);
    }
  }
}

void launch_build_recv(int dtype, const MsgRef *d_msgs, int nmsgs, long max_seg_per_msg, Seg *d_segs, uint32_t *d_keys,
                       uint64_t *d_items, cudaStream_t s) {
  if (nmsgs <= 0 || max_seg_per_msg <= 0) return;
  long blocks = (max_seg_per_msg + 7) / 8;  // 8 warps per block, one warp per segment
  const long cap = (long)kernel_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  CM_LAUNCH(dim3((unsigned)blocks, nmsgs), 256, 0, s, k_build_recv)(d_msgs, (int)dtype_size(dtype), d_segs, d_keys, d_items);
}

// ------------------------------------------------------------------------------------
// (5) sort items by destination column (stable LSD radix sort, CUB).
// ------------------------------------------------------------------------------------
size_t sort_temp_bytes(long nitems, int end_bit) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                  (const uint64_t *)nullptr, (uint64_t *)nullptr, nitems, 0, end_bit);
  return bytes;
}

void launch_sort(void *d_temp, size_t temp_bytes, const uint32_t *keys_in, uint32_t *keys_out, const uint64_t *items_in,
                 uint64_t *items_out, long nitems, int end_bit, cudaStream_t s) {
  if (nitems <= 0) return;
  cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, keys_in, keys_out, items_in, items_out, nitems, 0, end_bit, s);
}

// ------------------------------------------------------------------------------------
// (5b) after the sort: one ItemRec per item (absolute pointers, so the scatter-add has a
// single dependent load instead of item -> Seg -> lists), the start of every destination
// column's run in the sorted order.
// ------------------------------------------------------------------------------------
// col_start (optional, sorted keys only): col_start[c] = first sorted position whose key >= c, for
// c in [0, ncols]; the thread at the head of a run writes the entries of its column and of the
// empty columns before it, the last thread the tail.
// first position in the strictly increasing rows[0, n) whose value is >= key
__device__ __forceinline__ int first_row_ge(const int *__restrict__ rows, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(rows + mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// cuts (optional; local blocks of 2 .. kMaxCutTiles row bands): cuts[t * kMaxCutTiles + k - 1] = first element of
// item t whose row lies in band k or above, k = 1 .. ntiles - 1 — where the item's (strictly increasing) rows
// cross the band boundaries. The tile kernel trims a staged item to its band with two loads instead of two
// binary searches: 16 dependent loads per staged item were 40 % of a band's time at K2 / NS shapes.
constexpr int kMaxCutTiles = 8;
__global__ void __launch_bounds__(256) k_build_recs(const Seg *__restrict__ segs, const uint32_t *__restrict__ keys,
                                                    const uint64_t *__restrict__ items, long nitems, int esz,
                                                    ItemRec *__restrict__ recs, long ncols, long *__restrict__ col_start,
                                                    unsigned short *__restrict__ cuts, int ntiles, int tile_rows) {
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < nitems; t += (long)gridDim.x * blockDim.x) {
    if (col_start) {
      const long key = keys[t];
      const long prev = t == 0 ? -1 : (long)keys[t - 1];
      for (long c = prev + 1; c <= key && c <= ncols; ++c) col_start[c] = t;
      if (t == nitems - 1)
        for (long c = key + 1; c <= ncols; ++c) col_start[c] = nitems;
    }
    const uint64_t it = items[t];
    const long sid = (long)(it >> (kItemColBits + kItemChunkBits));
    const int b = (int)((it >> kItemChunkBits) & ((1u << kItemColBits) - 1u));
    const int ch = (int)(it & ((1u << kItemChunkBits) - 1u));
    const Seg s = segs[sid];
    const int a0 = ch * kRowChunk;
    ItemRec r;
    r.v = static_cast<const char *>(s.vals) + ((long)b * s.col_stride + (long)a0 * s.row_stride) * esz;
    r.lrow = s.lrow + a0;
    r.n = min(s.nr - a0, kRowChunk);
    r.row_stride = (int)s.row_stride;
    r.mask_flags = s.mask | (s.flags << 8);
    r.col = (int)keys[t];
    recs[t] = r;
    if (cuts && s.flags == 0 && s.row_stride == 1)
      for (int k = 1; k < ntiles; ++k) cuts[t * kMaxCutTiles + k - 1] = (unsigned short)first_row_ge(r.lrow, r.n, k * tile_rows);
  }
}

int tile_count(int dtype, long m_local) {
  const long tile_rows = 65536 / (long)dtype_size(dtype);
  return (int)((m_local + tile_rows - 1) / tile_rows);
}

size_t cuts_bytes(int dtype, long m_local, long nitems) {
  const int ntiles = tile_count(dtype, m_local);
  return ntiles > 1 && ntiles <= kMaxCutTiles ? (size_t)nitems * kMaxCutTiles * sizeof(unsigned short) : 0;
}

void launch_build_recs(int dtype, const Seg *d_segs, const uint32_t *d_keys, const uint64_t *d_items, long nitems,
                       ItemRec *d_recs, long n_local, long *d_col_start, unsigned short *d_cuts, long m_local, cudaStream_t s) {
  if (nitems <= 0) return;
  const long cap = (long)kernel_sm_count() * 8;
  long blocks = std::min((nitems + 255) / 256, cap);
  const int ntiles = tile_count(dtype, m_local);
  CM_LAUNCH((unsigned)blocks, 256, 0, s, k_build_recs)(d_segs, d_keys, d_items, nitems, (int)dtype_size(dtype), d_recs, n_local,
                                                d_col_start, d_cuts, ntiles, (int)(65536 / dtype_size(dtype)));
}

// ------------------------------------------------------------------------------------
// (6a) scatter-add with RED: apply<T> of the reference (:178-187). One warp per item (one
// column chunk of one segment, <= 256 rows). Values stream in coalesced; every destination
// element is a RED.ADD at local[col*LLD + lrow[a]]. Items arrive sorted by destination
// column, so the active window of the local block is a few columns wide and stays in L2.
// Best when the flush touches the local block sparsely (few hits per column).
// ------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void red_add(T *p, T v) {
  atomicAdd(p, v);  // result unused -> RED.E.ADD.{F64,F32,S32}.STRONG.GPU
}

__device__ __forceinline__ bool mask_keep(int mask, long grow, long gcol) {
  return mask == 1 ? (grow <= gcol) : (grow > gcol);
}

// G = lanes that share one item: 32 (a warp per item) for ordinary items; 16 or 8 when the items of a
// flush are tiny on average (a 32 x 32 update on a 4 x 2 grid leaves 8-row items), so that a warp
// works on 2 or 4 items at a time instead of idling three quarters of its lanes.
template <typename T, int G>
__global__ void __launch_bounds__(256) k_apply_red(const ItemRec *__restrict__ recs, long nitems, T *__restrict__ local,
                                                   Geom g) {
  constexpr int kPerWarp = 32 / G;
  const int lane = threadIdx.x & 31;
  const int sub = lane % G, grp = lane / G;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long t = warp * kPerWarp + grp; t < nitems; t += nwarps * kPerWarp) {
    const ItemRec r = recs[t];
    T *__restrict__ dcol = local + (long)r.col * g.lld;
    const T *__restrict__ v = static_cast<const T *>(r.v);
    const int *__restrict__ lrow = r.lrow;
    const int mask = r.mask_flags & 0xff;
    if (mask == 0 && r.row_stride == 1) {
      int a = sub;
      // 4 independent value/index loads in flight per lane before the REDs
      for (; a + 3 * G < r.n; a += 4 * G) {
        const T x0 = v[a], x1 = v[a + G], x2 = v[a + 2 * G], x3 = v[a + 3 * G];
        const int r0 = lrow[a], r1 = lrow[a + G], r2 = lrow[a + 2 * G], r3 = lrow[a + 3 * G];
        red_add(dcol + r0, x0);
        red_add(dcol + r1, x1);
        red_add(dcol + r2, x2);
        red_add(dcol + r3, x3);
      }
      for (; a < r.n; a += G) red_add(dcol + lrow[a], v[a]);
    } else {
      const long gcol = global_index(r.col, g.nb, g.npcol, g.mypcol);
      for (int a = sub; a < r.n; a += G) {
        const int row = lrow[a];
        if (mask != 0 && !mask_keep(mask, global_index(row, g.mb, g.nprow, g.myprow), gcol)) continue;
        red_add(dcol + row, v[(long)a * r.row_stride]);
      }
    }
  }
}

template <typename K>
static int resident_blocks(K kernel, int threads, size_t smem) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  return per_sm * kernel_sm_count();
}

template <typename T, int G>
static void launch_apply_red_g(const ItemRec *d_recs, long nitems, T *d_local, const Geom &g, cudaStream_t s) {
  // exactly one resident wave: the strided item order is also the time order of the sweep
  static const int resident = resident_blocks(k_apply_red<T, G>, 256, 0);
  constexpr int kPerBlock = 8 * (32 / G);
  const long blocks = std::min((nitems + kPerBlock - 1) / kPerBlock, (long)resident);
  CM_LAUNCH((unsigned)blocks, 256, 0, s, k_apply_red<T, G>)(d_recs, nitems, d_local, g);
}

template <typename T>
static void launch_apply_red_t(const ItemRec *d_recs, long nitems, long elements, T *d_local, const Geom &g, cudaStream_t s) {
  const long avg = elements / std::max(nitems, 1L);  // rows per item
  if (avg <= 8)
    launch_apply_red_g<T, 8>(d_recs, nitems, d_local, g, s);
  else if (avg <= 16)
    launch_apply_red_g<T, 16>(d_recs, nitems, d_local, g, s);
  else
    launch_apply_red_g<T, 32>(d_recs, nitems, d_local, g, s);
}

void launch_apply_red(int dtype, const ItemRec *d_recs, long nitems, long elements, void *d_local, const Geom &g,
                      cudaStream_t s) {
  if (nitems <= 0) return;
  switch (dtype) {
    case 0: launch_apply_red_t(d_recs, nitems, elements, static_cast<double *>(d_local), g, s); break;
    case 1: launch_apply_red_t(d_recs, nitems, elements, static_cast<float *>(d_local), g, s); break;
    default: launch_apply_red_t(d_recs, nitems, elements, static_cast<int *>(d_local), g, s); break;
  }
}

// ------------------------------------------------------------------------------------
// (6b) column-tile scatter-add. One block owns one tile of the local block — one column x up to
// kTileRows rows (64 KiB) — in shared memory. It walks the column's run of items in sorted order;
// thread t owns element t of every item (an item is at most kRowChunk = 256 rows of one segment
// column) and adds it into the tile with a plain shared-memory read-modify-write. An update's row
// indices are strictly increasing, so the elements of ONE item never collide; one __syncthreads per
// item keeps different items — which may hit the same rows — in order. The values and row indices
// of BLK items are in flight in registers while the previous BLK are being added (two register
// sets), with an L2 prefetch of the value lines a few blocks further ahead. Items whose rows may
// repeat, masked items (symmetric update / fill_lower) and strided payloads take a simple in-order
// loop. Local blocks taller than a tile are cut into row bands (kMulti): every band's block walks the
// same items and takes the rows that fall into its band (a contiguous sub-range of the item, found
// by binary search when the item is staged). Finally the tile is added to its column in HBM by ONE
// bulk reduce-add of the TMA unit (cp.reduce.async.bulk .add, SASS UBLKRED.G.S.ADD): no thread waits
// on a dependent global load for the write-back (B200, K1: 2.07 ms against 2.76 ms per step with a
// per-thread load / add / store write-back). Columns that are not 16-byte aligned fall back to
// per-thread 16-byte loads / stores. A tile has one owner and its items are applied in a fixed
// order, so the result is bit-exact run to run (this is also the deterministic-order mode). DRAM
// traffic is the compulsory s*E + 2*s*(touched columns) instead of one sector round trip per
// element. Best when the flush hits the local block densely.
// ------------------------------------------------------------------------------------
constexpr int kRecChunk = 144;  // items of a run staged in shared memory at a time (multiple of 2*BLK)

// The staged items of a chunk, one array per field: a thread fetches the fields of BLK consecutive
// items with a few 16-byte broadcast loads instead of one 32-byte record per item (the shared-memory
// pipe is the busiest unit of this kernel: every wavefront saved on bookkeeping goes to the tile).
template <typename T>
struct TileSmem {
  static constexpr int kTileRows = 65536 / (int)sizeof(T);
  alignas(128) T acc[kTileRows];       // 64 KiB: the tile (one column x kTileRows rows)
  alignas(16) const T *v[kRecChunk];   // first value of each staged item
  alignas(16) const int *rows[kRecChunk];  // its row indices
  alignas(16) int n[kRecChunk];        // rows (0 for padding)
  int stride[kRecChunk];               // elements between rows in v   (in-order path only)
  int mf[kRecChunk];                   // mask | flags << 8            (in-order path only)
};

// largest local row whose global row index is <= gcol (-1 if none): with it the triangular
// masks of the symmetric update / fill_lower (:654, :701) become one integer compare per row
__device__ __forceinline__ int mask_threshold(long gcol, long m_local, const Geom &g) {
  long lo = 0, hi = m_local;  // first local row with global index > gcol
  while (lo < hi) {
    const long mid = (lo + hi) >> 1;
    if (global_index(mid, g.mb, g.nprow, g.myprow) <= gcol) lo = mid + 1; else hi = mid;
  }
  return (int)lo - 1;
}

template <typename P>
struct alignas(16) Pair16 {  // two pointers, one 16-byte shared-memory load
  P x, y;
};
struct alignas(8) Pair8 {
  int x, y;
};

// fetch element `tid` of the BLK staged items starting at index `j0` (plain items: stride 1; j0 is a
// multiple of BLK, BLK is even, so the field arrays are read 16 / 8 bytes at a time). The loaded
// registers are NOT touched here — not even to subtract the tile's first row — so the warp leaves
// this phase with its loads still in flight and meets them again only in tile_add_block, after the
// previous block's adds; `okmask` (bit i: thread has an element in item i) depends on shared memory only.
template <typename T, int BLK>
__device__ __forceinline__ void tile_load_block(const TileSmem<T> &sm, int j0, int tid, int (&row)[BLK], T (&val)[BLK],
                                                unsigned &okmask) {
  static_assert(BLK % 2 == 0, "items are fetched two at a time");
  const T *vp[BLK];
  const int *rp[BLK];
  int nn[BLK];
#pragma unroll
  for (int i = 0; i < BLK; i += 2) {
    const Pair16<const T *> a = *reinterpret_cast<const Pair16<const T *> *>(&sm.v[j0 + i]);
    const Pair16<const int *> b = *reinterpret_cast<const Pair16<const int *> *>(&sm.rows[j0 + i]);
    const Pair8 c = *reinterpret_cast<const Pair8 *>(&sm.n[j0 + i]);
    vp[i] = a.x;
    vp[i + 1] = a.y;
    rp[i] = b.x;
    rp[i + 1] = b.y;
    nn[i] = c.x;
    nn[i + 1] = c.y;
  }
  okmask = 0;
#pragma unroll
  for (int i = 0; i < BLK; ++i) {
    const bool ok = tid < nn[i];
    const int aa = ok ? tid : 0;  // idle threads re-read element 0 (always valid memory)
    okmask |= (ok ? 1u : 0u) << i;
    row[i] = __ldg(rp[i] + aa);
    val[i] = __ldg(vp[i] + aa);
  }
}

// L2 prefetch of the value lines of the BLK items starting at `j0` (item i: threads 16*i .. 16*i+15,
// one 128-byte line each): issued kTilePrefetchBlocks blocks ahead, so that the register loads that
// follow later find their data in L2 instead of waiting a full DRAM round trip behind a barrier
constexpr int kTilePrefetchBlocks = 2;  // blocks ahead (B200, K1: 0 -> 2.30, 1 -> 2.16, 2 -> 2.10, 3 -> 2.16 ms per step)
template <typename T, int BLK>
__device__ __forceinline__ void tile_prefetch_block(const TileSmem<T> &sm, int j0, int tid) {
  const int i = tid >> 4, line = tid & 15;
  if (i < BLK) {
#ifndef CM_EMU
    if (line * 128 < sm.n[j0 + i] * (int)sizeof(T))
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(sm.v[j0 + i]) + line * 128));
#endif
  }
}

template <typename T, int BLK>
__device__ __forceinline__ void tile_add_block(T *acc, int r0, int h, const int (&row)[BLK], const T (&val)[BLK],
                                               unsigned okmask) {
#pragma unroll
  for (int i = 0; i < BLK; ++i) {
    const int rr = row[i] - r0;
    if (((okmask >> i) & 1u) && (unsigned)rr < (unsigned)h) acc[rr] += val[i];
    __syncthreads();  // items strictly in order
  }
}

// Item-parallel walk of a staged chunk (the default, order-free mode): warp w takes the staged items w, w + 8, ...
// whole — lane l their elements l, l + 32, ... — and adds them into the tile with shared-memory atomics, so no
// barrier separates the items and every lane has work whatever the height of the band (a 256-row item leaves 64
// rows in one of four row bands: thread-per-element keeps a quarter of the block busy between two barriers). G items
// x 2 passes of values and row indices are in flight per lane. The order in which items meet on an element is
// not fixed: results agree with the ordered walk within rounding, not bit for bit (the deterministic mode keeps
// the ordered walk). kGeneral: masks, strided payloads (repeated rows need nothing special here).
template <typename T, bool kGeneral>
__device__ __forceinline__ void tile_atomic_chunk(TileSmem<T> &sm, int cnt, int r0, int h, int thr, int tid) {
  constexpr int G = 4, PASS = 2, kWarps = 8;
  const int warp = tid >> 5, lane = tid & 31;
  T *acc = sm.acc;
  for (int j0 = warp; j0 < cnt; j0 += kWarps * G) {
    const T *vp[G];
    const int *rp[G];
    int nn[G], st[G], mk[G];
    int nmax = 0;
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int j = j0 + kWarps * i;
      const bool ok = j < cnt;
      vp[i] = sm.v[ok ? j : 0];
      rp[i] = sm.rows[ok ? j : 0];
      nn[i] = ok ? sm.n[j] : 0;
      st[i] = kGeneral ? sm.stride[ok ? j : 0] : 1;
      mk[i] = kGeneral ? (sm.mf[ok ? j : 0] & 0xff) : 0;
      nmax = max(nmax, nn[i]);
    }
    for (int a0 = 0; a0 < nmax; a0 += 32 * PASS) {
      int row[G][PASS];
      T val[G][PASS];
#pragma unroll
      for (int i = 0; i < G; ++i)
#pragma unroll
        for (int p = 0; p < PASS; ++p) {
          const int a = a0 + 32 * p + lane;
          const int aa = a < nn[i] ? a : 0;  // idle lanes re-read element 0 (always valid memory)
          row[i][p] = __ldg(rp[i] + aa);
          val[i][p] = __ldg(vp[i] + (long)aa * st[i]);
        }
#pragma unroll
      for (int i = 0; i < G; ++i)
#pragma unroll
        for (int p = 0; p < PASS; ++p) {
          const int a = a0 + 32 * p + lane;
          const int rr = row[i][p] - r0;
          bool ok = a < nn[i] && (unsigned)rr < (unsigned)h;
          if (kGeneral) ok = ok && (mk[i] == 0 || (mk[i] == 1) == (rr + r0 <= thr));
          if (ok) atomicAdd(acc + rr, val[i][p]);
        }
    }
  }
}

// Write-back of a tile by the TMA unit: one bulk reduce-add (cp.reduce.async.bulk, SASS UBLKRED) adds the
// whole shared-memory tile to its column in global memory. Needs 16-byte aligned ends; adds zeros where
// nothing was hit.
#ifndef CM_EMU
__device__ __forceinline__ unsigned tile_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_reduce_add(double *gdst, const double *ssrc, unsigned bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst),
               "r"(tile_smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_reduce_add(float *gdst, const float *ssrc, unsigned bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst),
               "r"(tile_smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_reduce_add(int *gdst, const int *ssrc, unsigned bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.s32 [%0], [%1], %2;" ::"l"(gdst),
               "r"(tile_smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
#endif

template <typename T>
struct alignas(16) Vec16 {
  T x[16 / sizeof(T)];
};

// kMulti: the local block is taller than one tile, so a block only owns part of each item's rows
// (staged items are trimmed to that part).
// BLK: items a thread keeps in flight per register set (two sets); CTAS: resident blocks per SM
// bulk: write the tile back with one bulk reduce-add (0: per-thread loads / stores; the kernel-logic
// build and misaligned columns always take those)
// kOrdered: items of a column are added strictly in sorted order, one barrier per item (bit-exact run to run: the
// deterministic mode); otherwise the item-parallel walk with shared-memory atomics (tile_atomic_chunk)
template <typename T, bool kMulti, int BLK, int CTAS, bool kOrdered>
__global__ void __launch_bounds__(256, CTAS) k_apply_tile(const ItemRec *__restrict__ recs, const long *__restrict__ col_start,
                                                       long ncols, int ntiles, long m_local, T *local, Geom g, int bulk,
                                                       const unsigned short *__restrict__ cuts, int strict) {
  CM_DYN_SMEM(smem_raw);
  TileSmem<T> &sm = *reinterpret_cast<TileSmem<T> *>(smem_raw);
  constexpr int kTileRows = TileSmem<T>::kTileRows;
  const int tid = threadIdx.x;
  const long nunits = ncols * ntiles;
  for (long unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
    const long c = unit / ntiles;
    const int tile = (int)(unit % ntiles);
    const long lo = col_start[c], hi = col_start[c + 1];
    if (lo == hi) continue;  // uniform: nothing for this column in this flush
    const int r0 = tile * kTileRows;
    const int h = (int)min((long)kTileRows, m_local - r0);
    {
      Vec16<T> z;
#pragma unroll
      for (int k = 0; k < (int)(16 / sizeof(T)); ++k) z.x[k] = T(0);
      Vec16<T> *acc16 = reinterpret_cast<Vec16<T> *>(sm.acc);
      for (int i = tid; i < (h * (int)sizeof(T) + 15) / 16; i += 256) acc16[i] = z;
    }

    for (long base = lo; base < hi; base += kRecChunk) {
      const int cnt = (int)min((long)kRecChunk, hi - base);
      __syncthreads();  // previous chunk fully consumed; acc zeroed
      int special = 0;
      if (tid < kRecChunk) {
        // padding entries (tid >= cnt) alias item 0 with n = 0: loads stay in bounds, nothing is added
        const ItemRec r = recs[base + (tid < cnt ? tid : 0)];
        const T *v = static_cast<const T *>(r.v);
        const int *rows = r.lrow;
        int n = tid < cnt ? r.n : 0;
        const int mf = tid < cnt ? r.mask_flags : 0;
        if (kMulti && n > 0 && (mf >> 8) == 0 && r.row_stride == 1) {
          // taller than one tile: this block owns the rows [r0, r0 + h). An item's rows are strictly
          // increasing, so the ones that fall into the band are a contiguous sub-range: keep only
          // that (two binary searches per staged item instead of a range check per element, and
          // the warps beyond the sub-range have nothing to load)
          int a_lo, a_hi;
          if (cuts) {  // where the item's rows cross the band boundaries was worked out when the records were built
            const unsigned short *ct = cuts + (base + tid) * kMaxCutTiles;
            a_lo = tile == 0 ? 0 : (int)ct[tile - 1];
            a_hi = tile == ntiles - 1 ? r.n : (int)ct[tile];
          } else {
            a_lo = first_row_ge(r.lrow, r.n, r0);
            a_hi = first_row_ge(r.lrow, r.n, r0 + h);
          }
          n = a_hi - a_lo;
          if (n > 0) {  // (an empty sub-range keeps the item's own, valid, base addresses)
            rows = r.lrow + a_lo;
            v = static_cast<const T *>(r.v) + (long)a_lo * r.row_stride;
          }
        }
        sm.v[tid] = v;
        sm.rows[tid] = rows;
        sm.n[tid] = n;
        sm.stride[tid] = r.row_stride;
        sm.mf[tid] = mf;
        special = tid < cnt ? ((kOrdered ? r.mask_flags : (r.mask_flags & 0xff)) | (r.row_stride != 1)) : 0;
      }
      const bool plain = __syncthreads_or(special) == 0;  // no mask, unit stride (ordered walk: no repeated rows either)
      if (!kOrdered) {
        if (plain)
          tile_atomic_chunk<T, false>(sm, cnt, r0, h, 0, tid);
        else
          tile_atomic_chunk<T, true>(sm, cnt, r0, h, mask_threshold(global_index(c, g.nb, g.npcol, g.mypcol), m_local, g), tid);
      } else if (plain) {
        // thread t owns element t of every item; block b+1's elements are in flight while block b
        // is added; one barrier per item keeps the items in order
        int rowA[BLK], rowB[BLK];
        T valA[BLK], valB[BLK];
        unsigned okA = 0, okB = 0;
        const int nblk = (cnt + BLK - 1) / BLK;  // staged padding covers up to kRecChunk
        constexpr int kTilePrefetch = kTilePrefetchBlocks;
        for (int b = 1; b <= kTilePrefetch && b < nblk; ++b) tile_prefetch_block<T, BLK>(sm, b * BLK, tid);
        tile_load_block<T, BLK>(sm, 0, tid, rowA, valA, okA);
        for (int b = 0; b < nblk; b += 2) {
          if (b + 1 + kTilePrefetch < nblk) tile_prefetch_block<T, BLK>(sm, (b + 1 + kTilePrefetch) * BLK, tid);
          if (b + 1 < nblk) tile_load_block<T, BLK>(sm, (b + 1) * BLK, tid, rowB, valB, okB);
          tile_add_block<T, BLK>(sm.acc, r0, h, rowA, valA, okA);
          if (b + 1 < nblk) {
            if (b + 2 + kTilePrefetch < nblk) tile_prefetch_block<T, BLK>(sm, (b + 2 + kTilePrefetch) * BLK, tid);
            if (b + 2 < nblk) tile_load_block<T, BLK>(sm, (b + 2) * BLK, tid, rowA, valA, okA);
            tile_add_block<T, BLK>(sm.acc, r0, h, rowB, valB, okB);
          }
        }
      } else {
        // masked (symmetric update / fill_lower), strided or repeated-row items: simple in-order loop
        const int thr = mask_threshold(global_index(c, g.nb, g.npcol, g.mypcol), m_local, g);
        for (int j = 0; j < cnt; ++j) {
          const T *v = sm.v[j];
          const int *rows = sm.rows[j];
          const int n = sm.n[j], stride = sm.stride[j], mf = sm.mf[j];
          const int mask = mf & 0xff;
          if ((mf >> 8) && !strict) {
            // rows may repeat (an index set that is not strictly increasing) and the order is free: every thread
            // adds its element with a shared-memory atomic
            if (tid < n) {
              const int rr = __ldg(rows + tid) - r0;
              if ((unsigned)rr < (unsigned)h && (mask == 0 || (mask == 1) == (rr + r0 <= thr)))
                atomicAdd(sm.acc + rr, __ldg(v + (long)tid * stride));
            }
          } else if (mf >> 8) {
            if (tid == 0) {  // deterministic mode: one thread applies the item serially, in index order
              for (int a = 0; a < n; ++a) {
                const int rr = rows[a] - r0;
                if ((unsigned)rr >= (unsigned)h) continue;
                if (mask != 0 && (mask == 1) != (rr + r0 <= thr)) continue;
                sm.acc[rr] += v[(long)a * stride];
              }
            }
          } else if (tid < n) {
            const int rr = __ldg(rows + tid) - r0;
            if ((unsigned)rr < (unsigned)h && (mask == 0 || (mask == 1) == (rr + r0 <= thr)))
              sm.acc[rr] += __ldg(v + (long)tid * stride);
          }
          __syncthreads();
        }
      }
    }

    // write-back: local[c, r0 + i] += acc[i]
    T *dcol = local + c * g.lld + r0;
    const bool aligned = (reinterpret_cast<uintptr_t>(dcol) & 15) == 0;
#ifndef CM_EMU
    if (bulk && aligned && ((h * (int)sizeof(T)) & 15) == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my tile writes -> async proxy
      __syncthreads();
      if (tid == 0) {
        bulk_reduce_add(dcol, sm.acc, (unsigned)(h * (int)sizeof(T)));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile has been read: it may be re-zeroed
      }
      __syncthreads();  // the next unit re-zeroes acc
      continue;
    }
#endif
    __syncthreads();
    constexpr int kPer = 16 / (int)sizeof(T);
    if (aligned) {
      // 16 bytes per thread step (LDS.128 / LDG.128 / STG.128), untouched pieces skipped
      const int hv = h / kPer;
      const Vec16<T> *acc16 = reinterpret_cast<const Vec16<T> *>(sm.acc);
      Vec16<T> *d16 = reinterpret_cast<Vec16<T> *>(dcol);
#pragma unroll 4
      for (int i = tid; i < hv; i += 256) {
        const Vec16<T> a = acc16[i];
        bool any = false;
#pragma unroll
        for (int k = 0; k < kPer; ++k) any |= a.x[k] != T(0);
        if (any) {
          Vec16<T> q = d16[i];
#pragma unroll
          for (int k = 0; k < kPer; ++k) q.x[k] += a.x[k];
          d16[i] = q;
        }
      }
      for (int i = hv * kPer + tid; i < h; i += 256)
        if (sm.acc[i] != T(0)) dcol[i] += sm.acc[i];
    } else {
      for (int i = tid; i < h; i += 256)
        if (sm.acc[i] != T(0)) dcol[i] += sm.acc[i];
    }
    __syncthreads();  // the next unit re-zeroes acc
  }
#ifndef CM_EMU
  if (bulk && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all reductions performed
#endif
}

template <typename T, bool kMulti, int BLK, int CTAS, bool kOrdered>
static void launch_tile_variant(const ItemRec *d_recs, const long *d_col_start, long n_local, int ntiles, long m_local,
                                T *d_local, const Geom &g, int max_blocks_per_sm, const unsigned short *d_cuts, int strict,
                                cudaStream_t s) {
  const size_t smem = sizeof(TileSmem<T>);
  static int resident = 0;
  const int bulk = 1;  // (0: per-thread write-back everywhere; kept for misaligned columns and the kernel-logic build)
  if (resident == 0) {
    cudaFuncSetAttribute(k_apply_tile<T, kMulti, BLK, CTAS, kOrdered>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    resident = resident_blocks(k_apply_tile<T, kMulti, BLK, CTAS, kOrdered>, 256, smem);
  }
  const long nunits = n_local * ntiles;
  // max_blocks_per_sm > 0 leaves room on every SM for a kernel of another stream (the pack of
  // the next panel in a pipelined commit)
  const long cap = max_blocks_per_sm > 0 ? (long)max_blocks_per_sm * kernel_sm_count() : (1L << 40);
  CM_LAUNCH((unsigned)std::min({nunits, (long)resident, cap}), 256, smem, s, k_apply_tile<T, kMulti, BLK, CTAS, kOrdered>)(
      d_recs, d_col_start, n_local, ntiles, m_local, d_local, g, bulk, d_cuts, strict);
}

template <typename T>
static void launch_apply_tile_t(const ItemRec *d_recs, const long *d_col_start, long n_local, long m_local, T *d_local,
                                const Geom &g, int max_blocks_per_sm, int ordered, const unsigned short *d_cuts, cudaStream_t s) {
  const int ntiles = tile_count(sizeof(T) == 8 ? 0 : 1, m_local);
  if (ntiles > kMaxCutTiles) d_cuts = nullptr;
  // 6-item register sets, 3 blocks per SM (measured best on B200; 8- and 12-item sets spill or lose residency).
  // Which walk (B200, one launch): one tile per column (K1) ordered 1.71 ms, item-parallel 2.01; two row bands
  // (K2 owner) 0.486 / 0.524 ms; four (NS owner) 2.44 / 2.18 ms — the atomics cost more per element than the
  // barrier per item until a block's share of an item falls to a quarter of its rows.
  // (a 64-register build that lets three scatter-add blocks and a pack block of the next panel share an SM was
  // measured at 2 GPUs: the scatter-add gained, the pack kernel lost more — both live on the L1 / LSU pipe)
  if (ntiles == 1)
    launch_tile_variant<T, false, 6, 3, true>(d_recs, d_col_start, n_local, ntiles, m_local, d_local, g, max_blocks_per_sm, d_cuts, ordered, s);
  else if (ordered || ntiles < 3)
    launch_tile_variant<T, true, 6, 3, true>(d_recs, d_col_start, n_local, ntiles, m_local, d_local, g, max_blocks_per_sm, d_cuts, ordered, s);
  else
    launch_tile_variant<T, true, 6, 3, false>(d_recs, d_col_start, n_local, ntiles, m_local, d_local, g, max_blocks_per_sm, d_cuts, ordered, s);
}

void launch_apply_tile(int dtype, const ItemRec *d_recs, const long *d_col_start, long nitems, void *d_local,
                       long n_local, long m_local, const Geom &g, int max_blocks_per_sm, int ordered,
                       const unsigned short *d_cuts, cudaStream_t s) {
  if (nitems <= 0) return;
  switch (dtype) {
    case 0: launch_apply_tile_t(d_recs, d_col_start, n_local, m_local, static_cast<double *>(d_local), g, max_blocks_per_sm, ordered, d_cuts, s); break;
    case 1: launch_apply_tile_t(d_recs, d_col_start, n_local, m_local, static_cast<float *>(d_local), g, max_blocks_per_sm, ordered, d_cuts, s); break;
    default: launch_apply_tile_t(d_recs, d_col_start, n_local, m_local, static_cast<int *>(d_local), g, max_blocks_per_sm, ordered, d_cuts, s); break;
  }
}

// ------------------------------------------------------------------------------------
// COO path (elemental updates, reference :623-629 -> :183)
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_apply_coo(const long *__restrict__ inds, const T *__restrict__ vals, long n,
                                                   T *__restrict__ local) {
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x)
    red_add(local + inds[k], vals[k]);
}

void launch_apply_coo(int dtype, const long *d_inds, const void *d_vals, long n, void *d_local, cudaStream_t s) {
  if (n <= 0) return;
  long blocks = (n + 255) / 256;
  const long cap = (long)kernel_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  switch (dtype) {
    case 0:
      CM_LAUNCH((unsigned)blocks, 256, 0, s, k_apply_coo<double>)(d_inds, static_cast<const double *>(d_vals), n,
                                                          static_cast<double *>(d_local));
      break;
    case 1:
      CM_LAUNCH((unsigned)blocks, 256, 0, s, k_apply_coo<float>)(d_inds, static_cast<const float *>(d_vals), n,
                                                         static_cast<float *>(d_local));
      break;
    default:
      CM_LAUNCH((unsigned)blocks, 256, 0, s, k_apply_coo<int>)(d_inds, static_cast<const int *>(d_vals), n,
                                                       static_cast<int *>(d_local));
      break;
  }
}


// deterministic COO: stable sort of (destination, arrival position); the thread at the head
// of each run adds the run's values in arrival order
__global__ void k_iota_u32(uint32_t *out, long n) {
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x) out[k] = (uint32_t)k;
}

template <typename T>
__global__ void __launch_bounds__(256) k_apply_coo_det(const unsigned long long *__restrict__ sorted_inds,
                                                       const uint32_t *__restrict__ perm, const T *__restrict__ vals,
                                                       long n, T *local) {
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x) {
    const unsigned long long ij = sorted_inds[k];
    if (k > 0 && sorted_inds[k - 1] == ij) continue;  // not a run head
    T acc = local[ij];
    for (long t = k; t < n && sorted_inds[t] == ij; ++t) acc += vals[perm[t]];
    local[ij] = acc;
  }
}

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

size_t coo_det_temp_bytes(long n) {
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, n);
  return align256(sort_bytes) + align256(8 * (size_t)n) + 2 * align256(4 * (size_t)n);
}

void launch_apply_coo_det(int dtype, const long *d_inds, const void *d_vals, long n, void *d_local, void *d_temp,
                          size_t temp_bytes, cudaStream_t s) {
  if (n <= 0) return;
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, n);
  char *base = static_cast<char *>(d_temp);
  void *sort_tmp = base;
  unsigned long long *sorted = reinterpret_cast<unsigned long long *>(base + align256(sort_bytes));
  uint32_t *iota = reinterpret_cast<uint32_t *>(base + align256(sort_bytes) + align256(8 * (size_t)n));
  uint32_t *perm = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(iota) + align256(4 * (size_t)n));
  (void)temp_bytes;
  long blocks = (n + 255) / 256;
  const long cap = (long)kernel_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  CM_LAUNCH((unsigned)blocks, 256, 0, s, k_iota_u32)(iota, n);
  cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, reinterpret_cast<const unsigned long long *>(d_inds), sorted, iota,
                                  perm, n, 0, 64, s);
  switch (dtype) {
    case 0:
      CM_LAUNCH((unsigned)blocks, 256, 0, s, k_apply_coo_det<double>)(sorted, perm, static_cast<const double *>(d_vals), n,
                                                              static_cast<double *>(d_local));
      break;
    case 1:
      CM_LAUNCH((unsigned)blocks, 256, 0, s, k_apply_coo_det<float>)(sorted, perm, static_cast<const float *>(d_vals), n,
                                                             static_cast<float *>(d_local));
      break;
    default:
      CM_LAUNCH((unsigned)blocks, 256, 0, s, k_apply_coo_det<int>)(sorted, perm, static_cast<const int *>(d_vals), n,
                                                           static_cast<int *>(d_local));
      break;
  }
}

// global indices of local indices 0..count-1 on process p (:697, :699)
__global__ void k_iota_global(long *out, long count, long blk, long np, long p) {
  for (long l = (long)blockIdx.x * blockDim.x + threadIdx.x; l < count; l += (long)gridDim.x * blockDim.x)
    out[l] = global_index(l, blk, np, p);
}

void launch_iota_global(long *d_out, long count, long blk, long np, long p, cudaStream_t s) {
  if (count <= 0) return;
  long blocks = (count + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  CM_LAUNCH((unsigned)blocks, 256, 0, s, k_iota_global)(d_out, count, blk, np, p);
}

}  // namespace cmb
