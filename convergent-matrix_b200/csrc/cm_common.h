// cm_common.h — structures shared by the host runtime and the sm_100a kernels.
//
// Vocabulary (follows the reference, swfrench/convergent-matrix):
//   update   one slice update(Mat, ix, jx)            include/convergent_matrix.hpp:604-615
//   owner    the rank whose local block holds (i,j)    :586, :609
//   segment  the part of one update owned by one rank: the dense sub-block Mat[R_p, C_q]
//            (R_p / C_q = rows / columns of the update whose owner row / column is p / q,
//            in original order) — exactly the elements the reference appends to bin
//            (p,q) for that update (:611), in the same column-major order
//   message  everything one source ships to one owner in one commit (reference: the
//            XBuff of one Bin::flush, :128-133, :220-245)
//   item     one column chunk of one segment: the unit of work of the scatter-add
#pragma once

#include <cuda_runtime.h>  // __host__ / __device__ (implicit under nvcc; explicit for host-compiled builds)
#include <stdint.h>

namespace cmb {

constexpr int kMaxGridDim = 64;     // max NPROW, NPCOL handled by the grouping kernel
constexpr int kRowChunk = 256;      // rows of one segment column covered by one item (= block size)
constexpr int kItemSegBits = 26;    // item encoding: seg | column | row chunk
constexpr int kItemColBits = 22;
constexpr int kItemChunkBits = 16;

// distribution geometry (template longs of the reference, :298-302) + this rank's coords
struct Geom {
  long m, n;
  long mb, nb;
  long nprow, npcol;
  long lld;
  int myprow, mypcol;  // row-major grid: rank = mypcol + npcol * myprow
};

// one staged slice update; built on the host, read by the map / pack kernels
struct UpdDesc {
  const void *data;  // DEVICE pointer to the payload
  const long *ix;    // DEVICE pointer, m global row indices
  const long *jx;    // DEVICE pointer, n global column indices
  long ld;           // leading dimension of the payload storage
  int m, n;          // VIEW dims (LocalMatrix::m(), n(); include/local_matrix.hpp:52-54)
  int trans;         // element (i,j) = trans ? data[j + i*ld] : data[i + ld*j]  (:71)
  int mask;          // CM_MASK_*
  long rbase;        // offset of this update's mapped row list in the row arenas
  long cbase;        // offset of this update's mapped column list in the column arenas
};

// one segment as the scatter-add sees it (absolute device pointers)
struct Seg {
  const void *vals;  // element (a,b) of the segment at vals[a*row_stride + b*col_stride]
  const int *lrow;   // nr local row indices    (:610  (ix/(MB*NPROW))*MB + ix%MB)
  const int *lcol;   // nc local column indices (:607  (jx/(NB*NPCOL))*NB + jx%NB)
  long row_stride;
  long col_stride;
  int nr, nc;
  int mask;
  int flags;  // kSegRowsMayRepeat
};
constexpr int kSegRowsMayRepeat = 1;  // row index set not strictly increasing (duplicates possible)

// one item after the sort, with absolute pointers (what the scatter-add kernels read)
struct ItemRec {
  const void *v;    // first value of the item
  const int *lrow;  // its local row indices
  int n;            // rows in the item (<= kRowChunk)
  int row_stride;   // elements between consecutive rows in v (1 unless a transposed direct payload)
  int mask_flags;   // mask | flags << 8
  int col;          // destination local column
};
static_assert(sizeof(ItemRec) == 32, "item record layout");

// ---- wire format of one message (metadata part); all offsets relative to the message ----
// Instead of the reference's 8-byte index per element (XBuff::inds, :131) a segment ships
// its nr local row indices and nc local column indices once; the per-element index stream
// is recoverable exactly: inds[a + b*nr] = LLD*lcol[b] + lrow[a].
struct WireHdr {
  long nseg;    // segments in this message (= updates staged by the source this commit)
  long nvals;   // total values
  long nidx;    // total int32 index words
  long nitems;  // total scatter-add items
};
struct WireSeg {
  int nr, nc;
  int mask;
  int flags;
  long val_off;   // first value of the segment (elements, within the message's values)
  long idx_off;   // first index word: lrow[nr] then lcol[nc]
  long item_off;  // first item of the segment within the message
};
static_assert(sizeof(WireHdr) == 32, "wire header layout");
static_assert(sizeof(WireSeg) == 40, "wire segment layout");

// Column panels. A source that stages a large commit cuts every update's columns into up to
// kMaxChunks panels of `panel_cols` consecutive LOCAL columns of the owner; every (source, panel,
// owner) triple is one message. The owner's scatter-add of panel k touches only panel k's columns
// of the local block, so a large commit is pipelined panel by panel (the scatter-add of panel k
// overlaps the pack / exchange of panel k+1) without sweeping the local block more than once. A
// small commit has one panel: one message per (source, owner), the reference's XBuff (:128-133).
// Implementation: panels widen the column-group key of the grouping kernel from the owner column q
// to gq = panel * NPCOL + q ("virtual owner columns"); the rest of the path sees npanel * NPCOL
// column groups per update.
constexpr int kMaxChunks = 8;
struct PanelPlan {
  int npanel;       // 1 .. kMaxChunks, and npanel * NPCOL <= kMaxGridDim
  long panel_cols;  // local columns per panel, a multiple of NB (the last panel takes the rest)
};
__host__ __device__ inline int panel_of(const PanelPlan &pp, long lcol) {
  if (pp.npanel <= 1) return 0;
  const long k = lcol / pp.panel_cols;
  return (int)(k < pp.npanel ? k : pp.npanel - 1);
}

// counts exchanged before the data exchange, per (source, panel, owner) triple
struct MsgCounts {
  long nvals;   // value SLOTS of the message (segments are padded, see seg_col_stride)
  long nidx, nseg, nitems;
  long nelems;  // update elements carried (what the scatter-add will add)
  long ncoo;    // elemental updates (COO lane) from this source to this owner; panel 0 only
  long nflushed;  // messages this source's threshold flushes packed for this owner since the last commit
                  // (reference flush(thresh) inside update(), :346, :614); panel 0 only
  long nadj;      // of nelems: elements whose local row directly follows the previous row of their segment
};

// Layout of one segment's values inside a message: column-major nr x nc with column stride
// seg_col_stride(nr). Tall segments pad every column to a multiple of 16 slots and start on a
// 16-slot boundary, so that the pack kernel's half-warps store whole, aligned 128-byte lines (fp64)
// into the owner's arena: peer stores that straddle lines cost ~25 % of the NVLink bandwidth
// (measured, B200: 532 vs 655 GB/s per direction). Short segments stay dense.
constexpr int kSegAlign = 16;
__host__ __device__ inline int seg_col_stride(int nr) { return nr >= 64 ? (nr + kSegAlign - 1) & ~(kSegAlign - 1) : nr; }
__host__ __device__ inline long seg_slots(int nr, int nc) {
  return ((long)seg_col_stride(nr) * nc + kSegAlign - 1) & ~(long)(kSegAlign - 1);
}

__host__ __device__ inline long meta_bytes(long nseg, long nidx) {
  long b = (long)sizeof(WireHdr) + nseg * (long)sizeof(WireSeg) + 4 * nidx;
  return (b + 15) & ~15L;
}

// (:610 / :607) local index of global index g along one axis
__host__ __device__ inline long local_index(long g, long blk, long np) {
  return (g / (blk * np)) * blk + g % blk;
}
// (:609 / :606) owner coordinate of global index g along one axis
__host__ __device__ inline int owner_coord(long g, long blk, long np) {
  return (int)((g / blk) % np);
}
// inverse of (owner_coord, local_index): global index of local index l on process p
// (the reconstruction fill_lower uses, :697, :699)
__host__ __device__ inline long global_index(long l, long blk, long np, long p) {
  return (l / blk) * (np * blk) + p * blk + l % blk;
}

}  // namespace cmb
