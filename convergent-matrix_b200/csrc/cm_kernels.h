// cm_kernels.h — launchers of the sm_100a kernels (cm_kernels.cu). Host-callable C++.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "cm_common.h"

namespace cmb {

// one received (or self) message as the owner sees it
struct MsgRef {
  const char *meta;  // WireHdr, WireSeg[nseg], int32 idx[nidx]
  const void *vals;  // nvals elements
  long seg_base;     // first entry of this message in the owner's Seg table
  long item_base;    // first item of this message in the owner's item arrays
};

// device-side views of the per-flush index arenas (all int32, see k_map_group)
struct MapArenas {
  int *rl;    // grouped local row indices        [sum m_u]
  int *rpos;  // original position of each grouped row
  int *roff;  // group starts, (nprow+1) per update
  int *cl;    // grouped local column indices     [sum n_u]
  int *cpos;
  int *coff;  // (npanel*npcol+1) per update: column groups gq = panel * npcol + owner column
  int *uflags;  // per update: kSegRowsMayRepeat if ix is not strictly increasing (zeroed by the caller)
  int *radj;    // (nprow+1) per update: rows of each owner row that directly follow the update's previous index
};

// per-(owner, update) exclusive scans produced by k_scan_dest
struct ScanArenas {
  long *valoff;    // [npanel * P * nupd], index (panel * P + owner) * nupd + update
  long *idxoff;    // same
  long *itemoff;   // same
  MsgCounts *totals;  // [kMaxChunks * P], index panel * P + owner
  void **dst_vals;    // [kMaxChunks * P] where this rank's message (panel, owner) starts: values ...
  char **dst_meta;    // ... and metadata. May point into a PEER's receive arena (NVLink stores).
};

size_t dtype_size(int dtype);

// (1) owner map + stable grouping by owner row / column (warp match + prefix sums)
void launch_map_group(const UpdDesc *d_descs, int nupd, const Geom &g, const PanelPlan &pp, const MapArenas &ma,
                      cudaStream_t s);
// (2) per-owner scans over the staged updates (sizes of every message)
void launch_scan(int nupd, const PanelPlan &pp, const Geom &g, const MapArenas &ma, const ScanArenas &sa,
                 const long *d_coo_counts, cudaStream_t s);
// (3) pack: gather Mat[R_p, C_q] into owner (p,q)'s segment + write the wire metadata, straight
// to sa.dst_vals / sa.dst_meta (local staging, or the owner's receive arena over NVLink)
// Packs the column panels [k0, k1) of every staged update.
// max_rows: largest m_u among the staged updates (sizes the block).
void launch_pack(int dtype, const UpdDesc *d_descs, int nupd, int k0, int k1, const PanelPlan &pp,
                 const Geom &g, const MapArenas &ma, const ScanArenas &sa, long max_rows, cudaStream_t s);
// (4a) single-rank path: segments point straight at the staged payloads (no pack)
void launch_build_direct(int dtype, const UpdDesc *d_descs, const long *d_item_base, int nupd,
                         const Geom &g, const MapArenas &ma, Seg *d_segs, uint32_t *d_keys,
                         uint64_t *d_items, cudaStream_t s);
// (4b) owner side: turn the received messages into Seg entries + (key, item) pairs
void launch_build_recv(int dtype, const MsgRef *d_msgs, int nmsgs, long max_seg_per_msg,
                       Seg *d_segs, uint32_t *d_keys, uint64_t *d_items, cudaStream_t s);
// (5) stable radix sort of the items by destination column (CUB)
size_t sort_temp_bytes(long nitems, int end_bit);
void launch_sort(void *d_temp, size_t temp_bytes, const uint32_t *keys_in, uint32_t *keys_out,
                 const uint64_t *items_in, uint64_t *items_out, long nitems, int end_bit,
                 cudaStream_t s);
// (5b) sorted items -> ItemRec (absolute pointers) + start of every destination column's run
int tile_count(int dtype, long m_local);  // column tiles stacked over the local rows
// d_cuts (cuts_bytes() > 0: local blocks of several row bands): where every item's rows cross the band boundaries
size_t cuts_bytes(int dtype, long m_local, long nitems);
void launch_build_recs(int dtype, const Seg *d_segs, const uint32_t *d_keys, const uint64_t *d_items,
                       long nitems, ItemRec *d_recs, long n_local, long *d_col_start, unsigned short *d_cuts,
                       long m_local, cudaStream_t s);
// (6a) scatter-add with RED.ADD, one warp per item (apply<T>, reference :178-187)
void launch_apply_red(int dtype, const ItemRec *d_recs, long nitems, long elements, void *d_local, const Geom &g,
                      cudaStream_t s);
// (6b) column-tile scatter-add in shared memory. ordered: items strictly in sorted order, no atomics (bit-exact run
// to run); otherwise item-parallel with shared-memory atomics (order-free, faster on tall blocks)
void launch_apply_tile(int dtype, const ItemRec *d_recs, const long *d_col_start, long nitems,
                       void *d_local, long n_local, long m_local, const Geom &g, int max_blocks_per_sm,
                       int ordered, const unsigned short *d_cuts, cudaStream_t s);
// COO path for elemental updates: local[inds[k]] += vals[k]
void launch_apply_coo(int dtype, const long *d_inds, const void *d_vals, long n, void *d_local,
                      cudaStream_t s);
// deterministic COO: stable sort by destination, one thread sums each run in arrival order
size_t coo_det_temp_bytes(long n);
void launch_apply_coo_det(int dtype, const long *d_inds, const void *d_vals, long n, void *d_local,
                          void *d_temp, size_t temp_bytes, cudaStream_t s);
// fill_lower helpers: global indices of this rank's local columns / rows (:697, :699)
void launch_iota_global(long *d_out, long count, long blk, long np, long p, cudaStream_t s);

int kernel_sm_count();

}  // namespace cmb
