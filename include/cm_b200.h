/* include/cm_b200.h — C ABI of libcm_b200.so, the B200 (sm_100a) implementation of
 * ConvergentMatrix's assembly hot path: update -> owner map / pack -> exchange ->
 * scatter-add -> commit.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types. The
 * header-only C++ templates in include/convergent_matrix.hpp / include/local_matrix.hpp
 * (same names and signatures as the reference's) are thin wrappers over these entry
 * points. Each entry point cites the reference interface it replaces; citations are into
 * swfrench/convergent-matrix include/convergent_matrix.hpp unless another file is named.
 *
 * Every function returns CM_OK (0) or a CM_ERR_* code; cm_last_error() gives the message
 * of the calling thread's last failure. There is NO CPU fallback: every entry point that
 * touches matrix data fails with CM_ERR_CUDA when no sm_100-class GPU / CUDA driver is
 * usable. The only calls that work without a GPU are the pure integer helpers
 * (cm_numroc, cm_owner_of, cm_descinit_raw, cm_version, cm_last_error).
 */
#ifndef CM_B200_H
#define CM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ---------------------------------------------------------------- */
#define CM_OK 0
#define CM_ERR_INVALID 1 /* bad argument / contract violation (the reference asserts)  */
#define CM_ERR_CUDA 2    /* CUDA runtime / driver failure, or no GPU                   */
#define CM_ERR_NCCL 3    /* NCCL failure or libnccl.so.2 not loadable                  */
#define CM_ERR_STATE 4   /* runtime not initialised / wrong mode / rank not bound      */
#define CM_ERR_NOMEM 5   /* device or pinned-host allocation failed                    */

/* ---- element types (template parameter T of the reference, :298) ----------------- */
#define CM_F64 0
#define CM_F32 1
#define CM_I32 2

/* ---- update masks (how one slice update is filtered on the owner) ----------------- */
#define CM_MASK_NONE 0        /* general update, :604-615                               */
#define CM_MASK_UPPER 1       /* keep elements with global row <= global col, :654      */
#define CM_MASK_STRICT_LOWER 2 /* keep elements with global row > global col (fill_lower
                                  targets, :701)                                        */

const char *cm_version(void);
const char *cm_last_error(void);

/* ====================================================================================
 * Runtime: replaces upcxx::init / finalize / rank_me / rank_n / barrier
 * (reference test/cm_test.cpp:367,389; include/convergent_matrix.hpp:84,409,421,741).
 * One process per GPU (cm_rt_init) is the production mode; ranks exchange over NCCL.
 * ==================================================================================== */

/* Rank 0 creates the 128-byte NCCL unique id; the launcher (torchrun + torch.distributed
 * broadcast, or the file rendezvous in include/convergent_matrix.hpp) hands the same bytes
 * to every rank's cm_rt_init. len must be >= 128. */
int cm_rt_unique_id(void *out, size_t len);

/* Bind this process to `rank` of `nranks` on CUDA device `device`. unique_id may be NULL
 * when nranks == 1 (no NCCL is loaded then). Collective over all ranks when nranks > 1. */
int cm_rt_init(int rank, int nranks, int device, const void *unique_id);

/* Convenience for SPMD C++ programs started by torchrun (or any launcher that exports
 * RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT): rank 0 creates the unique id and publishes
 * it through a file ($CM_B200_UID_FILE or /tmp/cm_b200_uid.<MASTER_PORT>.<ppid>), the other
 * ranks wait for it; then cm_rt_init(RANK, WORLD_SIZE, LOCAL_RANK, id). */
int cm_rt_init_from_env(void);

/* TEST TRANSPORT: all `nranks` ranks live in THIS process on ONE device, one host thread
 * per rank (threads-as-ranks, like the GASNet smp conduit the reference's tests use,
 * make/flags.local.mk:5). The exchange is device-to-device copies on the same GPU, the
 * kernels are the production ones. Each rank thread calls cm_rt_bind_rank(r) once. */
int cm_rt_init_loopback(int nranks, int device);
int cm_rt_bind_rank(int rank);

int cm_rt_finalize(void);
int cm_rt_rank_me(void); /* upcxx::rank_me(); -1 if not initialised */
int cm_rt_rank_n(void);  /* upcxx::rank_n();  -1 if not initialised */
int cm_rt_barrier(void); /* upcxx::barrier(), :741 */
int cm_rt_device(void);

/* ====================================================================================
 * ConvergentMatrix<T,NPROW,NPCOL,MB,NB,LLD>  (:298-971). Template longs become runtime
 * arguments. Grid coordinates are row-major: prow = rank / NPCOL, pcol = rank % NPCOL
 * (what the owner formula :586 implies; the reference ctor's rank / NPROW at :409 is a
 * defect for NPROW != NPCOL that update/commit never read — SURVEY.md Appendix C-1).
 * ==================================================================================== */
typedef struct cm_matrix cm_matrix;

/* ctor, :405-448 — COLLECTIVE. Allocates and zero-fills the LLD x n_local column-major
 * local block in HBM. Fails (the reference asserts, :417-428) unless m,n > 0,
 * nprow*npcol == cm_rt_rank_n(), m_local,n_local > 0 and m_local <= lld. */
int cm_create(cm_matrix **out, int dtype, long m, long n, long nprow, long npcol, long mb,
              long nb, long lld);
/* dtor, :457-460 — COLLECTIVE; performs a final commit. */
int cm_destroy(cm_matrix *M);

/* update(LocalMatrix<T>* Mat, long* ix, long* jx), :604-615. HOST pointers. (m_u, n_u) are
 * the VIEW dims Mat->m(), Mat->n(); element (i,j) = trans ? data[j + i*ld] : data[i + ld*j]
 * (include/local_matrix.hpp:71). Values and indices are copied before the call returns
 * (reference :611), so the caller may reuse data/ix/jx immediately. Global indices are
 * not range-checked (neither does the reference); out-of-range indices are undefined. */
int cm_update_slice(cm_matrix *M, const void *data, long m_u, long n_u, long ld, int trans,
                    const long *ix, const long *jx);
/* update(LocalMatrix<T>* Mat, long* ix), :641-680 — symmetric: only ix[i] <= ix[j]. */
int cm_update_sym(cm_matrix *M, const void *data, long n_u, long ld, int trans,
                  const long *ix);
/* update(T elem, long ix, long jx), :623-629. val points to one element of type T. */
int cm_update_elem(cm_matrix *M, const void *val, long i, long j);
/* n elemental updates in one call (same semantics as n calls of cm_update_elem). */
int cm_update_elems(cm_matrix *M, const void *vals, const long *is, const long *js, long n);
/* fill_lower(), :694-712. Call commit() before and after, as in the reference (:690-692). */
int cm_fill_lower(cm_matrix *M);

/* Device-resident variant of update(): d_data, d_ix, d_jx are DEVICE pointers on this
 * rank's GPU. The payload and index arrays are BORROWED (stream-ordered) until the next
 * cm_flush()/cm_commit() on this matrix returns — on any number of ranks: a flush packs (or
 * scatter-adds) everything staged before it returns; mask is one of CM_MASK_*. This is the
 * zero-copy path behind cm::DeviceLocalMatrix. */
int cm_update_slice_device(cm_matrix *M, const void *d_data, long m_u, long n_u, long ld,
                           int trans, const long *d_ix, const long *d_jx, int mask);

/* Batch forms: the effect of `count` consecutive cm_update_slice_device / cm_update_slice calls
 * (arrays of per-update arguments; trans may be NULL = all 0). The host form returns when every
 * payload of the batch has been copied, so the copies of one batch overlap each other on PCIe. */
int cm_update_slices_device(cm_matrix *M, long count, const void *const *d_data,
                            const long *m_u, const long *n_u, const long *ld, const int *trans,
                            const long *const *d_ix, const long *const *d_jx, int mask);
int cm_update_slices(cm_matrix *M, long count, const void *const *data, const long *m_u,
                     const long *n_u, const long *ld, const int *trans, const long *const *ix,
                     const long *const *jx);

/* Everything staged so far on THIS rank leaves the staging area (the reference's flush(thresh),
 * :330-375, with thresh = 0). NOT collective. With one rank the staged updates are mapped and
 * scatter-added. With several they are mapped and packed into one message per owner, which the pack
 * kernel stores straight into the owner's flush region of this rank over NVLink (the reference's rput,
 * :168-171; regions exist from the end of the first commit that carried a flushed message, and hold four
 * flushes per (owner, source) pair) or, without a region with room, into buffers of this rank that the
 * next commit delivers; the owner scatter-adds flushed messages inside its next commit, in front of that
 * commit's own (the reference: inside its next progress()). On return the staged payloads, index sets
 * and descriptors are released, borrowed device payloads included. update() calls this by itself when
 * more than bin_flush_threshold elements are staged (:346, :614). Elemental updates stay in host bins. */
int cm_flush(cm_matrix *M);
/* commit(), :723-742 — COLLECTIVE: pack, exchange (all-to-all-v), scatter-add, barrier.
 * On return every update issued by ANY rank before its matching commit() is applied. */
int cm_commit(cm_matrix *M);
/* reset(), :491-501 — COLLECTIVE: commit, zero-fill, barrier. */
int cm_reset(cm_matrix *M);

/* get_local_data(), :469 — the live LLD x n_local column-major block.
 *   _device: the HBM pointer (PBLAS-compatible layout, usable by cuSOLVERMp-style code).
 *   _host:   a host mirror refreshed by a D2H copy when stale; valid until cm_destroy. */
int cm_local_data_device(cm_matrix *M, const void **out);
int cm_local_data_host(cm_matrix *M, const void **out);
/* get_local_data_copy(), :478-483 — copies lld*n_local elements into caller memory. */
int cm_local_data_copy(cm_matrix *M, void *dst_host);
/* operator()(ix, jx), :585-596 — one-sided read of one element; the owner may be another
 * rank (its block is mapped through CUDA IPC at cm_create, the analogue of the cached
 * dist_object::fetch + rget). out points to one element of type T. */
int cm_get(cm_matrix *M, long i, long j, void *out);

/* save(fname) / load(fname), :788-873 / :888-967 — COLLECTIVE. The file is what the reference's
 * MPI-IO darray view writes: the dense global matrix, column-major, native endianness, no header
 * (element (i,j) at byte (i + j*M)*sizeof(T)). No MPI is involved: every rank pwrites / preads the
 * runs of rows it owns; the path must be visible to all ranks. As in the reference there is no
 * implicit commit before save (:778-779), and load replaces the local block's contents (:879). */
int cm_save(cm_matrix *M, const char *fname);
int cm_load(cm_matrix *M, const char *fname);

/* accessors, :550-576 */
long cm_m(const cm_matrix *M);
long cm_n(const cm_matrix *M);
long cm_m_local(const cm_matrix *M);
long cm_n_local(const cm_matrix *M);
long cm_lld(const cm_matrix *M);
long cm_pgrid_row(const cm_matrix *M);
long cm_pgrid_col(const cm_matrix *M);
int cm_dtype(const cm_matrix *M);

/* bin_flush_threshold, :507-514 — number of staged update ELEMENTS above which update()
 * triggers cm_flush (strict '>', :346), on any number of ranks. Default CM_DEFAULT_FLUSH_THRESHOLD
 * (2^30 elements, not the reference's 10000: batches are what make the sorted sweep pay); a per-rank
 * setting. Independently of it a rank flushes before it stages its 2^26-th update. */
#define CM_DEFAULT_FLUSH_THRESHOLD (1L << 30)
long cm_get_flush_threshold(const cm_matrix *M);
int cm_set_flush_threshold(cm_matrix *M, long elements);
/* progress_interval, :529-545 — stored for API compatibility; there is no progress engine. */
int cm_get_progress_interval(const cm_matrix *M);
int cm_set_progress_interval(cm_matrix *M, int interval);
/* deterministic-order mode: accumulation order fixed per element => bit-exact run to run. Off (default), a
 * local block taller than two shared-memory tiles (16384 fp64 rows) is scatter-added item-parallel with
 * shared-memory atomics: the order in which updates meet on an element is then not fixed (neither is it in
 * the reference, README.md:30-32); results agree within rounding. */
int cm_set_deterministic(cm_matrix *M, int on);
/* sort work items by destination column before the scatter-add (L2-resident sweep). on by
 * default; exposed so benches can show its effect. */
int cm_set_sorted_sweep(cm_matrix *M, int on);

/* which scatter-add kernel a flush uses: 0 = auto, 1 = always RED.ADD, 2 = always column tiles in shared
 * memory. Auto takes tiles when the flush holds >= tile_density * 26.8 / (26.8 - 25 c) update elements per element
 * of the local block it can reach, c being the share of its elements whose row directly follows the previous row of
 * their segment (RED pays per 32-byte sector, so runs of consecutive rows make it cheap): tile_density itself for
 * scattered rows, about 12 x tile_density for fully clustered ones. tile_density <= 0 keeps the current value
 * (default 0.125). Deterministic mode always uses column tiles. */
int cm_set_apply_kernel(cm_matrix *M, int kernel, double tile_density);

/* When the owner scatter-adds what commit() delivered (COLLECTIVE, nothing staged; extension — the
 * reference applies inside progress(), :178-187, and commit() waits for it, :737-738).
 * CM_APPLY_EAGER (default): inside commit(), as the reference's commit() guarantees.
 * CM_APPLY_DEFERRED (several ranks, peer-to-peer exchange modes): commit() returns when every message
 * has been DELIVERED into the owners' receive arenas, which become an append-only log; the owner
 * applies its log when its local block is read (cm_local_data_*, cm_get of an own element,
 * cm_fill_lower, cm_save), when the log of any owner exceeds budget_bytes (everybody applies, inside
 * that commit) or at cm_materialize / cm_reset / cm_load / cm_destroy. One scatter-add over several
 * commits' messages sweeps the local block once instead of once per commit.
 * One-sided reads of ANOTHER rank's element (cm_get) fail with CM_ERR_STATE while that rank may hold
 * unapplied updates: call cm_materialize first. budget_bytes <= 0 keeps the current budget (8 GiB). */
#define CM_APPLY_EAGER 0
#define CM_APPLY_DEFERRED 1
/* CM_APPLY_BACKGROUND: as DEFERRED, but once enough elements are logged (16 Mi; CM_B200_BG_MIN_ELEMS) the
 * owner starts their scatter-add on a second stream at the end of the commit and returns: it runs
 * while the following commits pack and exchange. Same rules for readers as DEFERRED. */
#define CM_APPLY_BACKGROUND 2
int cm_set_apply_policy(cm_matrix *M, int policy, long budget_bytes);
int cm_get_apply_policy(const cm_matrix *M);
/* COLLECTIVE: every rank applies its log, then a barrier; afterwards every local block is current. */
int cm_materialize(cm_matrix *M);

/* How commit() moves the per-owner messages (COLLECTIVE: same call on every rank, nothing staged):
 * 0 = auto, 1 = grouped ncclSend/ncclRecv all-to-all-v from a local send arena, 2 = peer-to-peer
 * stores: the pack kernel writes each message straight into the owner's receive arena over NVLink
 * (arenas mapped with CUDA IPC), so packing and exchange are one kernel. Auto picks 2 when every rank
 * can map every other rank's memory, else 1. The environment variable CM_B200_EXCHANGE=nccl|p2p sets
 * the default. cm_get_exchange_mode returns the mode in effect (1 or 2). */
int cm_set_exchange_mode(cm_matrix *M, int mode);
int cm_get_exchange_mode(const cm_matrix *M);

/* PBLAS array descriptor of the local block: {1, ictxt, M, N, MB, NB, 0, 0, LLD}. The
 * reference exposes only the accessors (:550-576); this is the descinit a user writes. */
int cm_descinit(const cm_matrix *M, int ictxt, int desc[9]);

/* ---- observability ----------------------------------------------------------------- */
typedef struct cm_stats {
  long updates;            /* slice updates accepted                                   */
  long elements_staged;    /* update elements accepted (sum of m_u*n_u, incl. masked)  */
  long elements_applied;   /* elements this rank's scatter-add consumed                */
  long flushes;            /* local flushes                                            */
  long commits;
  long kernel_launches;    /* launches of libcm_b200's own kernels                     */
  long bytes_sent;         /* exchange bytes sent to OTHER ranks (values + metadata)    */
  long bytes_received;
  long h2d_bytes;          /* host->device bytes moved by the host-pointer API         */
  long d2h_bytes;
  double apply_ms;         /* CUDA-event time of scatter-add kernels                    */
  double pack_ms;          /* map + pack kernels                                        */
  double exchange_ms;      /* NCCL / loopback exchange                                  */
  long apply_launches;
  long apply_elements;     /* elements covered by the timed scatter-add launches        */
  long tile_launches;      /* scatter-add launches that used the column-tile kernel     */
  long flushed_shipped;    /* messages of threshold flushes stored into their owner's flush region at flush time */
  long flushed_held;       /* ... packed into buffers of this rank and delivered by the next commit */
  long staging_peak_bytes; /* high-water mark of the host-API staging arenas (payload copies + index sets) */
} cm_stats;
int cm_get_stats(cm_matrix *M, cm_stats *out);
int cm_reset_stats(cm_matrix *M);
/* enable CUDA-event timing of each kernel class (costs a few us per flush) */
int cm_set_profiling(cm_matrix *M, int on);
/* the CUDA stream (cudaStream_t) every kernel of this matrix is launched on */
int cm_stream(cm_matrix *M, void **out);

/* ---- parity hooks (used by tests; not needed by applications) ---------------------- */
/* After a cm_commit() with cm_debug_keep_wire(M,1): the stream this rank SENT to `dst`,
 * expanded from the compact wire format into the reference's per-element (inds, data)
 * stream (XBuff, :128-133) in the reference's order (:605-613). */
int cm_debug_keep_wire(cm_matrix *M, int on);
/* lowers the number of updates a rank may stage before it flushes by itself (and the number of segments
 * one scatter-add addresses) from 2^26 - 1, so that tests reach that path with small inputs */
int cm_debug_set_staging_limit(cm_matrix *M, long max_updates);
int cm_debug_wire_count(cm_matrix *M, int dst, long *n_out);
int cm_debug_wire_expand(cm_matrix *M, int dst, long *inds_out, void *data_out);

/* ---- pure integer helpers (no GPU needed) ------------------------------------------ */
/* roc(), :379-395 */
long cm_numroc(long m, long np, long ip, long mb);
/* owner rank and local linear index of global (i,j), :586-588 */
void cm_owner_of(long i, long j, long nprow, long npcol, long mb, long nb, long lld,
                 int *rank, long *ij);
void cm_descinit_raw(int ictxt, long m, long n, long mb, long nb, long lld, int desc[9]);

#ifdef __cplusplus
}
#endif
#endif /* CM_B200_H */
