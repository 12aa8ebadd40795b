/* oracle/cm_oracle.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C, single-threaded restatement of the reference's assembly hot path
 * (swfrench/convergent-matrix, include/convergent_matrix.hpp). Each function cites the
 * reference lines it follows. All P ranks of the NPROW x NPCOL grid are simulated inside
 * one object; "messages" are applied in a fixed order (by source rank, then destination),
 * which is one of the orders the reference may legally produce.
 *
 * Parity pinning: the reference ships NO golden vectors (its tests are randomised and
 * self-consistent, test/cm_test.cpp). This restatement is pinned instead against the
 * reference header ITSELF, compiled unmodified over oracle/upcxx/upcxx.hpp
 * (oracle/_ref/libcm_ref.so): tests/test_oracle_vs_ref.py compares both on seeded inputs
 * in this container, and tests/golden/ holds fixtures generated from the reference build
 * by tests/golden/make_golden.py so the comparison travels to boxes without
 * /root/reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference leg may use it.
 */
#ifndef CM_ORACLE_H
#define CM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* dtype codes (same as include/cm_b200.h) */
#define CMO_F64 0
#define CMO_F32 1
#define CMO_I32 2

typedef struct cmo_matrix cmo_matrix;

/* reference :379-395 — ScaLAPACK numroc with source process 0 */
long cmo_roc(long m, long np, long ip, long mb);

/* reference :586-588 / :606-610 / :624-626 — owner rank and local linear index */
void cmo_owner(long i, long j, long nprow, long npcol, long mb, long nb, long lld, int *rank,
               long *ij);

/* reference :405-448. Grid coordinates are ROW-MAJOR (prow = rank / NPCOL), which is what
 * the owner formula implies; the reference ctor's `rank / NPROW` (:409) is a defect for
 * NPROW != NPCOL that does not affect update/commit results. */
cmo_matrix *cmo_create(int dtype, long m, long n, long nprow, long npcol, long mb, long nb,
                       long lld);
void cmo_destroy(cmo_matrix *M);

int cmo_nranks(const cmo_matrix *M);
long cmo_m_local(const cmo_matrix *M, int rank);
long cmo_n_local(const cmo_matrix *M, int rank);
const void *cmo_local_data(const cmo_matrix *M, int rank); /* lld * n_local elements */

/* flush threshold in elements, strict '>' (reference :74, :346) */
void cmo_set_flush_threshold(cmo_matrix *M, long thresh);

/* reference :604-615 — general slice update issued by `src`. (m_u, n_u) are view dims;
 * element (i,j) = trans ? data[j + i*ld] : data[i + ld*j] (include/local_matrix.hpp:71) */
void cmo_update_slice(cmo_matrix *M, int src, const void *data, long m_u, long n_u, long ld,
                      int trans, const long *ix, const long *jx);
/* reference :641-680 — symmetric update: only elements with ix[i] <= ix[j] */
void cmo_update_sym(cmo_matrix *M, int src, const void *data, long n_u, long ld, int trans,
                    const long *ix);
/* reference :623-629 */
void cmo_update_elem(cmo_matrix *M, int src, const void *val, long i, long j);
/* reference :694-712, executed for every rank in rank order */
void cmo_fill_lower(cmo_matrix *M);
/* reference :723-742 — drain every bin of every rank and apply (:178-187) */
void cmo_commit(cmo_matrix *M);
/* reference :491-501 */
void cmo_reset(cmo_matrix *M);
/* reference :585-596 */
void cmo_get(const cmo_matrix *M, long i, long j, void *out);

/* Wire capture: the (inds, data) streams the reference would rput from src to dst
 * (:168-171), concatenated over flushes since the last cmo_wire_clear(). */
void cmo_wire_enable(cmo_matrix *M, int on);
long cmo_wire_count(const cmo_matrix *M, int src, int dst);
void cmo_wire_copy(const cmo_matrix *M, int src, int dst, long *inds_out, void *data_out);
void cmo_wire_clear(cmo_matrix *M);

/* total number of update elements applied so far, and messages (flushes) sent */
long cmo_elements_applied(const cmo_matrix *M);
long cmo_messages_sent(const cmo_matrix *M);

#ifdef __cplusplus
}
#endif
#endif
