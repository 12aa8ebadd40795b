/* oracle/cm_oracle.c — TEST INFRASTRUCTURE ONLY. See cm_oracle.h for scope and pinning.
 * Citations are into /root/reference/include/convergent_matrix.hpp unless noted. */
#include "cm_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef struct {
  long n, cap;
  long *inds;
  unsigned char *data;
} cmo_bin; /* reference Bin<T>, :192-246 */

typedef struct {
  long n, cap;
  long *inds;
  unsigned char *data;
} cmo_wire;

struct cmo_matrix {
  int dtype;
  size_t esz;
  long m, n, nprow, npcol, mb, nb, lld;
  int P;
  long thresh;
  long *m_local, *n_local;
  unsigned char **local; /* [P] lld * n_local elements, zero-initialised (:431-432) */
  cmo_bin *bins;         /* [src * P + dst] */
  cmo_wire *wire;        /* [src * P + dst] */
  int wire_on;
  long elements_applied, messages_sent;
};

/* :379-395 */
long cmo_roc(long m, long np, long ip, long mb) {
  long nblocks = m / mb;
  long m_local = (nblocks / np) * mb;
  if (nblocks % np > ip)
    m_local += mb;
  else if (nblocks % np == ip)
    m_local += m % mb;
  return m_local;
}

/* :586-588 (== :606-610, :624-626) */
void cmo_owner(long i, long j, long nprow, long npcol, long mb, long nb, long lld, int *rank,
               long *ij) {
  *rank = (int)((j / nb) % npcol + npcol * ((i / mb) % nprow));
  *ij = lld * ((j / (nb * npcol)) * nb + j % nb) + (i / (mb * nprow)) * mb + i % mb;
}

cmo_matrix *cmo_create(int dtype, long m, long n, long nprow, long npcol, long mb, long nb,
                       long lld) {
  cmo_matrix *M = (cmo_matrix *)calloc(1, sizeof(cmo_matrix));
  int r;
  M->dtype = dtype;
  M->esz = dtype == CMO_F64 ? 8 : 4;
  M->m = m;
  M->n = n;
  M->nprow = nprow;
  M->npcol = npcol;
  M->mb = mb;
  M->nb = nb;
  M->lld = lld;
  M->P = (int)(nprow * npcol);
  M->thresh = 10000; /* DEFAULT_BIN_FLUSH_THRESHOLD, :73-75 */
  M->m_local = (long *)calloc(M->P, sizeof(long));
  M->n_local = (long *)calloc(M->P, sizeof(long));
  M->local = (unsigned char **)calloc(M->P, sizeof(unsigned char *));
  M->bins = (cmo_bin *)calloc((size_t)M->P * M->P, sizeof(cmo_bin));
  M->wire = (cmo_wire *)calloc((size_t)M->P * M->P, sizeof(cmo_wire));
  for (r = 0; r < M->P; ++r) {
    long myrow = r / npcol, mycol = r % npcol; /* row-major grid (see header) */
    M->m_local[r] = cmo_roc(m, nprow, myrow, mb); /* :412 */
    M->n_local[r] = cmo_roc(n, npcol, mycol, nb); /* :413 */
    /* the reference asserts these (:417-428); a checker must not scribble instead */
    if (m <= 0 || n <= 0 || M->m_local[r] <= 0 || M->n_local[r] <= 0 || M->m_local[r] > lld) {
      cmo_destroy(M);
      return NULL;
    }
    M->local[r] = (unsigned char *)calloc((size_t)(lld * M->n_local[r]) + 1, M->esz); /* :415,:432 */
  }
  return M;
}

void cmo_destroy(cmo_matrix *M) {
  int k;
  if (!M) return;
  for (k = 0; k < M->P; ++k) free(M->local[k]);
  for (k = 0; k < M->P * M->P; ++k) {
    free(M->bins[k].inds);
    free(M->bins[k].data);
    free(M->wire[k].inds);
    free(M->wire[k].data);
  }
  free(M->local);
  free(M->bins);
  free(M->wire);
  free(M->m_local);
  free(M->n_local);
  free(M);
}

int cmo_nranks(const cmo_matrix *M) { return M->P; }
long cmo_m_local(const cmo_matrix *M, int rank) { return M->m_local[rank]; }
long cmo_n_local(const cmo_matrix *M, int rank) { return M->n_local[rank]; }
const void *cmo_local_data(const cmo_matrix *M, int rank) { return M->local[rank]; }
void cmo_set_flush_threshold(cmo_matrix *M, long thresh) { M->thresh = thresh; }
long cmo_elements_applied(const cmo_matrix *M) { return M->elements_applied; }
long cmo_messages_sent(const cmo_matrix *M) { return M->messages_sent; }

/* Bin::append, :205-208 */
static void bin_append(cmo_matrix *M, cmo_bin *b, const void *val, long ij) {
  if (b->n == b->cap) {
    b->cap = b->cap ? 2 * b->cap : 1024;
    b->inds = (long *)realloc(b->inds, (size_t)b->cap * sizeof(long));
    b->data = (unsigned char *)realloc(b->data, (size_t)b->cap * M->esz);
  }
  b->inds[b->n] = ij;
  memcpy(b->data + (size_t)b->n * M->esz, val, M->esz);
  b->n++;
}

/* apply<T>, :178-187: elts[inds[i]] += data[i] */
static void apply(cmo_matrix *M, int dst, const cmo_bin *b) {
  long k;
  unsigned char *base = M->local[dst];
  switch (M->dtype) {
    case CMO_F64: {
      double *e = (double *)base;
      const double *d = (const double *)b->data;
      for (k = 0; k < b->n; ++k) e[b->inds[k]] += d[k];
      break;
    }
    case CMO_F32: {
      float *e = (float *)base;
      const float *d = (const float *)b->data;
      for (k = 0; k < b->n; ++k) e[b->inds[k]] += d[k];
      break;
    }
    default: {
      int *e = (int *)base;
      const int *d = (const int *)b->data;
      for (k = 0; k < b->n; ++k) e[b->inds[k]] += d[k];
      break;
    }
  }
  M->elements_applied += b->n;
}

/* flush(thresh), :330-375: ship every bin of `src` with size() > thresh (strict) */
static void flush(cmo_matrix *M, int src, long thresh) {
  int dst;
  for (dst = 0; dst < M->P; ++dst) {
    cmo_bin *b = &M->bins[(size_t)src * M->P + dst];
    if (b->n > thresh) {
      if (M->wire_on) {
        cmo_wire *w = &M->wire[(size_t)src * M->P + dst];
        if (w->n + b->n > w->cap) {
          w->cap = 2 * (w->n + b->n);
          w->inds = (long *)realloc(w->inds, (size_t)w->cap * sizeof(long));
          w->data = (unsigned char *)realloc(w->data, (size_t)w->cap * M->esz);
        }
        memcpy(w->inds + w->n, b->inds, (size_t)b->n * sizeof(long));
        memcpy(w->data + (size_t)w->n * M->esz, b->data, (size_t)b->n * M->esz);
        w->n += b->n;
      }
      apply(M, dst, b);
      b->n = 0; /* :240-241 */
      M->messages_sent++; /* :243 */
    }
  }
}

/* LocalMatrix::operator(), include/local_matrix.hpp:71 */
static const unsigned char *payload(const cmo_matrix *M, const void *data, long i, long j,
                                    long ld, int trans) {
  return (const unsigned char *)data + (size_t)(trans ? j + i * ld : i + ld * j) * M->esz;
}

/* :604-615 */
void cmo_update_slice(cmo_matrix *M, int src, const void *data, long m_u, long n_u, long ld,
                      int trans, const long *ix, const long *jx) {
  long i, j;
  for (j = 0; j < n_u; ++j) {
    int pcol = (int)((jx[j] / M->nb) % M->npcol);
    long off_j = M->lld * ((jx[j] / (M->nb * M->npcol)) * M->nb + jx[j] % M->nb);
    for (i = 0; i < m_u; ++i) {
      int rank = (int)(pcol + M->npcol * ((ix[i] / M->mb) % M->nprow));
      long ij = off_j + (ix[i] / (M->mb * M->nprow)) * M->mb + ix[i] % M->mb;
      bin_append(M, &M->bins[(size_t)src * M->P + rank], payload(M, data, i, j, ld, trans), ij);
    }
  }
  flush(M, src, M->thresh); /* :614 */
}

/* :641-680 */
void cmo_update_sym(cmo_matrix *M, int src, const void *data, long n_u, long ld, int trans,
                    const long *ix) {
  long i, j;
  for (j = 0; j < n_u; ++j) {
    int pcol = (int)((ix[j] / M->nb) % M->npcol);
    long off_j = M->lld * ((ix[j] / (M->nb * M->npcol)) * M->nb + ix[j] % M->nb);
    for (i = 0; i < n_u; ++i)
      if (ix[i] <= ix[j]) { /* :654 */
        int rank = (int)(pcol + M->npcol * ((ix[i] / M->mb) % M->nprow));
        long ij = off_j + (ix[i] / (M->mb * M->nprow)) * M->mb + ix[i] % M->mb;
        bin_append(M, &M->bins[(size_t)src * M->P + rank], payload(M, data, i, j, ld, trans), ij);
      }
  }
  flush(M, src, M->thresh); /* :670 */
}

/* :623-629 */
void cmo_update_elem(cmo_matrix *M, int src, const void *val, long i, long j) {
  int rank;
  long ij;
  cmo_owner(i, j, M->nprow, M->npcol, M->mb, M->nb, M->lld, &rank, &ij);
  bin_append(M, &M->bins[(size_t)src * M->P + rank], val, ij);
  flush(M, src, M->thresh);
}

/* :694-712, for every rank. The local block is snapshotted per rank before binning so
 * that threshold flushes landing in the same rank's lower triangle cannot be re-read
 * (they never are: sources are strictly upper, targets strictly lower). */
void cmo_fill_lower(cmo_matrix *M) {
  int r;
  for (r = 0; r < M->P; ++r) {
    long myrow = r / M->npcol, mycol = r % M->npcol;
    long i, j;
    const unsigned char *lp = M->local[r];
    for (j = 0; j < M->n_local[r]; ++j) {
      long ix = (j / M->nb) * (M->npcol * M->nb) + mycol * M->nb + j % M->nb; /* :697 */
      for (i = 0; i < M->m_local[r]; ++i) {
        long jx = (i / M->mb) * (M->nprow * M->mb) + myrow * M->mb + i % M->mb; /* :699 */
        if (ix > jx) { /* :701 — transposed global indices */
          int rank;
          long ij;
          cmo_owner(ix, jx, M->nprow, M->npcol, M->mb, M->nb, M->lld, &rank, &ij); /* :702-704 */
          bin_append(M, &M->bins[(size_t)r * M->P + rank], lp + (size_t)(i + j * M->lld) * M->esz,
                     ij); /* :705 */
        } else {
          break; /* :707 */
        }
      }
    }
    flush(M, r, M->thresh); /* :711 */
  }
}

/* :723-742 */
void cmo_commit(cmo_matrix *M) {
  int src;
  for (src = 0; src < M->P; ++src) flush(M, src, 0); /* :726; size() > 0 */
}

/* :491-501 */
void cmo_reset(cmo_matrix *M) {
  int r;
  cmo_commit(M);
  for (r = 0; r < M->P; ++r) memset(M->local[r], 0, (size_t)(M->lld * M->n_local[r]) * M->esz);
}

/* :585-596 */
void cmo_get(const cmo_matrix *M, long i, long j, void *out) {
  int rank;
  long ij;
  cmo_owner(i, j, M->nprow, M->npcol, M->mb, M->nb, M->lld, &rank, &ij);
  memcpy(out, M->local[rank] + (size_t)ij * M->esz, M->esz);
}

void cmo_wire_enable(cmo_matrix *M, int on) { M->wire_on = on; }
long cmo_wire_count(const cmo_matrix *M, int src, int dst) {
  return M->wire[(size_t)src * M->P + dst].n;
}
void cmo_wire_copy(const cmo_matrix *M, int src, int dst, long *inds_out, void *data_out) {
  const cmo_wire *w = &M->wire[(size_t)src * M->P + dst];
  if (w->n) {
    memcpy(inds_out, w->inds, (size_t)w->n * sizeof(long));
    memcpy(data_out, w->data, (size_t)w->n * M->esz);
  }
}
void cmo_wire_clear(cmo_matrix *M) {
  int k;
  for (k = 0; k < M->P * M->P; ++k) M->wire[k].n = 0;
}
