// include/convergent_matrix.hpp — cm::ConvergentMatrix<T,NPROW,NPCOL,MB,NB,LLD> over
// libcm_b200.so: the B200 drop-in for the assembly hot path of the reference class of the
// same name (swfrench/convergent-matrix include/convergent_matrix.hpp:298-971).
//
// Same template parameters, same member names and argument meaning; the bodies forward to
// the C ABI in include/cm_b200.h. Contract violations abort after printing the library's
// message, which is what the reference's asserts do (:417-428, :364-369); there are no
// exceptions and no return codes, as in the reference.
//
// What changed behind the API:
//   * update() stages the slice in HBM (values + indices are copied before it returns, as at
//     reference :611); owner mapping, per-owner packing and the scatter-add run in CUDA
//     kernels; there is no progress engine, so progress_interval() is stored but unused.
//   * With several ranks the exchange happens inside commit() (the pack kernel stores into the
//     owners' arenas over NVLink, or a grouped ncclSend/ncclRecv all-to-all-v) — and, as in the
//     reference (:346, :614), inside update() once more than bin_flush_threshold() elements are
//     staged: the rank then packs what it has staged and stores it into the owners' flush regions
//     without any collective call; the owners apply it inside their next commit().
//   * Process-grid coordinates are row-major (pgrid_row() = rank / NPCOL), which is what the
//     reference's owner formula implies; its ctor's rank / NPROW (:409) is a defect for
//     NPROW != NPCOL.
//   * get_local_data() returns a HOST mirror of the HBM-resident block (refreshed lazily);
//     get_local_data_device() is the PBLAS-layout block in HBM itself.
//   * save()/load() write / read the reference's on-disk format (:772-969) with pwrite/pread
//     instead of MPI-IO; they are always available (no ENABLE_MPIIO_SUPPORT switch).
#pragma once

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <thread>
#include <vector>

#include "cm_b200.h"
#include "local_matrix.hpp"

#ifndef DEFAULT_BIN_FLUSH_THRESHOLD
#define DEFAULT_BIN_FLUSH_THRESHOLD CM_DEFAULT_FLUSH_THRESHOLD  // elements (reference: 10000, :73-75)
#endif
#ifndef DEFAULT_PROGRESS_INTERVAL
#define DEFAULT_PROGRESS_INTERVAL 1  // reference :78-80; kept for API compatibility
#endif

namespace cm {

namespace internal {
inline void check(int rc, const char *what) {
  if (rc != CM_OK) {
    std::fprintf(stderr, "cm_b200: %s failed (%d): %s @ rank %d\n", what, rc, cm_last_error(), cm_rt_rank_me());
    std::abort();  // the reference asserts
  }
}
template <typename T>
struct dtype_of;
template <>
struct dtype_of<double> {
  static const int value = CM_F64;
};
template <>
struct dtype_of<float> {
  static const int value = CM_F32;
};
template <>
struct dtype_of<int> {
  static const int value = CM_I32;
};
}  // namespace internal

// ---- runtime: the analogue of upcxx::init / finalize / rank_me / rank_n / barrier ----------
// One process per GPU started by torchrun (RANK / WORLD_SIZE / LOCAL_RANK); a single process
// without those variables is rank 0 of 1 on device 0.
inline void init() { internal::check(cm_rt_init_from_env(), "cm::init"); }
inline void finalize() { internal::check(cm_rt_finalize(), "cm::finalize"); }
inline int rank_me() { return cm_rt_rank_me(); }
inline int rank_n() { return cm_rt_rank_n(); }
inline void barrier() { internal::check(cm_rt_barrier(), "cm::barrier"); }

// Threads-as-ranks on ONE GPU (the test transport of cm_b200.h): runs fn(rank) on `nranks`
// threads, like `upcxx-run -n nranks` over the smp conduit. Initialises and finalises the runtime.
inline void run_spmd_loopback(int nranks, int device, const std::function<void(int)> &fn) {
  internal::check(cm_rt_init_loopback(nranks, device), "cm_rt_init_loopback");
  std::vector<std::thread> ranks;
  for (int r = 0; r < nranks; ++r)
    ranks.emplace_back([r, &fn] {
      internal::check(cm_rt_bind_rank(r), "cm_rt_bind_rank");
      fn(r);
    });
  for (auto &t : ranks) t.join();
  internal::check(cm_rt_finalize(), "cm_rt_finalize");
}

template <typename T,              // matrix type: double, float or int
          long NPROW, long NPCOL,  // pblas process grid
          long MB, long NB,        // pblas local blocking factors
          long LLD>                // pblas local leading dim
class ConvergentMatrix {
 public:
  // collective (reference :405-448)
  ConvergentMatrix(long m, long n) {
    internal::check(cm_create(&h_, internal::dtype_of<T>::value, m, n, NPROW, NPCOL, MB, NB, LLD), "ConvergentMatrix");
    internal::check(cm_set_flush_threshold(h_, DEFAULT_BIN_FLUSH_THRESHOLD), "bin_flush_threshold");
  }
  ConvergentMatrix(const ConvergentMatrix &) = delete;
  ConvergentMatrix &operator=(const ConvergentMatrix &) = delete;
  // collective; performs a final commit (reference :457-460)
  ~ConvergentMatrix() { internal::check(cm_destroy(h_), "~ConvergentMatrix"); }

  // the live LLD x n_local column-major block, PBLAS-compatible (reference :469). Host mirror of
  // the HBM block, valid until the next update/commit/reset or destruction.
  const T *get_local_data() const {
    const void *p = nullptr;
    internal::check(cm_local_data_host(h_, &p), "get_local_data");
    return static_cast<const T *>(p);
  }
  // the same block where it lives, in HBM (for cuSOLVERMp-style consumers)
  const T *get_local_data_device() const {
    const void *p = nullptr;
    internal::check(cm_local_data_device(h_, &p), "get_local_data_device");
    return static_cast<const T *>(p);
  }
  // caller owns the copy, delete[] it (reference :478-483)
  T *get_local_data_copy() const {
    T *copy = new T[LLD * n_local()];
    internal::check(cm_local_data_copy(h_, copy), "get_local_data_copy");
    return copy;
  }

  void reset() { internal::check(cm_reset(h_), "reset"); }  // collective (reference :491-501)

  int bin_flush_threshold() const { return (int)cm_get_flush_threshold(h_); }  // reference :507
  void bin_flush_threshold(int thresh) { internal::check(cm_set_flush_threshold(h_, thresh), "bin_flush_threshold"); }
  int progress_interval() const { return cm_get_progress_interval(h_); }  // reference :529
  void progress_interval(int interval) { internal::check(cm_set_progress_interval(h_, interval), "progress_interval"); }

  long m() const { return cm_m(h_); }  // reference :550-576
  long n() const { return cm_n(h_); }
  long pgrid_row() const { return cm_pgrid_row(h_); }
  long pgrid_col() const { return cm_pgrid_col(h_); }
  long m_local() const { return cm_m_local(h_); }
  long n_local() const { return cm_n_local(h_); }

  // remote random access, read only (reference :585-596)
  T operator()(long ix, long jx) {
    T v;
    internal::check(cm_get(h_, ix, jx, &v), "operator()");
    return v;
  }

  // general slice update (reference :604-615)
  void update(LocalMatrix<T> *Mat, long *ix, long *jx) {
    internal::check(cm_update_slice(h_, Mat->data(), Mat->m(), Mat->n(), Mat->ld(), Mat->transposed() ? 1 : 0, ix, jx),
                    "update(Mat, ix, jx)");
  }
  // elemental update (reference :623-629)
  void update(T elem, long ix, long jx) { internal::check(cm_update_elem(h_, &elem, ix, jx), "update(elem, ix, jx)"); }
  // symmetric update: upper triangle of Mat only (reference :641-680)
  void update(LocalMatrix<T> *Mat, long *ix) {
#ifndef NOCHECK
    if (Mat->m() != Mat->n()) internal::check(CM_ERR_INVALID, "update(Mat, ix): Mat must be square");
#endif
    internal::check(cm_update_sym(h_, Mat->data(), Mat->n(), Mat->ld(), Mat->transposed() ? 1 : 0, ix),
                    "update(Mat, ix)");
  }
  // device-resident slice: payload and index arrays already in this rank's HBM (borrowed until
  // the next flush()/commit() returns)
  void update(const DeviceLocalMatrix<T> *Mat, const long *d_ix, const long *d_jx) {
    internal::check(cm_update_slice_device(h_, Mat->data(), Mat->m(), Mat->n(), Mat->ld(), Mat->transposed() ? 1 : 0,
                                           d_ix, d_jx, CM_MASK_NONE),
                    "update(DeviceLocalMatrix, d_ix, d_jx)");
  }

  void fill_lower() { internal::check(cm_fill_lower(h_), "fill_lower"); }  // reference :694-712
  void flush() { internal::check(cm_flush(h_), "flush"); }                 // local, non-collective
  void commit() { internal::check(cm_commit(h_), "commit"); }              // collective (reference :723-742)

  // map eq(value, i, j) over the locally stored elements (reference :756-768; test utility)
  bool verify_local_elements(std::function<bool(const T, long, long)> eq) {
    const T *local = get_local_data();
    const long myrow = pgrid_row(), mycol = pgrid_col();
    for (long j = 0; j < n(); j++)
      if ((j / NB) % NPCOL == mycol) {
        const long off_j = LLD * ((j / (NB * NPCOL)) * NB + j % NB);
        for (long i = 0; i < m(); i++)
          if ((i / MB) % NPROW == myrow) {
            const long ij = off_j + (i / (MB * NPROW)) * MB + i % MB;
            if (!eq(local[ij], i, j)) return false;
          }
      }
    return true;
  }

  // dense column-major global matrix on disk, the file format of the reference's MPI-IO save/load
  // (:788-873, :888-967); collective; commit() first, as in the reference. No MPI needed.
  void save(const char *fname) { internal::check(cm_save(h_, fname), "save"); }
  void load(const char *fname) { internal::check(cm_load(h_, fname), "load"); }

  // PBLAS array descriptor {1, ictxt, M, N, MB, NB, 0, 0, LLD} of the local block
  void descinit(int ictxt, int desc[9]) const { internal::check(cm_descinit(h_, ictxt, desc), "descinit"); }

  // knobs without a reference counterpart
  // extension: deferred scatter-add (cm_b200.h, cm_set_apply_policy); both calls are collective
  void apply_policy(int policy, long budget_bytes = 0) {
    internal::check(cm_set_apply_policy(h_, policy, budget_bytes), "apply_policy");
  }
  void materialize() { internal::check(cm_materialize(h_), "materialize"); }
  void deterministic(bool on) { internal::check(cm_set_deterministic(h_, on ? 1 : 0), "deterministic"); }
  cm_matrix *handle() const { return h_; }

 private:
  cm_matrix *h_ = nullptr;
};

}  // namespace cm
