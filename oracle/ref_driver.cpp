// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY.
//
// C-ABI driver around the UNMODIFIED reference header
// /root/reference/include/convergent_matrix.hpp compiled over the in-repo UPC++ shim
// (oracle/upcxx/upcxx.hpp). Built by oracle/Makefile into oracle/_ref/libcm_ref.so.
// One std::thread per rank (threads-as-ranks); every call below is executed on the
// rank's own thread so the reference's persona assumptions hold.
//
// Used by: tests/ (parity oracle, golden-vector generation in this container) and
// bench.py's cpu_baseline / `--impl reference` leg ("reference header over in-repo
// SMP shim"). Never by the product path.
#include <sched.h>

#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <upcxx/upcxx.hpp>

#include "include/convergent_matrix.hpp"  // the reference, from -I /root/reference
#include "include/local_matrix.hpp"

namespace {

// type-erased view of one rank's reference ConvergentMatrix instance
struct IRefMatrix {
  virtual ~IRefMatrix() {}
  virtual void update_slice(const void *data, long m_u, long n_u, long ld, int trans, long *ix,
                            long *jx) = 0;
  virtual void update_sym(const void *data, long n_u, long ld, int trans, long *ix) = 0;
  virtual void update_elem(const void *val, long i, long j) = 0;
  virtual void fill_lower() = 0;
  virtual void commit() = 0;
  virtual void reset() = 0;
  virtual const void *local_data() = 0;
  virtual long m_local() = 0;
  virtual long n_local() = 0;
  virtual long pgrid_row() = 0;
  virtual long pgrid_col() = 0;
  virtual void get(long i, long j, void *out) = 0;
  virtual void set_threshold(int t) = 0;
};

template <typename T, long NPROW, long NPCOL, long MB, long NB, long LLD>
struct RefMatrix final : IRefMatrix {
  cm::ConvergentMatrix<T, NPROW, NPCOL, MB, NB, LLD> mat;
  RefMatrix(long m, long n) : mat(m, n) {}

  // (m_u, n_u) are VIEW dims; element (i,j) = trans ? data[j + i*ld] : data[i + ld*j]
  // which is exactly LocalMatrix::operator() (reference include/local_matrix.hpp:71).
  void update_slice(const void *data, long m_u, long n_u, long ld, int trans, long *ix,
                    long *jx) override {
    T *d = const_cast<T *>(static_cast<const T *>(data));
    if (!trans) {
      cm::LocalMatrix<T> M(m_u, n_u, ld, d);
      mat.update(&M, ix, jx);
    } else {
      cm::LocalMatrix<T> base(n_u, m_u, ld, d);
      cm::LocalMatrix<T> *view = base.trans();
      mat.update(view, ix, jx);
      delete view;
    }
  }
  void update_sym(const void *data, long n_u, long ld, int trans, long *ix) override {
    T *d = const_cast<T *>(static_cast<const T *>(data));
    if (!trans) {
      cm::LocalMatrix<T> M(n_u, n_u, ld, d);
      mat.update(&M, ix);
    } else {
      cm::LocalMatrix<T> base(n_u, n_u, ld, d);
      cm::LocalMatrix<T> *view = base.trans();
      mat.update(view, ix);
      delete view;
    }
  }
  void update_elem(const void *val, long i, long j) override {
    mat.update(*static_cast<const T *>(val), i, j);
  }
  void fill_lower() override { mat.fill_lower(); }
  void commit() override { mat.commit(); }
  void reset() override { mat.reset(); }
  const void *local_data() override { return mat.get_local_data(); }
  long m_local() override { return mat.m_local(); }
  long n_local() override { return mat.n_local(); }
  long pgrid_row() override { return mat.pgrid_row(); }
  long pgrid_col() override { return mat.pgrid_col(); }
  void get(long i, long j, void *out) override { *static_cast<T *>(out) = mat(i, j); }
  void set_threshold(int t) override { mat.bin_flush_threshold(t); }
};

IRefMatrix *make_matrix(int dtype, long nprow, long npcol, long mb, long nb, long lld, long m,
                        long n) {
#define X(T, CODE, NPROW, NPCOL, MB, NB, LLD)                                              \
  if (dtype == CODE && nprow == NPROW && npcol == NPCOL && mb == MB && nb == NB &&         \
      lld == LLD)                                                                          \
    return new RefMatrix<T, NPROW, NPCOL, MB, NB, LLD>(m, n);
#include "ref_configs.def"
#undef X
  return nullptr;
}

bool have_config(int dtype, long nprow, long npcol, long mb, long nb, long lld) {
#define X(T, CODE, NPROW, NPCOL, MB, NB, LLD)                                              \
  if (dtype == CODE && nprow == NPROW && npcol == NPCOL && mb == MB && nb == NB &&         \
      lld == LLD)                                                                          \
    return true;
#include "ref_configs.def"
#undef X
  return false;
}

size_t dtype_size(int dtype) { return dtype == 0 ? 8 : 4; }

struct WireStream {
  std::vector<long> inds;
  std::vector<unsigned char> data;
};

struct Session {
  int P = 0;
  int dtype = 0;
  upcxx::shim::World *world = nullptr;
  std::vector<std::thread> threads;
  std::vector<IRefMatrix *> mats;
  std::vector<WireStream> wire;  // [src * P + dst]

  // per-rank task slots
  std::mutex mu;
  std::condition_variable cv_task, cv_done;
  std::vector<std::function<void()>> task;
  std::vector<bool> has_task;
  int pending = 0;
  bool stop = false;

  void worker(int rank) {
    upcxx::shim::enter(world, rank);
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_task.wait(lk, [&] { return stop || has_task[rank]; });
        if (!has_task[rank] && stop) return;
        f.swap(task[rank]);
        has_task[rank] = false;
      }
      f();
      {
        std::lock_guard<std::mutex> lk(mu);
        --pending;
      }
      cv_done.notify_all();
    }
  }

  // run fn(rank) on every rank's thread concurrently; returns when all are done
  void run_all(const std::function<void(int)> &fn) {
    {
      std::lock_guard<std::mutex> lk(mu);
      for (int r = 0; r < P; ++r) {
        task[r] = [fn, r] { fn(r); };
        has_task[r] = true;
      }
      pending = P;
    }
    cv_task.notify_all();
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { return pending == 0; });
  }

  void run_on(int rank, const std::function<void()> &fn) {
    {
      std::lock_guard<std::mutex> lk(mu);
      task[rank] = fn;
      has_task[rank] = true;
      pending = 1;
    }
    cv_task.notify_all();
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { return pending == 0; });
  }
};

Session *g_active = nullptr;  // the shim world is process-global: one session at a time

void pin_to_core(int core) {
  cpu_set_t set;
  CPU_ZERO(&set);
  CPU_SET(core, &set);
  sched_setaffinity(0, sizeof(set), &set);
}

}  // namespace

extern "C" {

int cmref_have_config(int dtype, long nprow, long npcol, long mb, long nb, long lld) {
  return have_config(dtype, nprow, npcol, mb, nb, lld) ? 1 : 0;
}

void *cmref_open(int dtype, long nprow, long npcol, long mb, long nb, long lld, long m, long n,
                 int record_wire) {
  if (g_active != nullptr) return nullptr;
  if (!have_config(dtype, nprow, npcol, mb, nb, lld)) return nullptr;
  Session *s = new Session();
  s->P = (int)(nprow * npcol);
  s->dtype = dtype;
  s->world = new upcxx::shim::World(s->P);
  s->mats.assign(s->P, nullptr);
  s->task.resize(s->P);
  s->has_task.assign(s->P, false);
  s->wire.resize((size_t)s->P * s->P);
  if (record_wire) {
    const int P = s->P;
    s->world->on_rput = [s, P](int src, int dst, const void *p, size_t n, size_t esz,
                               bool is_index) {
      WireStream &w = s->wire[(size_t)src * P + dst];
      if (is_index) {
        const long *q = static_cast<const long *>(p);
        w.inds.insert(w.inds.end(), q, q + n);
      } else {
        const unsigned char *q = static_cast<const unsigned char *>(p);
        w.data.insert(w.data.end(), q, q + n * esz);
      }
    };
  }
  for (int r = 0; r < s->P; ++r) s->threads.emplace_back([s, r] { s->worker(r); });
  // collective construction (reference :405-448)
  s->run_all([=](int r) { s->mats[r] = make_matrix(dtype, nprow, npcol, mb, nb, lld, m, n); });
  g_active = s;
  return s;
}

void cmref_close(void *h) {
  Session *s = static_cast<Session *>(h);
  if (s == nullptr) return;
  // collective destruction: the reference dtor commits (:457-460)
  s->run_all([s](int r) {
    delete s->mats[r];
    s->mats[r] = nullptr;
  });
  {
    std::lock_guard<std::mutex> lk(s->mu);
    s->stop = true;
  }
  s->cv_task.notify_all();
  for (auto &t : s->threads) t.join();
  delete s->world;
  if (g_active == s) g_active = nullptr;
  delete s;
}

int cmref_nranks(void *h) { return static_cast<Session *>(h)->P; }

void cmref_update_slice(void *h, int rank, const void *data, long m_u, long n_u, long ld,
                        int trans, const long *ix, const long *jx) {
  Session *s = static_cast<Session *>(h);
  s->run_on(rank, [=] {
    s->mats[rank]->update_slice(data, m_u, n_u, ld, trans, const_cast<long *>(ix),
                                const_cast<long *>(jx));
  });
}

void cmref_update_sym(void *h, int rank, const void *data, long n_u, long ld, int trans,
                      const long *ix) {
  Session *s = static_cast<Session *>(h);
  s->run_on(rank,
            [=] { s->mats[rank]->update_sym(data, n_u, ld, trans, const_cast<long *>(ix)); });
}

void cmref_update_elem(void *h, int rank, const void *val, long i, long j) {
  Session *s = static_cast<Session *>(h);
  s->run_on(rank, [=] { s->mats[rank]->update_elem(val, i, j); });
}

// n element updates issued back to back on `rank` (vals: n elements of the matrix type)
void cmref_update_elems(void *h, int rank, const void *vals, const long *is, const long *js,
                        long n) {
  Session *s = static_cast<Session *>(h);
  const size_t esz = dtype_size(s->dtype);
  s->run_on(rank, [=] {
    const unsigned char *v = static_cast<const unsigned char *>(vals);
    for (long k = 0; k < n; ++k) s->mats[rank]->update_elem(v + k * esz, is[k], js[k]);
  });
}

void cmref_fill_lower(void *h) {
  Session *s = static_cast<Session *>(h);
  s->run_all([s](int r) { s->mats[r]->fill_lower(); });
}

void cmref_commit(void *h) {
  Session *s = static_cast<Session *>(h);
  s->run_all([s](int r) { s->mats[r]->commit(); });
}

void cmref_reset(void *h) {
  Session *s = static_cast<Session *>(h);
  s->run_all([s](int r) { s->mats[r]->reset(); });
}

const void *cmref_local_data(void *h, int rank) {
  return static_cast<Session *>(h)->mats[rank]->local_data();
}
long cmref_m_local(void *h, int rank) { return static_cast<Session *>(h)->mats[rank]->m_local(); }
long cmref_n_local(void *h, int rank) { return static_cast<Session *>(h)->mats[rank]->n_local(); }
long cmref_pgrid_row(void *h, int rank) {
  return static_cast<Session *>(h)->mats[rank]->pgrid_row();
}
long cmref_pgrid_col(void *h, int rank) {
  return static_cast<Session *>(h)->mats[rank]->pgrid_col();
}

void cmref_get(void *h, int rank, long i, long j, void *out) {
  Session *s = static_cast<Session *>(h);
  s->run_on(rank, [=] { s->mats[rank]->get(i, j, out); });
}

void cmref_set_flush_threshold(void *h, int thresh) {
  Session *s = static_cast<Session *>(h);
  s->run_all([s, thresh](int r) { s->mats[r]->set_threshold(thresh); });
}

long cmref_wire_count(void *h, int src, int dst) {
  Session *s = static_cast<Session *>(h);
  return (long)s->wire[(size_t)src * s->P + dst].inds.size();
}

void cmref_wire_copy(void *h, int src, int dst, long *inds_out, void *data_out) {
  Session *s = static_cast<Session *>(h);
  const WireStream &w = s->wire[(size_t)src * s->P + dst];
  if (!w.inds.empty()) std::memcpy(inds_out, w.inds.data(), w.inds.size() * sizeof(long));
  if (!w.data.empty()) std::memcpy(data_out, w.data.data(), w.data.size());
}

void cmref_wire_clear(void *h) {
  Session *s = static_cast<Session *>(h);
  for (auto &w : s->wire) {
    w.inds.clear();
    w.data.clear();
  }
}

// CPU baseline: rank r issues nupd slice updates (each m_u x n_u, ld = m_u, payloads
// contiguous in data[r], index sets contiguous in ix[r] / jx[r]) with a commit() every
// `commit_every` updates (<=0: only at the end) and a final commit(). Timer boundaries
// follow SURVEY.md §8(d): first update() -> return of the last commit(), max over ranks,
// all ranks released together by a barrier. If pin != 0, rank r is pinned to core
// (r % ncores). out_secs[0] = max total, out_secs[1] = max time inside update() calls
// (binning + threshold flushes), out_secs[2] = max time inside commit().
// data_period > 0: update u reads payload (u % data_period) — the index sets are all distinct, the
// payload buffers of a bounded sample are cycled (payload values do not influence the time).
void cmref_bench_slices(void *h, const void *const *data, const long *const *ix,
                        const long *const *jx, long nupd, long m_u, long n_u,
                        long commit_every, int pin, double *out_secs, long data_period) {
  Session *s = static_cast<Session *>(h);
  const size_t esz = dtype_size(s->dtype);
  std::vector<double> t_total(s->P), t_upd(s->P), t_commit(s->P);
  const unsigned ncores = std::max(1u, std::thread::hardware_concurrency());
  s->run_all([&](int r) {
    using clk = std::chrono::steady_clock;
    if (pin) pin_to_core(r % ncores);
    IRefMatrix *M = s->mats[r];
    const unsigned char *d = static_cast<const unsigned char *>(data[r]);
    upcxx::barrier();
    const auto t0 = clk::now();
    double upd = 0, com = 0;
    for (long u = 0; u < nupd; ++u) {
      const auto a = clk::now();
      const long du = data_period > 0 ? u % data_period : u;
      M->update_slice(d + (size_t)du * m_u * n_u * esz, m_u, n_u, m_u, 0,
                      const_cast<long *>(ix[r] + u * m_u), const_cast<long *>(jx[r] + u * n_u));
      const auto b = clk::now();
      upd += std::chrono::duration<double>(b - a).count();
      if (commit_every > 0 && (u + 1) % commit_every == 0 && u + 1 < nupd) {
        M->commit();
        com += std::chrono::duration<double>(clk::now() - b).count();
      }
    }
    const auto c = clk::now();
    M->commit();
    const auto t1 = clk::now();
    com += std::chrono::duration<double>(t1 - c).count();
    t_total[r] = std::chrono::duration<double>(t1 - t0).count();
    t_upd[r] = upd;
    t_commit[r] = com;
  });
  out_secs[0] = out_secs[1] = out_secs[2] = 0;
  for (int r = 0; r < s->P; ++r) {
    out_secs[0] = std::max(out_secs[0], t_total[r]);
    out_secs[1] = std::max(out_secs[1], t_upd[r]);
    out_secs[2] = std::max(out_secs[2], t_commit[r]);
  }
}

}  // extern "C"
