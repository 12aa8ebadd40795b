// oracle/upcxx/upcxx.hpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A threads-as-ranks stand-in for the slice of UPC++ v1.0 that the reference
// header /root/reference/include/convergent_matrix.hpp uses,
// so that the reference's own update()/commit() code can be compiled UNMODIFIED
// here (`-I oracle -I /root/reference`) and used as the parity oracle and as the
// "reference header over in-repo SMP shim" CPU baseline.
//
// Semantics kept from UPC++ (because the reference relies on them):
//   * rpc_ff closures run on the TARGET rank's thread, inside that rank's
//     progress() (reference :242 -> :179-187), so owner-side accumulation stays
//     serialised per owner exactly as in the master-persona model.
//   * a dist_object<T>& passed as an RPC argument is re-bound to the target
//     rank's instance (reference :242 passes `state` by reference).
//   * barrier()/reduce_all()/future::wait() make user-level progress while waiting.
// Simplifications (all ranks share one address space):
//   * rpc() with a return value (only used for the remote allocation, :226-228)
//     executes immediately on the caller's thread; rput/rget are memcpy.
//   * futures are always ready; .then() runs its continuation at once.
//
// None of the hot path's arithmetic lives here: mapping (:606-610) and `+=`
// (:183) are executed by the reference header itself.
#pragma once

#include <atomic>
#include <cstddef>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

namespace upcxx {

// ----------------------------------------------------------------------------
// world / rank plumbing
// ----------------------------------------------------------------------------
namespace shim {

struct Inbox {
  std::mutex mu;
  std::deque<std::function<void()>> q;
};

// Called for every rput: (src rank, dst rank, payload, n elements, element size,
// is_index_stream). Lets tests capture the reference's exact wire content.
using RputHook = std::function<void(int, int, const void *, size_t, size_t, bool)>;

struct World {
  int n = 1;
  std::vector<std::unique_ptr<Inbox>> inbox;
  std::atomic<int> bar_count{0};
  std::atomic<long> bar_gen{0};
  std::mutex reg_mu;
  std::vector<std::vector<void *>> registry;  // [dist_object id][rank]
  std::vector<const void *> red_src;
  RputHook on_rput;

  explicit World(int nranks) : n(nranks), red_src(nranks, nullptr) {
    for (int i = 0; i < nranks; ++i) inbox.emplace_back(new Inbox());
  }
};

inline World *&world_ptr() {
  static World *w = nullptr;
  return w;
}
inline int &tl_rank() {
  static thread_local int r = 0;
  return r;
}
inline int &tl_next_dist_id() {
  static thread_local int id = 0;
  return id;
}
// Bind the calling thread to `rank` of world `w` (called by the SPMD launcher).
inline void enter(World *w, int rank) {
  world_ptr() = w;
  tl_rank() = rank;
  tl_next_dist_id() = 0;
}

}  // namespace shim

inline void init() {}
inline void finalize() {}
inline int rank_me() { return shim::tl_rank(); }
inline int rank_n() { return shim::world_ptr()->n; }

// user-level progress: run every closure other ranks injected for this rank
inline void progress() {
  shim::World *w = shim::world_ptr();
  shim::Inbox &in = *w->inbox[rank_me()];
  std::deque<std::function<void()>> todo;
  {
    std::lock_guard<std::mutex> g(in.mu);
    if (in.q.empty()) return;
    todo.swap(in.q);
  }
  for (auto &f : todo) f();
}

inline void barrier() {
  shim::World *w = shim::world_ptr();
  const long gen = w->bar_gen.load(std::memory_order_acquire);
  if (w->bar_count.fetch_add(1, std::memory_order_acq_rel) + 1 == w->n) {
    w->bar_count.store(0, std::memory_order_relaxed);
    w->bar_gen.fetch_add(1, std::memory_order_acq_rel);
  } else {
    while (w->bar_gen.load(std::memory_order_acquire) == gen) {
      progress();
      std::this_thread::yield();
    }
  }
}

// ----------------------------------------------------------------------------
// global pointers and the shared segment
// ----------------------------------------------------------------------------
template <typename T>
struct global_ptr {
  T *p = nullptr;
  int rank = 0;
  T *local() const { return p; }
  global_ptr operator+(std::ptrdiff_t d) const { return global_ptr{p + d, rank}; }
};

template <typename T>
global_ptr<T> new_array(std::size_t n) {
  T *p = new (std::nothrow) T[n];
  if (p == nullptr) throw std::bad_alloc();
  return global_ptr<T>{p, rank_me()};
}

template <typename T>
void delete_array(global_ptr<T> g) {
  delete[] g.p;
}

// ----------------------------------------------------------------------------
// futures / promises (always ready)
// ----------------------------------------------------------------------------
template <typename... T>
class future;

namespace shim {
template <typename R>
struct is_future : std::false_type {};
template <typename... U>
struct is_future<future<U...>> : std::true_type {};
}  // namespace shim

template <typename... T>
class future {
 public:
  std::tuple<T...> v;
  future() = default;
  template <typename... U,
            typename = std::enable_if_t<sizeof...(U) == sizeof...(T) && (sizeof...(U) > 0)>>
  explicit future(U &&... u) : v(std::forward<U>(u)...) {}
  explicit future(std::tuple<T...> &&t, int /*tag*/) : v(std::move(t)) {}

  // wait(): void for future<>, T for future<T>, tuple otherwise
  decltype(auto) wait() {
    progress();
    if constexpr (sizeof...(T) == 0) {
      return;
    } else if constexpr (sizeof...(T) == 1) {
      return std::get<0>(v);
    } else {
      return v;
    }
  }

  template <typename F>
  auto then(F &&f) {
    using R = std::invoke_result_t<F, T...>;
    if constexpr (std::is_void<R>::value) {
      std::apply(std::forward<F>(f), v);
      return future<>();
    } else if constexpr (shim::is_future<R>::value) {
      return std::apply(std::forward<F>(f), v);
    } else {
      return future<R>(std::apply(std::forward<F>(f), v));
    }
  }
};

inline future<> make_future() { return future<>(); }

template <typename T>
future<std::decay_t<T>> to_future(T &&x) {
  return future<std::decay_t<T>>(std::forward<T>(x));
}

template <typename... A, typename... B>
future<A..., B...> when_all(future<A...> a, future<B...> b) {
  return future<A..., B...>(std::tuple_cat(std::move(a.v), std::move(b.v)), 0);
}

template <typename... T>
class promise {
 public:
  future<T...> finalize() { return future<T...>(); }
};

namespace operation_cx {
struct as_promise_t {};
template <typename... T>
as_promise_t as_promise(promise<T...> &) {
  return as_promise_t{};
}
}  // namespace operation_cx

// ----------------------------------------------------------------------------
// one-sided transfers
// ----------------------------------------------------------------------------
template <typename T, typename Cx>
void rput(const T *src, global_ptr<T> dst, std::size_t n, Cx) {
  std::memcpy(dst.p, src, n * sizeof(T));
  shim::World *w = shim::world_ptr();
  if (w->on_rput)
    w->on_rput(rank_me(), dst.rank, src, n, sizeof(T), std::is_same<T, long>::value);
}

template <typename T>
future<T> rget(global_ptr<T> src) {
  return future<T>(*src.p);
}

// ----------------------------------------------------------------------------
// dist_object
// ----------------------------------------------------------------------------
template <typename T>
class dist_object {
  int id_;
  T value_;

 public:
  template <typename... A>
  explicit dist_object(A &&... a) : id_(shim::tl_next_dist_id()++), value_(std::forward<A>(a)...) {
    shim::World *w = shim::world_ptr();
    std::lock_guard<std::mutex> g(w->reg_mu);
    if ((int)w->registry.size() <= id_) w->registry.resize(id_ + 1, std::vector<void *>(w->n, nullptr));
    w->registry[id_][rank_me()] = this;
  }
  dist_object(const dist_object &) = delete;
  dist_object &operator=(const dist_object &) = delete;
  ~dist_object() {
    shim::World *w = shim::world_ptr();
    std::lock_guard<std::mutex> g(w->reg_mu);
    w->registry[id_][rank_me()] = nullptr;
  }

  int id() const { return id_; }
  T *operator->() { return &value_; }
  const T *operator->() const { return &value_; }
  T &operator*() { return value_; }
  const T &operator*() const { return value_; }

  static dist_object *lookup(int id, int rank) {
    shim::World *w = shim::world_ptr();
    for (;;) {
      {
        std::lock_guard<std::mutex> g(w->reg_mu);
        if ((int)w->registry.size() > id && w->registry[id][rank] != nullptr)
          return static_cast<dist_object *>(w->registry[id][rank]);
      }
      progress();
      std::this_thread::yield();
    }
  }

  future<T> fetch(int rank) const { return future<T>(lookup(id_, rank)->value_); }
};

// ----------------------------------------------------------------------------
// RPC
// ----------------------------------------------------------------------------
namespace shim {
// argument binder: plain values are decay-copied; dist_object references are
// carried as an id and re-bound on the target rank.
template <typename A>
struct bound {
  std::decay_t<A> v;
  explicit bound(A &&a) : v(std::forward<A>(a)) {}
  std::decay_t<A> &get() { return v; }
};
template <typename U>
struct bound<dist_object<U> &> {
  int id;
  explicit bound(dist_object<U> &d) : id(d.id()) {}
  dist_object<U> &get() { return *dist_object<U>::lookup(id, rank_me()); }
};
template <typename U>
struct bound<const dist_object<U> &> {
  int id;
  explicit bound(const dist_object<U> &d) : id(d.id()) {}
  dist_object<U> &get() { return *dist_object<U>::lookup(id, rank_me()); }
};
}  // namespace shim

// rpc with a result: executed immediately (shared address space); see header note
namespace shim {
// while alive, rank_me() answers as `rank` (the body of an rpc "runs on the target")
struct RankScope {
  int saved;
  explicit RankScope(int rank) : saved(tl_rank()) { tl_rank() = rank; }
  ~RankScope() { tl_rank() = saved; }
};
}  // namespace shim

template <typename F, typename... Args>
auto rpc(int rank, F &&f, Args &&... args) {
  using R = std::invoke_result_t<F, Args...>;
  shim::RankScope scope(rank);
  if constexpr (std::is_void<R>::value) {
    std::forward<F>(f)(std::forward<Args>(args)...);
    return future<>();
  } else {
    return future<R>(std::forward<F>(f)(std::forward<Args>(args)...));
  }
}

// fire-and-forget rpc: queued to the target, executed inside its progress()
template <typename F, typename... Args>
void rpc_ff(int rank, F f, Args &&... args) {
  auto pack = std::make_shared<std::tuple<shim::bound<Args &&>...>>(
      shim::bound<Args &&>(std::forward<Args>(args))...);
  shim::World *w = shim::world_ptr();
  shim::Inbox &in = *w->inbox[rank];
  std::lock_guard<std::mutex> g(in.mu);
  in.q.emplace_back([f, pack]() {
    std::apply([&](auto &... b) { f(b.get()...); }, *pack);
  });
}

// ----------------------------------------------------------------------------
// collectives
// ----------------------------------------------------------------------------
struct op_fast_add_t {};
constexpr op_fast_add_t op_fast_add{};

template <typename T>
future<> reduce_all(const T *src, T *dst, std::size_t n, op_fast_add_t) {
  shim::World *w = shim::world_ptr();
  w->red_src[rank_me()] = src;
  barrier();
  std::vector<T> tmp(n, T(0));
  for (int r = 0; r < w->n; ++r) {
    const T *s = static_cast<const T *>(w->red_src[r]);
    for (std::size_t i = 0; i < n; ++i) tmp[i] += s[i];
  }
  barrier();  // everyone has finished reading every src (dst may alias src)
  std::memcpy(dst, tmp.data(), n * sizeof(T));
  return future<>();
}

}  // namespace upcxx
