// include/local_matrix.hpp — cm::LocalMatrix<T>: the update payload container of the
// ConvergentMatrix API (drop-in for the storage/view part of the reference's
// include/local_matrix.hpp:12-85), plus cm::DeviceLocalMatrix<T>, its HBM-resident twin.
//
// A LocalMatrix is a column-major m x n array with leading dimension ld and an optional
// transpose *view* flag; update() reads it through operator()(i,j) semantics:
//     element (i,j) = transposed ? data[j + i*ld] : data[i + ld*j]        (reference :71)
// and m()/n() are the VIEW dimensions (reference :52-54).
//
// Out of scope here (not on the assembly path; reference include/local_matrix.hpp:87-169 and
// include/blas.hpp): the element-wise arithmetic operators and the BLAS-backed operator*.
#pragma once

#include <cassert>
#include <cstddef>

namespace cm {

template <typename T>
class LocalMatrix {
 public:
  // same five constructors as the reference (include/local_matrix.hpp:21-44)
  LocalMatrix() {}
  LocalMatrix(long m, long n) : rows_(m), cols_(n), ld_(m), owns_(true), data_(new T[m * n]) {}
  LocalMatrix(long m, long n, T *data) : rows_(m), cols_(n), ld_(m), data_(data) {}
  LocalMatrix(long m, long n, long ld) : rows_(m), cols_(n), ld_(ld), owns_(true), data_(new T[ld * n]) {}
  LocalMatrix(long m, long n, long ld, T *data) : rows_(m), cols_(n), ld_(ld), data_(data) {}
  LocalMatrix(const LocalMatrix &) = delete;  // the reference double-frees on copy (SURVEY C-5)
  LocalMatrix &operator=(const LocalMatrix &) = delete;
  ~LocalMatrix() {
    if (owns_) delete[] data_;
  }

  // take ownership of a caller-supplied buffer (reference :50)
  void override_free() { owns_ = true; }

  long m() const { return transposed_ ? cols_ : rows_; }
  long n() const { return transposed_ ? rows_ : cols_; }
  T *data() const { return data_; }
  long ld() const { return ld_; }                 // addition: needed by the C-ABI wrapper
  bool transposed() const { return transposed_; }  // addition: ditto

  // heap-allocated, NON-owning transposed view; the caller deletes it (reference :58-62)
  LocalMatrix<T> *trans() {
    LocalMatrix<T> *view = new LocalMatrix<T>(rows_, cols_, ld_, data_);
    view->transposed_ = !transposed_;
    return view;
  }

  T &operator()(long i, long j) const {
#ifndef NOCHECK
    assert(i >= 0 && j >= 0 && i < m() && j < n());
#endif
    return transposed_ ? data_[j + i * ld_] : data_[i + ld_ * j];
  }

  // vector access (reference :74-79)
  T &operator()(long i) const {
#ifndef NOCHECK
    assert((m() == 1 && transposed_) || (n() == 1 && !transposed_));
#endif
    return data_[i];
  }

  // fill every element of the view (reference :81-85)
  LocalMatrix<T> &operator=(T value) {
    for (long j = 0; j < n(); ++j)
      for (long i = 0; i < m(); ++i) (*this)(i, j) = value;
    return *this;
  }

 private:
  long rows_ = 0, cols_ = 0, ld_ = 0;  // storage dims
  bool owns_ = false;
  bool transposed_ = false;
  T *data_ = nullptr;
};

// A payload that already lives in this rank's HBM. Non-owning: wraps a device pointer (from
// cudaMalloc, a torch tensor, ...). ConvergentMatrix::update(DeviceLocalMatrix*, d_ix, d_jx)
// borrows it, stream-ordered, until the next flush()/commit() returns (include/cm_b200.h).
template <typename T>
class DeviceLocalMatrix {
 public:
  DeviceLocalMatrix(long m, long n, const T *device_data) : rows_(m), cols_(n), ld_(m), data_(device_data) {}
  DeviceLocalMatrix(long m, long n, long ld, const T *device_data) : rows_(m), cols_(n), ld_(ld), data_(device_data) {}
  long m() const { return transposed_ ? cols_ : rows_; }
  long n() const { return transposed_ ? rows_ : cols_; }
  long ld() const { return ld_; }
  bool transposed() const { return transposed_; }
  const T *data() const { return data_; }
  DeviceLocalMatrix<T> trans() const {
    DeviceLocalMatrix<T> view(rows_, cols_, ld_, data_);
    view.transposed_ = !transposed_;
    return view;
  }

 private:
  long rows_, cols_, ld_;
  bool transposed_ = false;
  const T *data_;
};

}  // namespace cm
