// tests/emu/include/cub/cub.cuh — TEST DOUBLE of the two CUB primitives csrc/ uses (see
// ../cuda_runtime.h for what tests/emu is and is not).
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <numeric>
#include <vector>

namespace cub {

struct Sum {
  template <class T>
  T operator()(const T &a, const T &b) const {
    return a + b;
  }
};

// block-wide exclusive scan over one item per thread (1-D blocks)
template <class T, int BLOCK>
class BlockScan {
 public:
  struct TempStorage {
    T item[BLOCK];
  };
  explicit BlockScan(TempStorage &t) : t_(t) {}
  template <class Op>
  void ExclusiveScan(T input, T &output, T initial, Op op, T &block_aggregate) {
    const int tid = (int)threadIdx.x;
    t_.item[tid] = input;
    __syncthreads();
    T acc = initial;
    for (int i = 0; i < tid; ++i) acc = op(acc, t_.item[i]);
    output = acc;
    T total = t_.item[0];
    for (int i = 1; i < BLOCK; ++i) total = op(total, t_.item[i]);
    block_aggregate = total;
    __syncthreads();
  }

 private:
  TempStorage &t_;
};

// stable LSD radix sort == a stable sort on the selected key bits
struct DeviceRadixSort {
  template <class K, class V, class N>
  static cudaError_t SortPairs(void *d_temp, size_t &temp_bytes, const K *keys_in, K *keys_out, const V *vals_in,
                               V *vals_out, N n, int begin_bit = 0, int end_bit = (int)sizeof(K) * 8,
                               cudaStream_t = nullptr) {
    if (!d_temp) {
      temp_bytes = 64;
      return cudaSuccess;
    }
    const int bits = end_bit - begin_bit;
    const K mask = bits >= (int)sizeof(K) * 8 ? ~K(0) : (K)(((K(1) << bits) - 1) << begin_bit);
    std::vector<size_t> perm((size_t)n);
    std::iota(perm.begin(), perm.end(), (size_t)0);
    std::stable_sort(perm.begin(), perm.end(),
                     [&](size_t a, size_t b) { return (keys_in[a] & mask) < (keys_in[b] & mask); });
    for (size_t i = 0; i < (size_t)n; ++i) {
      keys_out[i] = keys_in[perm[i]];
      vals_out[i] = vals_in[perm[i]];
    }
    return cudaSuccess;
  }
};

}  // namespace cub
