// TEST INFRASTRUCTURE ONLY (oracle/hostsim): host stand-in for cub::DeviceRadixSort::SortKeys (unsigned keys, all bits significant
// below end_bit: a plain sort is the same order).
#pragma once
#include <cuda_runtime.h>
namespace cub {
struct DeviceRadixSort {
    template <class K>
    static cudaError_t SortKeys(void* tmp, size_t& bytes, const K* in, K* out, int n, int = 0, int = sizeof(K) * 8, cudaStream_t = nullptr) {
        if (!tmp) { bytes = 1; return cudaSuccess; }
        std::copy(in, in + n, out);
        std::sort(out, out + n);
        return cudaSuccess;
    }
};
}  // namespace cub
