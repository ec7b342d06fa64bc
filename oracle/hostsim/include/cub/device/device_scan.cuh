// TEST INFRASTRUCTURE ONLY (oracle/hostsim): host stand-in for cub::DeviceScan (two-phase temp-storage protocol kept).
#pragma once
#include <cuda_runtime.h>
namespace cub {
struct DeviceScan {
    template <class In, class Out>
    static cudaError_t ExclusiveSum(void* tmp, size_t& bytes, In in, Out out, int n, cudaStream_t = nullptr) {
        if (!tmp) { bytes = 1; return cudaSuccess; }
        auto acc = decltype(*out + *out)(0);
        for (int i = 0; i < n; ++i) { const auto v = in[i]; out[i] = acc; acc += v; }
        return cudaSuccess;
    }
};
}  // namespace cub
