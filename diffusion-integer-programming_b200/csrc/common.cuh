// Shared helpers for the sm_100a kernels of libdip_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/dip_b200.h"

namespace dip {

constexpr int D = DIP_D;

void set_error(const char* fmt, ...);
int sm_count();
extern std::atomic<uint64_t> g_launches;

inline cudaStream_t as_stream(dip_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// In-library kernel timer (dip_prof_*): when enabled, every launcher brackets its launch with a
// pair of CUDA events on the launching stream; dip_prof_collect sums them per kernel class.
struct ProfScope {
    int slot;
    cudaStream_t st;
    ProfScope(int cls, cudaStream_t s);
    ~ProfScope();
};
#define DIP_PROF(cls, stream) dip::ProfScope _prof_scope((cls), dip::as_stream(stream))

#define DIP_FAIL(...)                                  \
    do {                                               \
        dip::set_error(__VA_ARGS__);                   \
        return 1;                                      \
    } while (0)

#define DIP_REQUIRE(cond, ...)                         \
    do {                                               \
        if (!(cond)) DIP_FAIL(__VA_ARGS__);            \
    } while (0)

#define DIP_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) DIP_FAIL("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
    } while (0)

// after a <<<>>> launch
#define DIP_LAUNCHED()                                                                      \
    do {                                                                                    \
        dip::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess) DIP_FAIL("%s:%d launch: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
    } while (0)

#define DIP_TRY(call)                  \
    do {                               \
        int _r = (call);               \
        if (_r) return _r;             \
    } while (0)

// pointer to row r (global row index b*T + t) of a two-source activation
__device__ __forceinline__ const float* row_ptr(const dip_rows& s, int64_t r) {
    if (s.split == 0) return s.batch + r * D;
    int64_t b = r / s.T;
    int t = (int)(r - b * s.T);
    return t < s.split ? s.shared + b * s.shared_stride + (int64_t)t * D : s.batch + (b * (s.T - s.split) + (t - s.split)) * D;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float quick_gelu(float u) {
    // u * sigmoid(1.702 u), model/modules.py:270-272
    return u / (1.0f + expf(-1.702f * u));
}
__device__ __forceinline__ float quick_gelu_grad(float u) {
    float s = 1.0f / (1.0f + expf(-1.702f * u));
    return s + 1.702f * u * s * (1.0f - s);
}
__device__ __forceinline__ float sigmoidf_(float z) { return 1.0f / (1.0f + expf(-z)); }

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace dip
