// Partial-solution hand-off on the device (reference: sample_partialsol.py:172-199): for every decoded sample draw the
// `int(n * coverage)` variables that get a value in the SCIP partial solution, and emit the selection mask and the selected values
// bit-packed (numpy.packbits order: index 8*j is bit 7 of byte j), so the host solver workers receive ceil(n/8) + ceil(n/8) bytes per
// sample instead of a Python dict.  The reference draws with Python's Mersenne twister (`random.sample`), a sequential stream that
// ties the result to the order in which samples are visited; here the subset is a pure function of (seed, instance, sample): index i
// gets the 64-bit key (philox(seed; i, sample, instance) << 32) | i and the k smallest keys are selected -- a uniform k-subset,
// identical for any sharding of samples over GPUs.  One CTA per sample; the k-th smallest key is found by an 8-pass radix select
// (256-bin histograms in shared memory), keys are recomputed instead of stored.
#include "common.cuh"

namespace dip {

__device__ __forceinline__ uint32_t philox_word(uint32_t i, uint32_t sample, uint64_t instance, uint64_t seed) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t c0 = i, c1 = sample, c2 = (uint32_t)instance, c3 = (uint32_t)(instance >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0, hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

__device__ __forceinline__ uint64_t select_key(int i, int sample, uint64_t instance, uint64_t seed) {
    return ((uint64_t)philox_word((uint32_t)i, (uint32_t)sample, instance, seed) << 32) | (uint32_t)i;
}

__global__ void __launch_bounds__(256) partial_select_kernel(uint8_t* __restrict__ mask_bits, uint8_t* __restrict__ value_bits,
                                                            const uint8_t* __restrict__ xbits, int n, int k, uint64_t seed, uint64_t instance,
                                                            int sample_base, int first_sample_values) {
    __shared__ uint32_t hist[256];
    __shared__ uint64_t s_prefix;
    __shared__ int s_k;
    const int b = blockIdx.x, sample = sample_base + b;
    const int nbytes = (n + 7) / 8;
    uint64_t thr = 0;                      // keys <= thr are selected
    if (k >= n) thr = ~0ull;
    if (k > 0 && k < n) {
        if (threadIdx.x == 0) { s_prefix = 0; s_k = k; }
        for (int pass = 7; pass >= 0; --pass) {
            hist[threadIdx.x] = 0;
            __syncthreads();
            const uint64_t prefix = s_prefix;
            const uint64_t himask = pass == 7 ? 0ull : (~0ull << (8 * (pass + 1)));
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const uint64_t key = select_key(i, sample, instance, seed);
                if ((key & himask) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1u);     // integer counts: order-independent
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int need = s_k, bin = 0;
                while (bin < 255 && (int)hist[bin] < need) { need -= (int)hist[bin]; ++bin; }
                s_k = need;                                                // rank of the wanted key inside this bin
                s_prefix = prefix | ((uint64_t)bin << (8 * pass));
            }
            __syncthreads();
        }
        thr = s_prefix;                    // the k-th smallest key itself (keys are unique: the low word is the index)
    }
    const uint8_t* xrow = xbits + (int64_t)(first_sample_values ? 0 : b) * n;      // sample_partialsol.py:179 reads pred_x[0, k] for every j
    for (int j = threadIdx.x; j < nbytes; j += blockDim.x) {
        uint32_t m = 0, v = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int i = 8 * j + t;
            if (i < n && k > 0 && select_key(i, sample, instance, seed) <= thr) {
                m |= 0x80u >> t;
                if (xrow[i]) v |= 0x80u >> t;
            }
        }
        mask_bits[(int64_t)b * nbytes + j] = (uint8_t)m;
        value_bits[(int64_t)b * nbytes + j] = (uint8_t)v;
    }
}

// out = a * x + b * y   (y nullable), n % 4 == 0
__global__ void __launch_bounds__(256) axpby_kernel(float4* __restrict__ out, const float4* __restrict__ x, const float4* __restrict__ y, float a,
                                                    float b, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 u = x[i], r;
        if (y) {
            float4 w = y[i];
            r = make_float4(a * u.x + b * w.x, a * u.y + b * w.y, a * u.z + b * w.z, a * u.w + b * w.w);
        } else {
            r = make_float4(a * u.x, a * u.y, a * u.z, a * u.w);
        }
        out[i] = r;
    }
}

}  // namespace dip

using namespace dip;

extern "C" {

int dip_partial_select(uint8_t* mask_bits, uint8_t* value_bits, const uint8_t* xbits, int B, int n, int k, uint64_t seed, uint64_t instance_id,
                       int sample_base, int first_sample_values, dip_stream_t stream) {
    DIP_REQUIRE(mask_bits && value_bits && xbits && B > 0 && n > 0, "dip_partial_select: null/empty argument");
    DIP_REQUIRE(k >= 0 && k <= n, "dip_partial_select: k = %d outside [0, n = %d]", k, n);
    DIP_PROF(DIP_PROF_OTHER, stream);
    partial_select_kernel<<<B, 256, 0, as_stream(stream)>>>(mask_bits, value_bits, xbits, n, k, seed, instance_id, sample_base, first_sample_values);
    DIP_LAUNCHED();
    return 0;
}

int dip_axpby(float* out, const float* x, const float* y, float a, float b, int64_t n, dip_stream_t stream) {
    DIP_REQUIRE(out && x && n > 0 && n % 4 == 0, "dip_axpby: bad argument");
    DIP_PROF(DIP_PROF_OTHER, stream);
    int64_t n4 = n / 4, blocks = (n4 + 255) / 256, cap = (int64_t)sm_count() * 8;
    axpby_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(out), reinterpret_cast<const float4*>(x),
                                                                                   reinterpret_cast<const float4*>(y), a, b, n4);
    DIP_LAUNCHED();
    return 0;
}

}  // extern "C"
