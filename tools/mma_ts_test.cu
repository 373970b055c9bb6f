// Development test: tcgen05.mma with the A operand in tensor memory (TS mode): layout check + rate.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
namespace dip { void set_error(const char*, ...) {} std::atomic<uint64_t> g_launches{0}; int sm_count() { return 148; } }
using namespace dip::tc;

__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// A[i][k] = ((i * 3 + k * 5) % 7) - 3, B[n][k] = ((n + 2 * k) % 5) - 2, K = 64 (4 k-steps), N = 64
template <int N>
__global__ void __launch_bounds__(128, 1) ts_kernel(float* dout, long long* tim, int n_time) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    // B tile [N rows x 64 k] bf16, K-major SWIZZLE_128B
    for (int idx = threadIdx.x; idx < N * 8; idx += 128) {
        const int n = idx >> 3, c = idx & 7;
        uint32_t w[4];
        for (int e = 0; e < 4; ++e) {
            const int k0 = c * 8 + 2 * e;
            const float b0 = (float)(((n + 2 * k0) % 5) - 2), b1 = (float)(((n + 2 * (k0 + 1)) % 5) - 2);
            w[e] = pack2(__float2bfloat16_rn(b0), __float2bfloat16_rn(b1));
        }
        *reinterpret_cast<uint4*>(sm + sw128_off(n, c)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    // A operand into TMEM columns [256, 256+32): lane = row, column c holds (A[row][2c], A[row][2c+1])
    for (int blk = 0; blk < 4; ++blk) {
        uint32_t r[8];
        for (int c = 0; c < 8; ++c) {
            const int k0 = blk * 16 + 2 * c;
            const float a0 = (float)(((row * 3 + k0 * 5) % 7) - 3), a1 = (float)(((row * 3 + (k0 + 1) * 5) % 7) - 3);
            r[c] = pack2(__float2bfloat16_rn(a0), __float2bfloat16_rn(a1));
        }
        tmem_st8(lane_addr + 256 + blk * 8, r);
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = instr_desc(128, N, 0, 0);
        for (int ks = 0; ks < 4; ++ks) tc_mma_ts(tmem, tmem + 256 + ks * 8, smem_desc(smem_u32(sm) + ks * 32, 16, 1024), idesc, ks > 0);
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        // rate: n_time more MMAs into other columns
        long long t0 = clock64();
        for (int i = 0; i < n_time; ++i) tc_mma_ts(tmem + 128, tmem + 256 + (i & 3) * 8, smem_desc(smem_u32(sm) + (i & 3) * 32, 16, 1024), idesc, i > 0);
        long long t1 = clock64();
        tc_commit(&bar);
        mbar_wait(&bar, 1);
        long long t2 = clock64();
        tim[0] = t1 - t0; tim[1] = t2 - t0;
    }
    __syncthreads();
    tc_fence_after();
    for (int ch = 0; ch < N / 32; ++ch) {
        float v[32];
        tmem_ld32(lane_addr + ch * 32, v);
        for (int i = 0; i < 32; ++i) dout[row * N + ch * 32 + i] = v[i];
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N>
void run() {
    float* d; long long* t;
    cudaMalloc(&d, 128 * N * 4); cudaMalloc(&t, 16);
    cudaFuncSetAttribute(ts_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    ts_kernel<N><<<1, 128, 65536>>>(d, t, 2048);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> h(128 * N); long long ht[2];
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(ht, t, 16, cudaMemcpyDeviceToHost);
    int bad = 0; double maxerr = 0;
    for (int i = 0; i < 128; ++i) for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < 64; ++k) ref += (double)(((i * 3 + k * 5) % 7) - 3) * (double)(((n + 2 * k) % 5) - 2);
        double err = fabs(ref - h[i * N + n]); if (err > 1e-3) ++bad; if (err > maxerr) maxerr = err;
    }
    printf("TS mode N=%3d: %s (mismatches %d, max err %.3g); issue %.1f cyc/mma, complete %.1f cyc/mma (ideal %d) %s\n", N, bad ? "LAYOUT WRONG" : "layout OK",
           bad, maxerr, ht[0] / 2048.0, ht[1] / 2048.0, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() { run<32>(); run<64>(); run<128>(); run<256>(); return 0; }
