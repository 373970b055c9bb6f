// Linear layers on tcgen05:  C = act(LN?(A) . W^T + bias + addmat + residual)  with the same fused
// epilogue set as the fp32 SIMT GEMM, "bf16x2 split" numerics (tc_common.cuh), optional LayerNorm fused
// into the operand load (every CTA tile holds complete 128-feature rows, so the statistics are
// row-local).  These GEMMs have only 32 FLOP per HBM byte: the kernel is built to stream --
// persistent CTAs, warp-specialised:
//   warps 0-3  loaders : coalesced fp32 row loads (one warp per row, float4 per lane), optional LN by warp
//                        shuffles, split into bf16 hi/lo, swizzled K-major stores into a 2-stage ring
//   warp  4    MMA     : one thread issues 3 passes x K/16 tcgen05.mma per stage into a double-buffered
//                        TMEM accumulator; weights (bf16 hi/lo planes, TMA, SWIZZLE_128B) stay resident
//   warps 5-8  epilogue: thread = row; TMEM -> registers -> bias / residual / activation -> global
//                        (fp32 and/or bf16 hi/lo planes for the attention kernels)
// Supported shapes: K = 128 (BN = 128 columns per CTA column-block) and K = 384 (BN = 64), N % BN == 0.
#include "tc_common.cuh"

namespace dip {
namespace tcgemm {

using namespace tc;

constexpr uint32_t A_PLANE = 32768;                 // [2 kb][128 rows][64] bf16
constexpr uint32_t A_STAGE = 2 * A_PLANE;           // hi + lo
constexpr uint32_t OFF_W = 2 * A_STAGE;             // after the 2-stage A ring (128 KB)
constexpr uint32_t W_BYTES = 98304;                 // up to 96 KB of weight planes
constexpr uint32_t OFF_BAR = OFF_W + W_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256;

enum { G_WFULL = 0, G_AFULL = 1, G_AEMPTY = 3, G_CFULL = 5, G_CEMPTY = 7 };

struct Params {
    dip_gemm_args g;
    dip_rows a_rows;          // A source (two-source rows allowed when K == 128); a_rows.batch == NULL -> g.A / g.lda
    const float* ln_gamma;    // LayerNorm prologue (K == 128 only), NULL = off
    const float* ln_beta;
    float* ln_mean;           // optional statistics output (written by column-block 0)
    float* ln_rstd;
    int n_tiles;
};

__device__ __forceinline__ float act_apply(float v, int act, float aux) {
    switch (act) {
        case DIP_ACT_QUICKGELU: return quick_gelu(v);
        case DIP_ACT_RELU: return fmaxf(v, 0.f);
        case DIP_ACT_SIGMOID: return sigmoidf_(v);
        case DIP_ACT_MUL_QUICKGELU_GRAD: return v * quick_gelu_grad(aux);
        default: return v;
    }
}

// one float4 of one output row through the whole epilogue (bias, addmat, residual, preact, activation, stores)
__device__ __forceinline__ void emit4(const dip_gemm_args& g, int64_t row, int c, float (&x)[4]) {
    int64_t out_row = row;
    int t_in = 0;
    if (g.rows_in > 0) {
        const int64_t bb = row / g.rows_in;
        t_in = (int)(row - bb * g.rows_in);
        out_row = bb * g.rows_out + g.row_off + t_in;
    }
    if (g.bias) {
        const float rbs = g.row_bias_scale ? g.row_bias_scale[row] : 1.0f;
        const float4 bv = *reinterpret_cast<const float4*>(g.bias + c);
        x[0] += bv.x * rbs; x[1] += bv.y * rbs; x[2] += bv.z * rbs; x[3] += bv.w * rbs;
    }
    if (g.addmat) {
        const float4 am = *reinterpret_cast<const float4*>(g.addmat + (int64_t)t_in * g.N + c);
        x[0] += am.x; x[1] += am.y; x[2] += am.z; x[3] += am.w;
    }
    if (g.res.batch) {
        const float4 rr = *reinterpret_cast<const float4*>(row_ptr(g.res, out_row) + c);
        x[0] += rr.x; x[1] += rr.y; x[2] += rr.z; x[3] += rr.w;
    }
    if (g.preact) *reinterpret_cast<float4*>(g.preact + out_row * g.ldc + c) = make_float4(x[0], x[1], x[2], x[3]);
    if (g.act != DIP_ACT_NONE) {
        float4 ax = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.act == DIP_ACT_MUL_QUICKGELU_GRAD) ax = *reinterpret_cast<const float4*>(g.aux + out_row * g.ldc + c);
        x[0] = act_apply(x[0], g.act, ax.x); x[1] = act_apply(x[1], g.act, ax.y);
        x[2] = act_apply(x[2], g.act, ax.z); x[3] = act_apply(x[3], g.act, ax.w);
    }
    if (g.C) *reinterpret_cast<float4*>(g.C + out_row * g.ldc + c) = make_float4(x[0], x[1], x[2], x[3]);
    if (g.split_hi) {
        uint32_t h0, l0, h1, l1;
        split2(x[0], x[1], h0, l0);
        split2(x[2], x[3], h1, l1);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.split_hi) + out_row * g.ldc + c) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.split_lo) + out_row * g.ldc + c) = make_uint2(l0, l1);
    }
}

constexpr int STG_LD = 36;                          // staging row stride in floats (conflict-free float4 access)
constexpr uint32_t STG_WARP = 32 * STG_LD * 4;      // per epilogue warp

template <int KCH, int BN>
__global__ void __launch_bounds__(288, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo, const Params P) {
    extern __shared__ __align__(1024) uint8_t sm[];
    if ((smem_u32(sm) & 1023u) != 0u) __trap();
    uint8_t* sW = sm + OFF_W;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const dip_gemm_args& g = P.g;
    constexpr uint32_t W_BLOCK = BN * 128;                 // one 64-column k-block of BN weight rows
    constexpr uint32_t W_PLANE = 2 * KCH * W_BLOCK;
    constexpr uint32_t IDESC = instr_desc(128, BN, 0, 0);
    // K = 128: the weights take 64 KB of the 96 KB region; the rest stages the output tile for coalesced stores
    constexpr bool STAGED = (2 * W_PLANE + 4 * STG_WARP <= W_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * BN;

    if (threadIdx.x == 0) {
        mbar_init(&bars[G_WFULL], 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars[G_AFULL + s], 128);
            mbar_init(&bars[G_AEMPTY + s], 1);
            mbar_init(&bars[G_CFULL + s], 1);
            mbar_init(&bars[G_CEMPTY + s], 128);
        }
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 4) {
        // ------------------------------------------------------------------ loaders
        float4 gam = make_float4(1.f, 1.f, 1.f, 1.f), bet = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool ln = P.ln_gamma != nullptr;
        if (ln) {
            gam = reinterpret_cast<const float4*>(P.ln_gamma)[lane];
            bet = reinterpret_cast<const float4*>(P.ln_beta)[lane];
        }
        // lane l holds columns 4l..4l+3 of its row: k-block l/16, 16-byte chunk (l%16)/2, half l%2
        const int kb = lane >> 4, chunk = (lane & 15) >> 1, half = lane & 1;
        int it = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
            for (int kc = 0; kc < KCH; ++kc, ++it) {
                const int s = it & 1;
                if (it >= 2) mbar_wait(&bars[G_AEMPTY + s], ((it >> 1) - 1) & 1);
                uint8_t* dst = sm + s * A_STAGE;
#pragma unroll 1
                for (int rr0 = 0; rr0 < 32; rr0 += 8) {
                    float4 v[8];
                    // 8 independent 512-byte row loads in flight per warp before any is consumed
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int64_t row = (int64_t)tile * 128 + warp * 32 + rr0 + u;
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (row < g.M) {
                            const float* src = P.a_rows.batch ? row_ptr(P.a_rows, row) : g.A + row * (int64_t)g.lda + kc * 128;
                            v[u] = reinterpret_cast<const float4*>(src)[lane];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int r = warp * 32 + rr0 + u;
                        const int64_t row = (int64_t)tile * 128 + r;
                        float4 x = v[u];
                        if (ln && row < g.M) {
                            const float mu = warp_sum(x.x + x.y + x.z + x.w) * (1.0f / D);
                            const float dx = x.x - mu, dy = x.y - mu, dz = x.z - mu, dw = x.w - mu;
                            const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / D);
                            float rs = rsqrtf(var + 1e-5f);
                            rs = rs * (1.5f - 0.5f * (var + 1e-5f) * rs * rs);
                            x = make_float4(dx * rs * gam.x + bet.x, dy * rs * gam.y + bet.y, dz * rs * gam.z + bet.z, dw * rs * gam.w + bet.w);
                            if (P.ln_mean && blockIdx.y == 0 && lane == 0) {
                                P.ln_mean[row] = mu;
                                P.ln_rstd[row] = rs;
                            }
                        }
                        uint32_t h0, l0, h1, l1;
                        split2(x.x, x.y, h0, l0);
                        split2(x.z, x.w, h1, l1);
                        const uint32_t off = kb * 16384 + sw128_off(r, chunk) + half * 8;
                        *reinterpret_cast<uint2*>(dst + off) = make_uint2(h0, h1);
                        *reinterpret_cast<uint2*>(dst + A_PLANE + off) = make_uint2(l0, l1);
                    }
                }
                fence_proxy_async();
                mbar_arrive(&bars[G_AFULL + s]);
            }
        }
    } else if (warp == 4) {
        // ------------------------------------------------------------------ weights (TMA) + MMA issue
        if (lane == 0) {
            mbar_expect_tx(&bars[G_WFULL], 2 * W_PLANE);
            for (int kb = 0; kb < 2 * KCH; ++kb) {
                tma_load_2d(sW + kb * W_BLOCK, &tm_w_hi, &bars[G_WFULL], kb * 64, n0);
                tma_load_2d(sW + W_PLANE + kb * W_BLOCK, &tm_w_lo, &bars[G_WFULL], kb * 64, n0);
            }
            mbar_wait(&bars[G_WFULL], 0);
            const uint32_t aA = smem_u32(sm), aW = smem_u32(sW);
            int it = 0, ti = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++ti) {
                const int cb = ti & 1;
                if (ti >= 2) mbar_wait(&bars[G_CEMPTY + cb], ((ti >> 1) - 1) & 1);
                uint32_t acc = 0;
                for (int kc = 0; kc < KCH; ++kc, ++it) {
                    const int s = it & 1;
                    mbar_wait(&bars[G_AFULL + s], (it >> 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t a = aA + s * A_STAGE + (pass == 2 ? A_PLANE : 0);      // hi, hi, lo
                        const uint32_t w = aW + (pass == 1 ? W_PLANE : 0);                    // hi, lo, hi
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                tc_mma(tmem + cb * BN, smem_desc(a + kb * 16384 + ks * 32, 16, 1024),
                                       smem_desc(w + (kc * 2 + kb) * W_BLOCK + ks * 32, 16, 1024), IDESC, acc);
                                acc = 1;
                            }
                    }
                    tc_commit(&bars[G_AEMPTY + s]);
                }
                tc_commit(&bars[G_CFULL + cb]);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: TMEM lane quarter = warp % 4
        const int q = warp & 3;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        float* stg = reinterpret_cast<float*>(sW + 2 * W_PLANE + (warp - 5) * STG_WARP);
        int ti = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++ti) {
            const int cb = ti & 1;
            mbar_wait(&bars[G_CFULL + cb], (ti >> 1) & 1);
            tc_fence_after();
            const int64_t row0 = (int64_t)tile * 128 + q * 32;
#pragma unroll
            for (int ch = 0; ch < BN / 32; ++ch) {
                float v[32];
                tmem_ld32(lane_addr + cb * BN + ch * 32, v);
                if (ch == BN / 32 - 1) {
                    tc_fence_before();
                    mbar_arrive(&bars[G_CEMPTY + cb]);           // accumulator drained: the MMA warp may reuse it
                }
                const int c0 = n0 + ch * 32;
                if (STAGED) {
                    // transpose through shared memory: afterwards every warp instruction touches 8 rows x 128 contiguous bytes
#pragma unroll
                    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(stg + lane * STG_LD + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    __syncwarp();
#pragma unroll
                    for (int rr0 = 0; rr0 < 32; rr0 += 4) {
                        const int rr = rr0 + (lane >> 3), cc = (lane & 7) * 4;
                        const float4 t = *reinterpret_cast<const float4*>(stg + rr * STG_LD + cc);
                        float x[4] = {t.x, t.y, t.z, t.w};
                        if (row0 + rr < g.M) emit4(g, row0 + rr, c0 + cc, x);
                    }
                    __syncwarp();
                } else {
                    const int64_t row = row0 + lane;
                    if (row < g.M) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float x[4] = {v[i], v[i + 1], v[i + 2], v[i + 3]};
                            emit4(g, row, c0 + i, x);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 2 * BN);
}

template <int KCH, int BN>
static int launch(const CUtensorMap& whi, const CUtensorMap& wlo, const Params& P, int n_blocks_y, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        DIP_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<KCH, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr = true;
    }
    int gx = sm_count() / n_blocks_y;
    if (gx < 1) gx = 1;
    if (gx > P.n_tiles) gx = P.n_tiles;
    dim3 grid(gx, n_blocks_y);
    gemm_tc_kernel<KCH, BN><<<grid, 288, SMEM_BYTES, st>>>(whi, wlo, P);
    return 0;
}

}  // namespace tcgemm
}  // namespace dip

using namespace dip;

extern "C" {

int dip_gemm_tc(const dip_gemm_args* args, const void* w_hi, const void* w_lo, const dip_rows* a_rows, const float* ln_gamma,
                const float* ln_beta, float* ln_mean, float* ln_rstd, dip_stream_t stream) {
    DIP_REQUIRE(args && w_hi && w_lo, "dip_gemm_tc: null argument");
    const dip_gemm_args& g = *args;
    DIP_REQUIRE((g.A || (a_rows && a_rows->batch)) && (g.C || g.split_hi || g.preact), "dip_gemm_tc: no input / no output");
    DIP_REQUIRE(g.M > 0 && (g.K == 128 || g.K == 384), "dip_gemm_tc: K must be 128 or 384 (K=%d)", g.K);
    const int BN = g.K == 128 ? 128 : 64;
    DIP_REQUIRE(g.N % BN == 0, "dip_gemm_tc: N must be a multiple of %d (N=%d)", BN, g.N);
    DIP_REQUIRE(g.ldc % 4 == 0 && (g.A == nullptr || g.lda % 4 == 0), "dip_gemm_tc: lda/ldc must be multiples of 4");
    DIP_REQUIRE(g.res.batch == nullptr || g.N == D, "dip_gemm_tc: residual needs N == 128");
    DIP_REQUIRE(g.act != DIP_ACT_MUL_QUICKGELU_GRAD || g.aux, "dip_gemm_tc: aux missing");
    DIP_REQUIRE((g.split_hi == nullptr) == (g.split_lo == nullptr), "dip_gemm_tc: split_hi and split_lo go together");
    DIP_REQUIRE(!(a_rows && a_rows->batch) || g.K == 128, "dip_gemm_tc: row sources need K == 128");
    DIP_REQUIRE(ln_gamma == nullptr || (g.K == 128 && ln_beta), "dip_gemm_tc: LayerNorm prologue needs K == 128 and beta");
    DIP_REQUIRE((ln_mean == nullptr) == (ln_rstd == nullptr), "dip_gemm_tc: mean and rstd go together");
    tcgemm::Params P;
    P.g = g;
    if (a_rows && a_rows->batch) P.a_rows = *a_rows;
    else { P.a_rows.shared = nullptr; P.a_rows.batch = nullptr; P.a_rows.T = 0; P.a_rows.split = 0; }
    P.ln_gamma = ln_gamma; P.ln_beta = ln_beta; P.ln_mean = ln_mean; P.ln_rstd = ln_rstd;
    P.n_tiles = cdiv(g.M, 128);
    CUtensorMap whi, wlo;
    DIP_TRY(make_tmap_bf16_2d(&whi, w_hi, g.N, g.K, g.K, BN, 64));
    DIP_TRY(make_tmap_bf16_2d(&wlo, w_lo, g.N, g.K, g.K, BN, 64));
    DIP_PROF(DIP_PROF_GEMM, stream);
    if (g.K == 128) DIP_TRY((tcgemm::launch<1, 128>(whi, wlo, P, g.N / 128, as_stream(stream))));
    else DIP_TRY((tcgemm::launch<3, 64>(whi, wlo, P, g.N / 64, as_stream(stream))));
    DIP_LAUNCHED();
    return 0;
}

}  // extern "C"
