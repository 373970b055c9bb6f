// fp32 SIMT GEMM with fused epilogue:  C = act(A . W^T + bias + addmat + residual).
// This is the exact-fp32 path of the Linear layers (model/modules.py:293-296, model/diffusion.py:160).
// CTA tile 128x128x16, 256 threads, 8x8 outputs per thread held in registers; operand tiles are
// transposed into shared memory so the inner product reads are conflict-free float4 broadcasts.
#include <cuda_bf16.h>
#include "common.cuh"

namespace dip {

int sm_count();

constexpr int GBM = 128, GBN = 128, GBK = 16, GPAD = 4;

struct GemmParams {
    dip_gemm_args a;
};

__device__ __forceinline__ float apply_act(float v, int act, float aux) {
    switch (act) {
        case DIP_ACT_QUICKGELU: return quick_gelu(v);
        case DIP_ACT_RELU: return fmaxf(v, 0.f);
        case DIP_ACT_SIGMOID: return sigmoidf_(v);
        case DIP_ACT_MUL_QUICKGELU_GRAD: return v * quick_gelu_grad(aux);
        default: return v;
    }
}

__global__ void __launch_bounds__(256) gemm_nt_kernel(const GemmParams P) {
    const dip_gemm_args& g = P.a;
    __shared__ __align__(16) float As[GBK][GBM + GPAD];
    __shared__ __align__(16) float Bs[GBK][GBN + GPAD];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * GBM;
    const int n0 = blockIdx.y * GBN;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // loader mapping: 128 rows x 16 k = 512 float4; thread loads rows (tid>>2) and (tid>>2)+64, k-quad tid&3
    const int lrow = tid >> 2, lkq = tid & 3;

    for (int k0 = 0; k0 < g.K; k0 += GBK) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int r = lrow + 64 * h;
            float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < g.M) av = *reinterpret_cast<const float4*>(g.A + (m0 + r) * g.lda + k0 + lkq * 4);
            As[lkq * 4 + 0][r] = av.x; As[lkq * 4 + 1][r] = av.y; As[lkq * 4 + 2][r] = av.z; As[lkq * 4 + 3][r] = av.w;
            float4 bv = *reinterpret_cast<const float4*>(g.W + (int64_t)(n0 + r) * g.K + k0 + lkq * 4);
            Bs[lkq * 4 + 0][r] = bv.x; Bs[lkq * 4 + 1][r] = bv.y; Bs[lkq * 4 + 2][r] = bv.z; Bs[lkq * 4 + 3][r] = bv.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int64_t r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= g.M) continue;
        int64_t out_row = r, b_in = 0;
        int t_in = 0;
        if (g.rows_in > 0) {
            int64_t b = r / g.rows_in;
            t_in = (int)(r - b * g.rows_in);
            out_row = b * g.rows_out + g.row_off + t_in;
            b_in = b;
        }
        float rbs = g.row_bias_scale ? g.row_bias_scale[r] : 1.0f;
        const float* resp = g.res.batch ? row_ptr(g.res, out_row) : nullptr;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int c = n0 + h * 64 + tx * 4;
            float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
            if (g.bias) {
                float4 bb = *reinterpret_cast<const float4*>(g.bias + c);
                v[0] += bb.x * rbs; v[1] += bb.y * rbs; v[2] += bb.z * rbs; v[3] += bb.w * rbs;
            }
            if (g.addmat) {
                float4 am = *reinterpret_cast<const float4*>(g.addmat + b_in * g.addmat_stride + (int64_t)t_in * g.N + c);
                v[0] += am.x; v[1] += am.y; v[2] += am.z; v[3] += am.w;
            }
            if (resp) {
                // residual rows are 128 wide (only used with N == 128)
                float4 rr = *reinterpret_cast<const float4*>(resp + c);
                v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
            }
            if (g.preact) *reinterpret_cast<float4*>(g.preact + out_row * g.ldc + c) = make_float4(v[0], v[1], v[2], v[3]);
            if (g.act != DIP_ACT_NONE) {
                float4 ax = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g.act == DIP_ACT_MUL_QUICKGELU_GRAD) ax = *reinterpret_cast<const float4*>(g.aux + out_row * g.ldc + c);
                v[0] = apply_act(v[0], g.act, ax.x); v[1] = apply_act(v[1], g.act, ax.y);
                v[2] = apply_act(v[2], g.act, ax.z); v[3] = apply_act(v[3], g.act, ax.w);
            }
            *reinterpret_cast<float4*>(g.C + out_row * g.ldc + c) = make_float4(v[0], v[1], v[2], v[3]);
            if (g.split_hi) {
                __nv_bfloat16 h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    h[e] = __float2bfloat16_rn(v[e]);
                    l[e] = __float2bfloat16_rn(v[e] - __bfloat162float(h[e]));
                }
                uint2 hv, lv;
                hv.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
                hv.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
                lv.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
                lv.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.split_hi) + out_row * g.ldc + c) = hv;
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.split_lo) + out_row * g.ldc + c) = lv;
            }
        }
    }
}

// tiny-K linear for the 5 / 17 raw feature columns: one thread per output element
__global__ void __launch_bounds__(256) linear_smallk_kernel(float* __restrict__ y, const float* __restrict__ x, int64_t rows, int K,
                                                            const float* __restrict__ shift, const float* __restrict__ scale,
                                                            const float* __restrict__ W, const float* __restrict__ bias, int act) {
    int64_t total = rows * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / D;
        int j = (int)(i - r * D);
        float acc = 0.f;
        for (int k = 0; k < K; ++k) {
            float v = x[r * K + k];
            if (shift) v = v + shift[k];
            if (scale) v = v * scale[k];
            acc = fmaf(v, W[j * K + k], acc);
        }
        if (bias) acc += bias[j];
        y[i] = apply_act(acc, act, 0.f);
    }
}

}  // namespace dip

using namespace dip;

extern "C" {

int dip_gemm_nt(const dip_gemm_args* args, dip_stream_t stream) {
    DIP_REQUIRE(args && args->A && args->W && args->C, "dip_gemm_nt: null argument");
    DIP_REQUIRE(args->in_T == 0, "dip_gemm_nt: input row windows are a dip_gemm_tc feature");
    const dip_gemm_args& g = *args;
    DIP_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "dip_gemm_nt: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
    DIP_REQUIRE(g.N % GBN == 0 && g.K % GBK == 0, "dip_gemm_nt: N must be a multiple of 128 and K of 16 (N=%d K=%d)", g.N, g.K);
    DIP_REQUIRE(g.lda % 4 == 0 && g.ldc % 4 == 0, "dip_gemm_nt: lda/ldc must be multiples of 4");
    DIP_REQUIRE(g.res.batch == nullptr || g.N == D, "dip_gemm_nt: residual needs N == 128");
    DIP_REQUIRE(g.act != DIP_ACT_MUL_QUICKGELU_GRAD || g.aux, "dip_gemm_nt: aux missing");
    DIP_REQUIRE((g.split_hi == nullptr) == (g.split_lo == nullptr), "dip_gemm_nt: split_hi and split_lo go together");
    DIP_REQUIRE(g.rows_in == 0 || (g.rows_out >= g.rows_in + g.row_off), "dip_gemm_nt: bad row remap");
    GemmParams P;
    P.a = g;
    dim3 grid(cdiv(g.M, GBM), g.N / GBN);
    DIP_PROF(DIP_PROF_GEMM, stream);
    gemm_nt_kernel<<<grid, 256, 0, as_stream(stream)>>>(P);
    DIP_LAUNCHED();
    return 0;
}

int dip_linear_smallk(float* y, const float* x, int64_t rows, int K, const float* shift, const float* scale, const float* W,
                      const float* bias, int act, dip_stream_t stream) {
    DIP_REQUIRE(y && x && W && rows > 0 && K > 0 && K <= 64, "dip_linear_smallk: bad argument");
    int64_t blocks = (rows * D + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    linear_smallk_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(y, x, rows, K, shift, scale, W, bias, act);
    DIP_LAUNCHED();
    return 0;
}

}  // extern "C"
