// Exact-fp32 flash attention, one head of width 128, with key-padding mask: forward (saves the
// log-sum-exp) and deterministic input-gradient backward.  This is the fp32 reference path of
// nn.MultiheadAttention's scaled_dot_product_attention (model/modules.py:291) and of its autograd
// backward (model/diffusion.py:634); it never materialises the T x T score matrix.
//
// Tiles are 64 (queries) x 64 (keys) x 128; 256 threads; score tiles use a 4x4 register micro-tile
// per thread (rows ty+16*ii, cols tx+16*jj), output tiles a 4x8 micro-tile.  Shared-memory rows are
// padded to 132 / 68 floats so every float4 access below is bank-conflict-free.
#include <math.h>
#include "common.cuh"

namespace dip {

constexpr int AT = 64;
constexpr int ALD = D + 4;    // row stride of [64][128] tiles
constexpr int PLD = AT + 4;   // row stride of [64][64] tiles
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float SM_SCALE = 0.08838834764831845f;   // 1/sqrt(128)

// copy a 64 x 128 tile (rows t0.. of sample b) into shared memory, zero-filling rows >= T
__device__ __forceinline__ void load_tile(float* __restrict__ dst, const float* __restrict__ src, int ld, int64_t row0, int t0, int T) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        int idx = threadIdx.x + 256 * u;
        int r = idx >> 5, c4 = idx & 31;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t0 + r < T) v = *reinterpret_cast<const float4*>(src + (row0 + r) * (int64_t)ld + c4 * 4);
        *reinterpret_cast<float4*>(dst + r * ALD + c4 * 4) = v;
    }
}

// acc[ii][jj] += sum_k X[ty+16ii][k] * Y[tx+16jj][k]
__device__ __forceinline__ void tile_dot(float (&acc)[4][4], const float* __restrict__ X, const float* __restrict__ Y, int tx, int ty) {
#pragma unroll 4
    for (int k = 0; k < D; k += 4) {
        float4 xv[4], yv[4];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) xv[ii] = *reinterpret_cast<const float4*>(X + (ty + 16 * ii) * ALD + k);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) yv[jj] = *reinterpret_cast<const float4*>(Y + (tx + 16 * jj) * ALD + k);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                acc[ii][jj] = fmaf(xv[ii].x, yv[jj].x, acc[ii][jj]);
                acc[ii][jj] = fmaf(xv[ii].y, yv[jj].y, acc[ii][jj]);
                acc[ii][jj] = fmaf(xv[ii].z, yv[jj].z, acc[ii][jj]);
                acc[ii][jj] = fmaf(xv[ii].w, yv[jj].w, acc[ii][jj]);
            }
    }
}

// acc[r][0..7] += sum_{j<64} Pm[prow(r)][j] * Yt[j][cols]   (Pm row-major [64][PLD], rows ty*4+r)
__device__ __forceinline__ void tile_pv_rows(float (&acc)[4][8], const float* __restrict__ Pm, const float* __restrict__ Yt, int tx, int ty) {
#pragma unroll 2
    for (int j = 0; j < AT; j += 4) {
        float4 pv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) pv[r] = *reinterpret_cast<const float4*>(Pm + (ty * 4 + r) * PLD + j);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            float4 y0 = *reinterpret_cast<const float4*>(Yt + (j + jj) * ALD + tx * 4);
            float4 y1 = *reinterpret_cast<const float4*>(Yt + (j + jj) * ALD + 64 + tx * 4);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float p = jj == 0 ? pv[r].x : jj == 1 ? pv[r].y : jj == 2 ? pv[r].z : pv[r].w;
                acc[r][0] = fmaf(p, y0.x, acc[r][0]); acc[r][1] = fmaf(p, y0.y, acc[r][1]);
                acc[r][2] = fmaf(p, y0.z, acc[r][2]); acc[r][3] = fmaf(p, y0.w, acc[r][3]);
                acc[r][4] = fmaf(p, y1.x, acc[r][4]); acc[r][5] = fmaf(p, y1.y, acc[r][5]);
                acc[r][6] = fmaf(p, y1.z, acc[r][6]); acc[r][7] = fmaf(p, y1.w, acc[r][7]);
            }
        }
    }
}

// acc[r][0..7] += sum_{i<64} Pm[i][ty*4+r] * Yt[i][cols]   (transposed use of Pm: rows of the result are columns of Pm)
__device__ __forceinline__ void tile_ptv_cols(float (&acc)[4][8], const float* __restrict__ Pm, const float* __restrict__ Yt, int tx, int ty) {
#pragma unroll 4
    for (int i = 0; i < AT; ++i) {
        float4 pv = *reinterpret_cast<const float4*>(Pm + i * PLD + ty * 4);
        float4 y0 = *reinterpret_cast<const float4*>(Yt + i * ALD + tx * 4);
        float4 y1 = *reinterpret_cast<const float4*>(Yt + i * ALD + 64 + tx * 4);
        float p[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            acc[r][0] = fmaf(p[r], y0.x, acc[r][0]); acc[r][1] = fmaf(p[r], y0.y, acc[r][1]);
            acc[r][2] = fmaf(p[r], y0.z, acc[r][2]); acc[r][3] = fmaf(p[r], y0.w, acc[r][3]);
            acc[r][4] = fmaf(p[r], y1.x, acc[r][4]); acc[r][5] = fmaf(p[r], y1.y, acc[r][5]);
            acc[r][6] = fmaf(p[r], y1.z, acc[r][6]); acc[r][7] = fmaf(p[r], y1.w, acc[r][7]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) attn_fwd_kernel(float* __restrict__ o, float* __restrict__ lse, const float* __restrict__ q,
                                                          const float* __restrict__ k, const float* __restrict__ v, int ld,
                                                          const uint8_t* __restrict__ key_mask, int64_t mask_stride, int T) {
    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;
    float* Ks = Qs + AT * ALD;
    float* Vs = Ks + AT * ALD;
    float* Ps = Ks;   // P overwrites the K tile once the scores are in registers

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int b = blockIdx.y, q0 = blockIdx.x * AT;
    const int64_t base = (int64_t)b * T;
    const int64_t mask_base = (int64_t)b * mask_stride;

    load_tile(Qs, q, ld, base + q0, q0, T);

    float m_run[4], l_part[4], oacc[4][8];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        m_run[ii] = -INFINITY;
        l_part[ii] = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) oacc[ii][c] = 0.f;
    }

    for (int k0 = 0; k0 < T; k0 += AT) {
        __syncthreads();   // previous tile's P / V fully consumed (and Q visible on the first pass)
        load_tile(Ks, k, ld, base + k0, k0, T);
        load_tile(Vs, v, ld, base + k0, k0, T);
        __syncthreads();

        float s[4][4];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) s[ii][jj] = 0.f;
        tile_dot(s, Qs, Ks, tx, ty);

        bool masked[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            int j = k0 + tx + 16 * jj;
            masked[jj] = (j >= T) || (key_mask && key_mask[mask_base + j]);
        }
        float alpha[4];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            float mx = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                s[ii][jj] = masked[jj] ? -INFINITY : s[ii][jj] * (SM_SCALE * LOG2E);
                mx = fmaxf(mx, s[ii][jj]);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            float m_new = fmaxf(m_run[ii], mx);
            alpha[ii] = (m_new == -INFINITY) ? 1.0f : exp2f(m_run[ii] - m_new);
            float rowsum = 0.f;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float p = (m_new == -INFINITY) ? 0.f : exp2f(s[ii][jj] - m_new);
                s[ii][jj] = p;
                rowsum += p;
            }
            l_part[ii] = l_part[ii] * alpha[ii] + rowsum;
            m_run[ii] = m_new;
        }
        __syncthreads();   // all score reads of Ks done before P overwrites it
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) Ps[(ty + 16 * ii) * PLD + tx + 16 * jj] = s[ii][jj];
        __syncthreads();

        // O rows are ty+16*ii here (same rows as the softmax state); accumulate P.V
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int c = 0; c < 8; ++c) oacc[ii][c] *= alpha[ii];
#pragma unroll 2
        for (int j = 0; j < AT; j += 4) {
            float4 pv[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) pv[ii] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * ii) * PLD + j);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float4 y0 = *reinterpret_cast<const float4*>(Vs + (j + jj) * ALD + tx * 4);
                float4 y1 = *reinterpret_cast<const float4*>(Vs + (j + jj) * ALD + 64 + tx * 4);
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    float p = jj == 0 ? pv[ii].x : jj == 1 ? pv[ii].y : jj == 2 ? pv[ii].z : pv[ii].w;
                    oacc[ii][0] = fmaf(p, y0.x, oacc[ii][0]); oacc[ii][1] = fmaf(p, y0.y, oacc[ii][1]);
                    oacc[ii][2] = fmaf(p, y0.z, oacc[ii][2]); oacc[ii][3] = fmaf(p, y0.w, oacc[ii][3]);
                    oacc[ii][4] = fmaf(p, y1.x, oacc[ii][4]); oacc[ii][5] = fmaf(p, y1.y, oacc[ii][5]);
                    oacc[ii][6] = fmaf(p, y1.z, oacc[ii][6]); oacc[ii][7] = fmaf(p, y1.w, oacc[ii][7]);
                }
            }
        }
    }

#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        float l = l_part[ii];
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
        int t = q0 + ty + 16 * ii;
        if (t < T) {
            float inv = 1.0f / l;
            float* orow = o + (base + t) * D;
            *reinterpret_cast<float4*>(orow + tx * 4) = make_float4(oacc[ii][0] * inv, oacc[ii][1] * inv, oacc[ii][2] * inv, oacc[ii][3] * inv);
            *reinterpret_cast<float4*>(orow + 64 + tx * 4) = make_float4(oacc[ii][4] * inv, oacc[ii][5] * inv, oacc[ii][6] * inv, oacc[ii][7] * inv);
            if (lse && tx == 0) lse[base + t] = m_run[ii] * LN2 + logf(l);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_delta_kernel(float* __restrict__ delta, const float* __restrict__ o,
                                                         const float* __restrict__ d_o, int64_t rows) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        float4 a = reinterpret_cast<const float4*>(o + r * D)[lane];
        float4 g = reinterpret_cast<const float4*>(d_o + r * D)[lane];
        float s = warp_sum(a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w);
        if (lane == 0) delta[r] = s;
    }
}

// shared phase 1 of both backward kernels: for the (query tile, key tile) pair held in shared
// memory compute P and dS = P * (dP - delta) / sqrt(d) into registers.
__device__ __forceinline__ void bwd_scores(float (&p)[4][4], float (&ds)[4][4], const float* Qs, const float* Ks, const float* dOs,
                                           const float* Vs, const float* lse_s, const float* delta_s, const uint8_t* key_mask,
                                           int64_t mask_base, int q0, int k0, int T, int tx, int ty) {
    float dp[4][4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) { p[ii][jj] = 0.f; dp[ii][jj] = 0.f; }
    tile_dot(p, Qs, Ks, tx, ty);
    tile_dot(dp, dOs, Vs, tx, ty);
    bool masked[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        int j = k0 + tx + 16 * jj;
        masked[jj] = (j >= T) || (key_mask && key_mask[mask_base + j]);
    }
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        int i = ty + 16 * ii;
        bool row_ok = (q0 + i) < T;
        float l2 = lse_s[i] * LOG2E, dl = delta_s[i];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            float pv = (masked[jj] || !row_ok) ? 0.f : exp2f(p[ii][jj] * (SM_SCALE * LOG2E) - l2);
            p[ii][jj] = pv;
            ds[ii][jj] = pv * (dp[ii][jj] - dl) * SM_SCALE;
        }
    }
}

// key-stationary: dK, dV for one tile of 64 keys, looping over all query tiles
__global__ void __launch_bounds__(256, 1) attn_bwd_dkdv_kernel(float* __restrict__ dk, float* __restrict__ dv, int ld_d,
                                                               const float* __restrict__ q, const float* __restrict__ k,
                                                               const float* __restrict__ v, int ld, const float* __restrict__ d_o,
                                                               const float* __restrict__ lse, const float* __restrict__ delta,
                                                               const uint8_t* __restrict__ key_mask, int64_t mask_stride, int T) {
    extern __shared__ __align__(16) float smem[];
    float* Ks = smem;
    float* Vs = Ks + AT * ALD;
    float* Qs = Vs + AT * ALD;
    float* dOs = Qs + AT * ALD;
    float* Ps = dOs + AT * ALD;
    float* dSs = Ps + AT * PLD;
    float* lse_s = dSs + AT * PLD;
    float* delta_s = lse_s + AT;

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int b = blockIdx.y, k0 = blockIdx.x * AT;
    const int64_t base = (int64_t)b * T;

    load_tile(Ks, k, ld, base + k0, k0, T);
    load_tile(Vs, v, ld, base + k0, k0, T);

    float dkacc[4][8], dvacc[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) { dkacc[r][c] = 0.f; dvacc[r][c] = 0.f; }

    for (int q0 = 0; q0 < T; q0 += AT) {
        __syncthreads();
        load_tile(Qs, q, ld, base + q0, q0, T);
        load_tile(dOs, d_o, D, base + q0, q0, T);
        if (tid < AT) {
            bool ok = q0 + tid < T;
            lse_s[tid] = ok ? lse[base + q0 + tid] : 0.f;
            delta_s[tid] = ok ? delta[base + q0 + tid] : 0.f;
        }
        __syncthreads();
        float p[4][4], ds[4][4];
        bwd_scores(p, ds, Qs, Ks, dOs, Vs, lse_s, delta_s, key_mask, (int64_t)b * mask_stride, q0, k0, T, tx, ty);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                Ps[(ty + 16 * ii) * PLD + tx + 16 * jj] = p[ii][jj];
                dSs[(ty + 16 * ii) * PLD + tx + 16 * jj] = ds[ii][jj];
            }
        __syncthreads();
        tile_ptv_cols(dvacc, Ps, dOs, tx, ty);    // dV += P^T dO
        tile_ptv_cols(dkacc, dSs, Qs, tx, ty);    // dK += dS^T Q
    }

#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int t = k0 + ty * 4 + r;
        if (t < T) {
            float* kr = dk + (base + t) * (int64_t)ld_d;
            float* vr = dv + (base + t) * (int64_t)ld_d;
            *reinterpret_cast<float4*>(kr + tx * 4) = make_float4(dkacc[r][0], dkacc[r][1], dkacc[r][2], dkacc[r][3]);
            *reinterpret_cast<float4*>(kr + 64 + tx * 4) = make_float4(dkacc[r][4], dkacc[r][5], dkacc[r][6], dkacc[r][7]);
            *reinterpret_cast<float4*>(vr + tx * 4) = make_float4(dvacc[r][0], dvacc[r][1], dvacc[r][2], dvacc[r][3]);
            *reinterpret_cast<float4*>(vr + 64 + tx * 4) = make_float4(dvacc[r][4], dvacc[r][5], dvacc[r][6], dvacc[r][7]);
        }
    }
}

// query-stationary: dQ for one tile of 64 queries, looping over all key tiles
__global__ void __launch_bounds__(256, 1) attn_bwd_dq_kernel(float* __restrict__ dq, int ld_d, const float* __restrict__ q,
                                                             const float* __restrict__ k, const float* __restrict__ v, int ld,
                                                             const float* __restrict__ d_o, const float* __restrict__ lse,
                                                             const float* __restrict__ delta, const uint8_t* __restrict__ key_mask, int64_t mask_stride, int T) {
    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;
    float* dOs = Qs + AT * ALD;
    float* Ks = dOs + AT * ALD;
    float* Vs = Ks + AT * ALD;
    float* dSs = Vs + AT * ALD;
    float* lse_s = dSs + AT * PLD;
    float* delta_s = lse_s + AT;

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int b = blockIdx.y, q0 = blockIdx.x * AT;
    const int64_t base = (int64_t)b * T;

    load_tile(Qs, q, ld, base + q0, q0, T);
    load_tile(dOs, d_o, D, base + q0, q0, T);
    if (tid < AT) {
        bool ok = q0 + tid < T;
        lse_s[tid] = ok ? lse[base + q0 + tid] : 0.f;
        delta_s[tid] = ok ? delta[base + q0 + tid] : 0.f;
    }

    float dqacc[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) dqacc[r][c] = 0.f;

    for (int k0 = 0; k0 < T; k0 += AT) {
        __syncthreads();
        load_tile(Ks, k, ld, base + k0, k0, T);
        load_tile(Vs, v, ld, base + k0, k0, T);
        __syncthreads();
        float p[4][4], ds[4][4];
        bwd_scores(p, ds, Qs, Ks, dOs, Vs, lse_s, delta_s, key_mask, (int64_t)b * mask_stride, q0, k0, T, tx, ty);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) dSs[(ty + 16 * ii) * PLD + tx + 16 * jj] = ds[ii][jj];
        __syncthreads();
        tile_pv_rows(dqacc, dSs, Ks, tx, ty);     // dQ += dS K
    }

#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int t = q0 + ty * 4 + r;
        if (t < T) {
            float* qr = dq + (base + t) * (int64_t)ld_d;
            *reinterpret_cast<float4*>(qr + tx * 4) = make_float4(dqacc[r][0], dqacc[r][1], dqacc[r][2], dqacc[r][3]);
            *reinterpret_cast<float4*>(qr + 64 + tx * 4) = make_float4(dqacc[r][4], dqacc[r][5], dqacc[r][6], dqacc[r][7]);
        }
    }
}

constexpr size_t FWD_SMEM = (size_t)3 * AT * ALD * sizeof(float);
constexpr size_t DKDV_SMEM = ((size_t)4 * AT * ALD + 2 * AT * PLD + 2 * AT) * sizeof(float);
constexpr size_t DQ_SMEM = ((size_t)4 * AT * ALD + AT * PLD + 2 * AT) * sizeof(float);

static bool g_attr_set = false;
static int set_attrs() {
    if (g_attr_set) return 0;
    DIP_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
    DIP_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DKDV_SMEM));
    DIP_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DQ_SMEM));
    g_attr_set = true;
    return 0;
}

int sm_count();

}  // namespace dip

using namespace dip;

extern "C" {

int dip_attn_fwd(float* o, float* lse, const float* q, const float* k, const float* v, int ld, const uint8_t* key_mask, int64_t mask_stride, int B, int T,
                 dip_stream_t stream) {
    DIP_REQUIRE(o && q && k && v && B > 0 && T > 0, "dip_attn_fwd: null/empty argument");
    DIP_REQUIRE(ld % 4 == 0 && ld >= D, "dip_attn_fwd: ld must be a multiple of 4 and >= 128");
    DIP_REQUIRE(B <= 65535, "dip_attn_fwd: B too large for one launch");
    DIP_TRY(set_attrs());
    dim3 grid(cdiv(T, AT), B);
    DIP_PROF(DIP_PROF_ATTN_FWD, stream);
    attn_fwd_kernel<<<grid, 256, FWD_SMEM, as_stream(stream)>>>(o, lse, q, k, v, ld, key_mask, mask_stride, T);
    DIP_LAUNCHED();
    return 0;
}

int dip_attn_bwd(float* dq, float* dk, float* dv, int ld_d, const float* q, const float* k, const float* v, int ld, const float* o,
                 const float* d_o, const float* lse, const uint8_t* key_mask, int64_t mask_stride, float* delta_ws, int B, int T,
                 dip_stream_t stream) {
    DIP_REQUIRE(dq && dk && dv && q && k && v && o && d_o && lse && delta_ws && B > 0 && T > 0, "dip_attn_bwd: null/empty argument");
    DIP_REQUIRE(ld % 4 == 0 && ld_d % 4 == 0, "dip_attn_bwd: ld must be a multiple of 4");
    DIP_REQUIRE(B <= 65535, "dip_attn_bwd: B too large for one launch");
    DIP_TRY(set_attrs());
    cudaStream_t st = as_stream(stream);
    int64_t rows = (int64_t)B * T;
    int64_t blocks = (rows + 7) / 8, cap = (int64_t)sm_count() * 8;
    {
        DIP_PROF(DIP_PROF_OTHER, stream);
        attn_delta_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(delta_ws, o, d_o, rows);
        DIP_LAUNCHED();
    }
    dim3 grid(cdiv(T, AT), B);
    {
        DIP_PROF(DIP_PROF_ATTN_BWD_DKDV, stream);
        attn_bwd_dkdv_kernel<<<grid, 256, DKDV_SMEM, st>>>(dk, dv, ld_d, q, k, v, ld, d_o, lse, delta_ws, key_mask, mask_stride, T);
        DIP_LAUNCHED();
    }
    {
        DIP_PROF(DIP_PROF_ATTN_BWD_DQ, stream);
        attn_bwd_dq_kernel<<<grid, 256, DQ_SMEM, st>>>(dq, ld_d, q, k, v, ld, d_o, lse, delta_ws, key_mask, mask_stride, T);
        DIP_LAUNCHED();
    }
    return 0;
}

}  // extern "C"
