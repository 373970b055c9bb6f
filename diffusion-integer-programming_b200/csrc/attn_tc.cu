// Flash attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), one head of width 128,
// key-padding mask, "bf16x2 split" numerics (see tc_common.cuh): forward here, backward below.
//
// Operands are the bf16 (hi, lo) planes of the packed in_proj output, shape (B*T, 384) each, written
// by the producing GEMM's epilogue.  Every tile is moved by TMA (SWIZZLE_128B boxes of 64 columns)
// straight into the layout tcgen05.mma consumes; the same [rows x 128] tile serves as a K-major
// operand (Q, K in S = Q K^T) or as an MN-major B operand (V in P V) by descriptor only.
//
// Forward CTA = 128 queries (one TMEM lane / one softmax thread per query), key tiles of 64:
//   warp 4 : TMA producer (Q once, then a 2-stage K/V ring)
//   warp 5 : one thread issues tcgen05.mma:  S_j = Q K_j^T (3 passes) into a double-buffered TMEM
//            tile, PV_j = P_j V_j (3 passes) into a second TMEM tile; S_{j+1} is queued before PV_j so the
//            tensor pipe works while the softmax warps turn S_j into P_j
//   warps 0-3 : online softmax in registers (exp2 domain), P_j written as bf16 hi/lo into swizzled
//            shared memory, running output kept in registers and folded with each PV_j tile.
#include <math.h>
#include "tc_common.cuh"

namespace dip {
namespace tcattn {

using namespace tc;

constexpr int BM = 128, BN = 64;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float SM_SCALE = 0.08838834764831845f;

constexpr uint32_t Q_BYTES = 65536;        // hi [2][128][64] + lo [2][128][64] bf16
constexpr uint32_t KV_STAGE = 65536;       // K hi, K lo, V hi, V lo: [2][64][64] bf16 each
constexpr uint32_t P_BYTES = 32768;        // hi [128][64] + lo [128][64]
constexpr uint32_t OFF_KV = Q_BYTES, OFF_P = Q_BYTES + 2 * KV_STAGE, OFF_BAR = OFF_P + P_BYTES;
constexpr uint32_t FWD_SMEM = OFF_BAR + 256 + 1024;   // + barriers + alignment slack

constexpr uint32_t IDESC_S = instr_desc(128, BN, 0, 0);     // Q (K-major) x K (K-major) -> 128 x 64
constexpr uint32_t IDESC_PV = instr_desc(128, 128, 0, 1);   // P (K-major) x V (MN-major) -> 128 x 128

enum { B_QFULL = 0, B_KVFULL = 1, B_KVEMPTY = 3, B_SFULL = 5, B_SFREE = 7, B_PFULL = 9, B_PVFULL = 10, B_COUNT = 11 };

__global__ void __launch_bounds__(192, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_hi128, const __grid_constant__ CUtensorMap tm_lo128,
                   const __grid_constant__ CUtensorMap tm_hi64, const __grid_constant__ CUtensorMap tm_lo64, float* __restrict__ o,
                   float* __restrict__ lse, const uint8_t* __restrict__ key_mask, int64_t mask_stride, int T) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = sm;
    uint8_t* sKV = sm + OFF_KV;
    uint8_t* sP = sm + OFF_P;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, q0 = blockIdx.x * BM;
    const int row_base = b * T;
    const int n_tiles = (T + BN - 1) / BN;

    if (threadIdx.x == 0) {
        mbar_init(&bars[B_QFULL], 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars[B_KVFULL + s], 1);
            mbar_init(&bars[B_KVEMPTY + s], 1);
            mbar_init(&bars[B_SFULL + s], 1);
            mbar_init(&bars[B_SFREE + s], 128);
        }
        mbar_init(&bars[B_PFULL], 128);
        mbar_init(&bars[B_PVFULL], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;      // columns [0,64) S buf 0, [64,128) S buf 1, [128,256) PV

    if (warp == 4) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            tma_prefetch_desc(&tm_hi128); tma_prefetch_desc(&tm_lo128); tma_prefetch_desc(&tm_hi64); tma_prefetch_desc(&tm_lo64);
            mbar_expect_tx(&bars[B_QFULL], Q_BYTES);
            for (int kb = 0; kb < 2; ++kb) {
                tma_load_2d(sQ + kb * 16384, &tm_hi128, &bars[B_QFULL], kb * 64, row_base + q0);
                tma_load_2d(sQ + 32768 + kb * 16384, &tm_lo128, &bars[B_QFULL], kb * 64, row_base + q0);
            }
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j & 1;
                if (j >= 2) mbar_wait(&bars[B_KVEMPTY + s], ((j >> 1) - 1) & 1);
                mbar_expect_tx(&bars[B_KVFULL + s], KV_STAGE);
                uint8_t* dst = sKV + s * KV_STAGE;
                const int krow = row_base + j * BN;
                for (int kb = 0; kb < 2; ++kb) {
                    tma_load_2d(dst + kb * 8192, &tm_hi64, &bars[B_KVFULL + s], 128 + kb * 64, krow);            // K hi
                    tma_load_2d(dst + 16384 + kb * 8192, &tm_lo64, &bars[B_KVFULL + s], 128 + kb * 64, krow);    // K lo
                    tma_load_2d(dst + 32768 + kb * 8192, &tm_hi64, &bars[B_KVFULL + s], 256 + kb * 64, krow);    // V hi
                    tma_load_2d(dst + 49152 + kb * 8192, &tm_lo64, &bars[B_KVFULL + s], 256 + kb * 64, krow);    // V lo
                }
            }
        }
    } else if (warp == 5) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t aQ = smem_u32(sQ), aKV = smem_u32(sKV), aP = smem_u32(sP);
            mbar_wait(&bars[B_QFULL], 0);
            auto issue_S = [&](int j) {
                const int s = j & 1;
                mbar_wait(&bars[B_KVFULL + s], (j >> 1) & 1);
                if (j >= 2) mbar_wait(&bars[B_SFREE + s], ((j >> 1) - 1) & 1);
                tc_fence_after();
                const uint32_t kbase = aKV + s * KV_STAGE;
                uint32_t acc = 0;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t qa = aQ + (pass == 2 ? 32768 : 0);            // hi, hi, lo
                    const uint32_t ka = kbase + (pass == 1 ? 16384 : 0);        // hi, lo, hi
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            tc_mma(tmem + s * BN, smem_desc(qa + kb * 16384 + ks * 32, 16, 1024), smem_desc(ka + kb * 8192 + ks * 32, 16, 1024),
                                   IDESC_S, acc);
                            acc = 1;
                        }
                }
                tc_commit(&bars[B_SFULL + s]);
            };
            auto issue_PV = [&](int j) {
                const int s = j & 1;
                mbar_wait(&bars[B_PFULL], j & 1);
                tc_fence_after();
                const uint32_t vbase = aKV + s * KV_STAGE + 32768;
                uint32_t acc = 0;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t pa = aP + (pass == 2 ? 16384 : 0);           // hi, hi, lo
                    const uint32_t va = vbase + (pass == 1 ? 16384 : 0);        // hi, lo, hi
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        // V tile: two [64 keys x 64 d] blocks 8192 bytes apart (MN blocks), 16 keys per k-step = 2048 bytes
                        tc_mma(tmem + 128, smem_desc(pa + ks * 32, 16, 1024), smem_desc(va + ks * 2048, 8192, 1024), IDESC_PV, acc);
                        acc = 1;
                    }
                }
                tc_commit(&bars[B_PVFULL]);
                tc_commit(&bars[B_KVEMPTY + s]);
            };
            issue_S(0);
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) issue_S(j + 1);
                issue_PV(j);
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax: thread = query row
        const int row = threadIdx.x;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        const uint8_t* mrow = key_mask ? key_mask + (int64_t)b * mask_stride : nullptr;
        float oacc[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) oacc[i] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            const int s = j & 1;
            mbar_wait(&bars[B_SFULL + s], (j >> 1) & 1);
            tc_fence_after();
            float sv[64];
            {
                float t0[32], t1[32];
                tmem_ld32(lane_addr + s * BN, t0);
                tmem_ld32(lane_addr + s * BN + 32, t1);
#pragma unroll
                for (int i = 0; i < 32; ++i) { sv[i] = t0[i]; sv[32 + i] = t1[i]; }
            }
            tc_fence_before();
            mbar_arrive(&bars[B_SFREE + s]);
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 64; ++c) {
                const int key = j * BN + c;
                const bool masked = (key >= T) || (mrow && mrow[key]);
                sv[c] = masked ? -INFINITY : sv[c] * (SM_SCALE * LOG2E);
                mx = fmaxf(mx, sv[c]);
            }
            const float m_new = fmaxf(m_run, mx);
            const bool dead = (m_new == -INFINITY);
            const float alpha = dead ? 1.0f : exp2f(m_run - m_new);
            float rowsum = 0.f;
#pragma unroll
            for (int c = 0; c < 64; ++c) {
                const float p = dead ? 0.f : exp2f(sv[c] - m_new);
                sv[c] = p;
                rowsum += p;
            }
            l_run = l_run * alpha + rowsum;
            m_run = m_new;
            // previous tile's P.V must be complete before P is overwritten; fold it into the running output
            if (j > 0) {
                mbar_wait(&bars[B_PVFULL], (j - 1) & 1);
                tc_fence_after();
            }
            // P_j -> shared memory (bf16 hi / lo, K-major SWIZZLE_128B: row = query, 8 chunks of 8 keys)
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    __nv_bfloat16 h0, l0, h1, l1;
                    split_bf16(sv[c8 * 8 + 2 * e], h0, l0);
                    split_bf16(sv[c8 * 8 + 2 * e + 1], h1, l1);
                    hw[e] = pack2(h0, h1);
                    lw[e] = pack2(l0, l1);
                }
                const uint32_t off = sw128_off(row, c8);
                *reinterpret_cast<uint4*>(sP + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(sP + 16384 + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
            if (j > 0) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    float t[32];
                    tmem_ld32(lane_addr + 128 + ch * 32, t);
#pragma unroll
                    for (int i = 0; i < 32; ++i) oacc[ch * 32 + i] = (oacc[ch * 32 + i] + t[i]) * alpha;
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars[B_PFULL]);
        }
        mbar_wait(&bars[B_PVFULL], (n_tiles - 1) & 1);
        tc_fence_after();
        const float inv = 1.0f / l_run;
        const int t_q = q0 + row;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            float t[32];
            tmem_ld32(lane_addr + 128 + ch * 32, t);
            if (t_q < T) {
                float* orow = o + ((int64_t)row_base + t_q) * D + ch * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(orow + i) = make_float4((oacc[ch * 32 + i] + t[i]) * inv, (oacc[ch * 32 + i + 1] + t[i + 1]) * inv,
                                                                       (oacc[ch * 32 + i + 2] + t[i + 2]) * inv, (oacc[ch * 32 + i + 3] + t[i + 3]) * inv);
            }
        }
        if (lse && t_q < T) lse[(int64_t)row_base + t_q] = m_run * LN2 + logf(l_run);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace tcattn

// ---------------------------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows,
                      uint32_t box_cols) {
    EncodeTiledFn fn = encode_fn();
    DIP_REQUIRE(fn, "cuTensorMapEncodeTiled entry point not available");
    DIP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (row_stride_elems * 2) % 16 == 0, "tensor map: base / row stride must be 16-byte aligned");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DIP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return 0;
}

}  // namespace dip

using namespace dip;

extern "C" {

int dip_attn_fwd_tc(float* o, float* lse, const void* qkv_hi, const void* qkv_lo, const uint8_t* key_mask, int64_t mask_stride, int B, int T,
                    dip_stream_t stream) {
    DIP_REQUIRE(o && qkv_hi && qkv_lo && B > 0 && T > 0, "dip_attn_fwd_tc: null/empty argument");
    DIP_REQUIRE(B <= 65535, "dip_attn_fwd_tc: B too large for one launch");
    static bool attr = false;
    if (!attr) {
        DIP_CUDA(cudaFuncSetAttribute(tcattn::attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcattn::FWD_SMEM));
        attr = true;
    }
    const uint64_t R = (uint64_t)B * T;
    CUtensorMap hi128, lo128, hi64, lo64;
    DIP_TRY(make_tmap_bf16_2d(&hi128, qkv_hi, R, 3 * D, 3 * D, 128, 64));
    DIP_TRY(make_tmap_bf16_2d(&lo128, qkv_lo, R, 3 * D, 3 * D, 128, 64));
    DIP_TRY(make_tmap_bf16_2d(&hi64, qkv_hi, R, 3 * D, 3 * D, 64, 64));
    DIP_TRY(make_tmap_bf16_2d(&lo64, qkv_lo, R, 3 * D, 3 * D, 64, 64));
    dim3 grid(cdiv(T, tcattn::BM), B);
    DIP_PROF(DIP_PROF_ATTN_FWD, stream);
    tcattn::attn_fwd_tc_kernel<<<grid, 192, tcattn::FWD_SMEM, as_stream(stream)>>>(hi128, lo128, hi64, lo64, o, lse, key_mask, mask_stride, T);
    DIP_LAUNCHED();
    return 0;
}

}  // extern "C"
