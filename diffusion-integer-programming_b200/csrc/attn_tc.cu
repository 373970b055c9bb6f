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


// =============================================================================================
// backward (input gradients only, deterministic: no atomics)
// =============================================================================================
// Two launches of one templated kernel, both with a 128-row stationary operand pair in shared memory
// and 32-row streamed tiles (2-stage TMA ring):
//   KS = true  (key-stationary)   rows = keys   : X = K, Y = V stationary; U = Q, W = dO streamed
//        G1 = K Q^T (= S^T), G2 = V dO^T (= dP^T);  dV += P^T dO,  dK += dS^T Q
//   KS = false (query-stationary) rows = queries: X = Q, Y = dO stationary; U = K, W = V streamed
//        G1 = Q K^T (= S),   G2 = dO V^T (= dP);    dQ += dS K
// P = exp(S/sqrt(d) - lse[query]) and dS = P (dP - delta[query]) / sqrt(d) are formed by the 128 row
// threads from TMEM and written back to shared memory as bf16 hi/lo in the un-swizzled ("interleaved")
// K-major core-matrix layout; the streamed tiles double as MN-major B operands of the accumulating GEMMs.
constexpr int BT = 32;                         // streamed tile rows
constexpr uint32_t ST_PLANE = 32768;           // stationary plane: [2][128][64] bf16
constexpr uint32_t TL_PLANE = 8192;            // streamed plane:   [2][32][64] bf16
constexpr uint32_t TL_STAGE = 4 * TL_PLANE;    // U hi, U lo, W hi, W lo
constexpr uint32_t PD_PLANE = 8192;            // [128][32] bf16 interleaved
constexpr uint32_t BOFF_X = 0, BOFF_Y = 2 * ST_PLANE, BOFF_TL = 4 * ST_PLANE, BOFF_PD = BOFF_TL + 2 * TL_STAGE,
                   BOFF_BAR = BOFF_PD + 4 * PD_PLANE;
constexpr uint32_t BWD_SMEM = BOFF_BAR + 256 + 1024;
constexpr uint32_t IDESC_G = instr_desc(128, BT, 0, 0);       // X (K-major) x U (K-major) -> 128 x 32
constexpr uint32_t IDESC_ACC = instr_desc(128, 128, 0, 1);    // P/dS (K-major) x U/W (MN-major) -> 128 x 128

enum { C_XFULL = 0, C_TLFULL = 1, C_TLEMPTY = 3, C_GFULL = 5, C_GFREE = 7, C_PDFULL = 9, C_PDFREE = 10, C_DONE = 11 };

// un-swizzled K-major descriptor for the [128 x 32] P / dS tiles: core matrix = 8 rows x 16 bytes,
// K-adjacent core matrices 128 bytes apart (LBO), 8-row groups 512 bytes apart (SBO)
__device__ __forceinline__ uint64_t nosw_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(128u >> 4) << 16;
    d |= (uint64_t)(512u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

template <bool KS>
__global__ void __launch_bounds__(192, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv_hi128, const __grid_constant__ CUtensorMap tm_qkv_lo128,
                   const __grid_constant__ CUtensorMap tm_qkv_hi32, const __grid_constant__ CUtensorMap tm_qkv_lo32,
                   const __grid_constant__ CUtensorMap tm_do_hi128, const __grid_constant__ CUtensorMap tm_do_lo128,
                   const __grid_constant__ CUtensorMap tm_do_hi32, const __grid_constant__ CUtensorMap tm_do_lo32,
                   float* __restrict__ out1, float* __restrict__ out2, int ld_d, const float* __restrict__ lse, const float* __restrict__ delta,
                   const uint8_t* __restrict__ key_mask, int64_t mask_stride, int T) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + BOFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, r0 = blockIdx.x * 128;      // first stationary row (key for KS, query otherwise)
    const int row_base = b * T;
    const int n_tiles = (T + BT - 1) / BT;

    if (threadIdx.x == 0) {
        mbar_init(&bars[C_XFULL], 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars[C_TLFULL + s], 1);
            mbar_init(&bars[C_TLEMPTY + s], 1);
            mbar_init(&bars[C_GFULL + s], 1);
            mbar_init(&bars[C_GFREE + s], 128);
        }
        mbar_init(&bars[C_PDFULL], 128);
        mbar_init(&bars[C_PDFREE], 1);
        mbar_init(&bars[C_DONE], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;     // [0,64) G1 x2, [64,128) G2 x2, [128,256) acc1 (dV), [256,384) acc2 (dK / dQ)

    // column of the packed (q|k|v) planes each operand lives in
    constexpr int COL_X = KS ? 128 : 0;      // K : Q
    constexpr int COL_U = KS ? 0 : 128;      // Q : K

    if (warp == 4) {
        if (lane == 0) {
            const CUtensorMap* yhi = KS ? &tm_qkv_hi128 : &tm_do_hi128;
            const CUtensorMap* ylo = KS ? &tm_qkv_lo128 : &tm_do_lo128;
            const int col_y = KS ? 256 : 0;  // V : dO
            mbar_expect_tx(&bars[C_XFULL], 4 * ST_PLANE);
            for (int kb = 0; kb < 2; ++kb) {
                tma_load_2d(sm + BOFF_X + kb * 16384, &tm_qkv_hi128, &bars[C_XFULL], COL_X + kb * 64, row_base + r0);
                tma_load_2d(sm + BOFF_X + ST_PLANE + kb * 16384, &tm_qkv_lo128, &bars[C_XFULL], COL_X + kb * 64, row_base + r0);
                tma_load_2d(sm + BOFF_Y + kb * 16384, yhi, &bars[C_XFULL], col_y + kb * 64, row_base + r0);
                tma_load_2d(sm + BOFF_Y + ST_PLANE + kb * 16384, ylo, &bars[C_XFULL], col_y + kb * 64, row_base + r0);
            }
            const CUtensorMap* whi = KS ? &tm_do_hi32 : &tm_qkv_hi32;
            const CUtensorMap* wlo = KS ? &tm_do_lo32 : &tm_qkv_lo32;
            const int col_w = KS ? 0 : 256;  // dO : V
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j & 1;
                if (j >= 2) mbar_wait(&bars[C_TLEMPTY + s], ((j >> 1) - 1) & 1);
                mbar_expect_tx(&bars[C_TLFULL + s], TL_STAGE);
                uint8_t* dst = sm + BOFF_TL + s * TL_STAGE;
                const int trow = row_base + j * BT;
                for (int kb = 0; kb < 2; ++kb) {
                    tma_load_2d(dst + kb * 4096, &tm_qkv_hi32, &bars[C_TLFULL + s], COL_U + kb * 64, trow);
                    tma_load_2d(dst + TL_PLANE + kb * 4096, &tm_qkv_lo32, &bars[C_TLFULL + s], COL_U + kb * 64, trow);
                    tma_load_2d(dst + 2 * TL_PLANE + kb * 4096, whi, &bars[C_TLFULL + s], col_w + kb * 64, trow);
                    tma_load_2d(dst + 3 * TL_PLANE + kb * 4096, wlo, &bars[C_TLFULL + s], col_w + kb * 64, trow);
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            const uint32_t aX = smem_u32(sm + BOFF_X), aY = smem_u32(sm + BOFF_Y), aTL = smem_u32(sm + BOFF_TL), aPD = smem_u32(sm + BOFF_PD);
            mbar_wait(&bars[C_XFULL], 0);
            auto issue_G = [&](int j) {
                const int s = j & 1;
                mbar_wait(&bars[C_TLFULL + s], (j >> 1) & 1);
                if (j >= 2) mbar_wait(&bars[C_GFREE + s], ((j >> 1) - 1) & 1);
                tc_fence_after();
                const uint32_t tl = aTL + s * TL_STAGE;
#pragma unroll
                for (int g = 0; g < 2; ++g) {                       // g = 0: X.U^T, g = 1: Y.W^T
                    const uint32_t st = g == 0 ? aX : aY;
                    const uint32_t tp = tl + g * 2 * TL_PLANE;
                    uint32_t acc = 0;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t a = st + (pass == 2 ? ST_PLANE : 0);          // hi, hi, lo
                        const uint32_t bb = tp + (pass == 1 ? TL_PLANE : 0);         // hi, lo, hi
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                tc_mma(tmem + g * 64 + s * BT, smem_desc(a + kb * 16384 + ks * 32, 16, 1024),
                                       smem_desc(bb + kb * 4096 + ks * 32, 16, 1024), IDESC_G, acc);
                                acc = 1;
                            }
                    }
                }
                tc_commit(&bars[C_GFULL + s]);
            };
            auto issue_acc = [&](int j) {
                const int s = j & 1;
                mbar_wait(&bars[C_PDFULL], j & 1);
                tc_fence_after();
                const uint32_t tl = aTL + s * TL_STAGE;
#pragma unroll
                for (int g = (KS ? 0 : 1); g < 2; ++g) {            // g = 0: acc1 += P.W (dV), g = 1: acc2 += dS.U (dK / dQ)
                    const uint32_t pd = aPD + g * 2 * PD_PLANE;
                    const uint32_t tp = tl + (g == 0 ? 2 * TL_PLANE : 0);
                    uint32_t acc = j > 0 ? 1u : 0u;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t a = pd + (pass == 2 ? PD_PLANE : 0);
                        const uint32_t bb = tp + (pass == 1 ? TL_PLANE : 0);
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            // streamed tile as MN-major B: d-halves 4096 bytes apart, 16 rows per k-step = 2048 bytes
                            tc_mma(tmem + 128 + g * 128, nosw_desc(a + ks * 256), smem_desc(bb + ks * 2048, 4096, 1024), IDESC_ACC, acc);
                            acc = 1;
                        }
                    }
                }
                tc_commit(&bars[C_PDFREE]);
                tc_commit(&bars[C_TLEMPTY + s]);
            };
            issue_G(0);
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) issue_G(j + 1);
                issue_acc(j);
            }
            tc_commit(&bars[C_DONE]);
        }
    } else {
        const int row = threadIdx.x;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        const uint8_t* mrow = key_mask ? key_mask + (int64_t)b * mask_stride : nullptr;
        const int r_idx = r0 + row;
        bool row_ok = r_idx < T;
        float row_lse2 = 0.f, row_delta = 0.f;
        if (KS) {
            if (row_ok && mrow && mrow[r_idx]) row_ok = false;      // masked key: its column of P is zero
        } else if (row_ok) {
            row_lse2 = lse[(int64_t)row_base + r_idx] * LOG2E;
            row_delta = delta[(int64_t)row_base + r_idx];
        }
        uint8_t* sPD = sm + BOFF_PD;
        const uint32_t wr_off = (uint32_t)(row >> 3) * 512u + (uint32_t)(row & 7) * 16u;
        for (int j = 0; j < n_tiles; ++j) {
            const int s = j & 1;
            mbar_wait(&bars[C_GFULL + s], (j >> 1) & 1);
            tc_fence_after();
            float g1[32], g2[32];
            tmem_ld32(lane_addr + s * BT, g1);
            tmem_ld32(lane_addr + 64 + s * BT, g2);
            tc_fence_before();
            mbar_arrive(&bars[C_GFREE + s]);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int cidx = j * BT + c;
                bool ok = row_ok && cidx < T;
                float l2 = row_lse2, dl = row_delta;
                if (KS) {
                    if (ok) {
                        l2 = __ldg(lse + (int64_t)row_base + cidx) * LOG2E;
                        dl = __ldg(delta + (int64_t)row_base + cidx);
                    }
                } else {
                    if (ok && mrow && mrow[cidx]) ok = false;
                }
                const float p = ok ? exp2f(g1[c] * (SM_SCALE * LOG2E) - l2) : 0.f;
                g1[c] = p;
                g2[c] = p * (g2[c] - dl) * SM_SCALE;
            }
            if (j > 0) mbar_wait(&bars[C_PDFREE], (j - 1) & 1);     // previous accumulating MMAs are done with the P / dS tiles
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
                uint32_t ph[4], pl[4], dh[4], dlw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    __nv_bfloat16 h0, l0, h1, l1;
                    split_bf16(g1[kc * 8 + 2 * e], h0, l0);
                    split_bf16(g1[kc * 8 + 2 * e + 1], h1, l1);
                    ph[e] = pack2(h0, h1); pl[e] = pack2(l0, l1);
                    split_bf16(g2[kc * 8 + 2 * e], h0, l0);
                    split_bf16(g2[kc * 8 + 2 * e + 1], h1, l1);
                    dh[e] = pack2(h0, h1); dlw[e] = pack2(l0, l1);
                }
                const uint32_t off = wr_off + kc * 128;
                if (KS) {
                    *reinterpret_cast<uint4*>(sPD + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                    *reinterpret_cast<uint4*>(sPD + PD_PLANE + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                }
                *reinterpret_cast<uint4*>(sPD + 2 * PD_PLANE + off) = make_uint4(dh[0], dh[1], dh[2], dh[3]);
                *reinterpret_cast<uint4*>(sPD + 3 * PD_PLANE + off) = make_uint4(dlw[0], dlw[1], dlw[2], dlw[3]);
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars[C_PDFULL]);
        }
        mbar_wait(&bars[C_DONE], 0);
        tc_fence_after();
        const bool store_ok = r_idx < T;
#pragma unroll
        for (int g = (KS ? 0 : 1); g < 2; ++g) {
            float* dst = (g == 0 ? out1 : out2) + ((int64_t)row_base + r_idx) * ld_d;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float t[32];
                tmem_ld32(lane_addr + 128 + g * 128 + ch * 32, t);
                if (store_ok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + ch * 32 + i) = make_float4(t[i], t[i + 1], t[i + 2], t[i + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(256) delta_kernel(float* __restrict__ delta, const float* __restrict__ o, const float* __restrict__ d_o, int64_t rows) {
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


int dip_attn_bwd_tc(float* dq, float* dk, float* dv, int ld_d, const void* qkv_hi, const void* qkv_lo, const void* do_hi, const void* do_lo,
                    const float* o, const float* d_o, const float* lse, const uint8_t* key_mask, int64_t mask_stride, float* delta_ws, int B, int T,
                    dip_stream_t stream) {
    DIP_REQUIRE(dq && dk && dv && qkv_hi && qkv_lo && do_hi && do_lo && o && d_o && lse && delta_ws && B > 0 && T > 0,
                "dip_attn_bwd_tc: null/empty argument");
    DIP_REQUIRE(ld_d % 4 == 0 && B <= 65535, "dip_attn_bwd_tc: bad ld_d / B");
    static bool attr = false;
    if (!attr) {
        DIP_CUDA(cudaFuncSetAttribute(tcattn::attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcattn::BWD_SMEM));
        DIP_CUDA(cudaFuncSetAttribute(tcattn::attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcattn::BWD_SMEM));
        attr = true;
    }
    cudaStream_t st = as_stream(stream);
    const uint64_t R = (uint64_t)B * T;
    CUtensorMap q_hi128, q_lo128, q_hi32, q_lo32, d_hi128, d_lo128, d_hi32, d_lo32;
    DIP_TRY(make_tmap_bf16_2d(&q_hi128, qkv_hi, R, 3 * D, 3 * D, 128, 64));
    DIP_TRY(make_tmap_bf16_2d(&q_lo128, qkv_lo, R, 3 * D, 3 * D, 128, 64));
    DIP_TRY(make_tmap_bf16_2d(&q_hi32, qkv_hi, R, 3 * D, 3 * D, 32, 64));
    DIP_TRY(make_tmap_bf16_2d(&q_lo32, qkv_lo, R, 3 * D, 3 * D, 32, 64));
    DIP_TRY(make_tmap_bf16_2d(&d_hi128, do_hi, R, D, D, 128, 64));
    DIP_TRY(make_tmap_bf16_2d(&d_lo128, do_lo, R, D, D, 128, 64));
    DIP_TRY(make_tmap_bf16_2d(&d_hi32, do_hi, R, D, D, 32, 64));
    DIP_TRY(make_tmap_bf16_2d(&d_lo32, do_lo, R, D, D, 32, 64));
    {
        DIP_PROF(DIP_PROF_OTHER, stream);
        int64_t blocks = ((int64_t)R + 7) / 8, cap = (int64_t)sm_count() * 8;
        tcattn::delta_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(delta_ws, o, d_o, (int64_t)R);
        DIP_LAUNCHED();
    }
    dim3 grid(cdiv(T, 128), B);
    {
        DIP_PROF(DIP_PROF_ATTN_BWD_DKDV, stream);
        tcattn::attn_bwd_tc_kernel<true><<<grid, 192, tcattn::BWD_SMEM, st>>>(q_hi128, q_lo128, q_hi32, q_lo32, d_hi128, d_lo128, d_hi32, d_lo32, dv, dk,
                                                                              ld_d, lse, delta_ws, key_mask, mask_stride, T);
        DIP_LAUNCHED();
    }
    {
        DIP_PROF(DIP_PROF_ATTN_BWD_DQ, stream);
        tcattn::attn_bwd_tc_kernel<false><<<grid, 192, tcattn::BWD_SMEM, st>>>(q_hi128, q_lo128, q_hi32, q_lo32, d_hi128, d_lo128, d_hi32, d_lo32, nullptr,
                                                                               dq, ld_d, lse, delta_ws, key_mask, mask_stride, T);
        DIP_LAUNCHED();
    }
    return 0;
}

}  // extern "C"
