// Flash attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), one head of width 128,
// key-padding mask, "bf16x2 split" numerics (see tc_common.cuh): forward and input-gradient backward.
//
// Operands are the bf16 (hi, lo) planes of the packed in_proj output, shape (B*T, 384) each, written
// by the producing GEMM's epilogue.  Every tile is moved by TMA (SWIZZLE_128B boxes of 64 columns)
// straight into the layout tcgen05.mma consumes; the same [rows x 128] tile serves as a K-major
// operand (Q, K in S = Q K^T) or as an MN-major B operand (V in P V) by descriptor only.
//
// Two measured facts of tcgen05.mma on B200 shape these kernels (tools/mma_bench.cu, tools/mma_ts_test.cu):
//  (1) with M = 128 every instruction costs max(~66, N/2) cycles when A is read from shared memory -- small-N
//      MMAs waste the pipe.  The score GEMMs therefore stack the hi and lo planes of the streamed operand
//      along N: one B operand [X_hi ; X_lo] of 2 x rows, issued once with A_hi and once with A_lo into the
//      SAME accumulator.  Column block 0 then holds A.X_hi, block 1 holds A.X_lo; their sum -- formed by the
//      row threads -- is the exact four-term product, for 2 MMAs per k-step instead of 3.
//  (2) the A operand may live in tensor memory (lane = row, 32-bit column = bf16 pair).  P and dS are
//      written there by the row threads (tcgen05.st) and never touch shared memory; the space pays for a
//      third pipeline stage.
//
// All kernels: 352 threads = 8 "row" warps + 1 producer warp + 2 MMA warps.
//   warps 0-7 : two warps per TMEM lane quarter; thread (row r, half h) owns row r of the 128-row tile
//               and half of its columns.
//   warp 8    : TMA producer; its 32 lanes also ballot the key-padding mask of every streamed tile into a
//               bit word, so the row warps never touch global memory inside the loop.
//   warp 11   : (forward only) second TMA producer, for the V ring
//   warps 9,10: MMA issue (one thread each).  Forward: warp 9 the score GEMMs, warp 10 the accumulating GEMMs -- issue is
//               throttled to the tensor pipe's rate and an mbarrier wait costs up to ~250 cycles (tools/ktrace.py), two
//               issuers hide each other's waits.  Backward kernels: warp 9 alone issues everything in an explicit order.
#include <math.h>
#include <stdlib.h>
#include "tc_common.cuh"

namespace dip {
namespace tcattn {

using namespace tc;

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float SM_SCALE = 0.08838834764831845f;
constexpr float SC2 = SM_SCALE * LOG2E;         // scores are exponentiated in base 2
constexpr int NTHREADS = 352, ROW_THREADS = 256, W_PRODUCER = 8, W_MMA = 9, W_MMA2 = 10;
constexpr int FWD_THREADS = 384, W_PRODUCER_V = 11;       // the forward has a second TMA producer warp (V ring)
constexpr uint32_t SMEM_LIMIT = 232448;

#ifdef DIP_TRACE
// Development aid (compile with -DDIP_TRACE): CTA (0,0) logs (event << 48 | tile << 32 | SM clock low bits) from its
// producer lane, MMA lane and row thread 0 into three regions of a global buffer; plain stores, no atomics.
__device__ unsigned long long* g_trace = nullptr;
struct Tracer {
    unsigned long long* p; int n;
    __device__ Tracer(int region) : p(nullptr), n(0) { if (g_trace && blockIdx.x == 0 && blockIdx.y == 0 && region >= 0) p = g_trace + region * 16384; }
    __device__ __forceinline__ void ev(int e, int tile) { if (p && n < 16384) p[n++] = ((unsigned long long)e << 48) | ((unsigned long long)tile << 32) | (unsigned long long)(unsigned int)clock64(); }
};
#define TR(e, t) tracer.ev((e), (t))
#else
struct Tracer { __device__ Tracer(int) {} };
#define TR(e, t)
#endif

__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* r) {
    asm volatile("st.global.cs.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]),
                 "r"(r[6]), "r"(r[7])
                 : "memory");
}

// 256-bit store of 8 floats (row-owner epilogues: 32 rows x 32 bytes per warp instruction = whole sectors)
__device__ __forceinline__ void st_global_v8f(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]),
                 "f"(v[6]), "f"(v[7])
                 : "memory");
}

__device__ __forceinline__ void require_aligned_smem(const void* p) {
    if ((smem_u32(p) & 1023u) != 0u) {
        if (threadIdx.x == 0) printf("dip: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
}

// =============================================================================================
// forward: CTA = 128 queries, key tiles of 64
// =============================================================================================
// shared memory: Q hi/lo resident (64 KB) | K ring, 3 stages x 32 KB, each [kb][hi,lo][64 keys][64 d]
//                (k-block kb: 128 stacked rows = B operand of the score MMA) | V ring, 2 stages x 32 KB,
//                each [hi,lo][d-half][64 keys][64 d] (MN-major B operand of P.V)
// tensor memory: [0,256) two score buffers of 128 columns (Q.K_hi | Q.K_lo), [256,384) P.V,
//                [384,512) two P buffers of (32 hi + 32 lo) packed columns
constexpr int BM = 128, BN = 64;
constexpr uint32_t Q_BYTES = 65536, K_STAGE = 32768, V_STAGE = 32768;
constexpr int K_STAGES = 3, V_STAGES = 2;
constexpr uint32_t OFF_K = Q_BYTES, OFF_V = OFF_K + K_STAGES * K_STAGE, OFF_BAR = OFF_V + V_STAGES * V_STAGE;
constexpr uint32_t OFF_BITS = OFF_BAR + 192, OFF_XM = OFF_BAR + 256;      // mask-bit ring (8 x u64), row-max exchange [2][2][128] f32
constexpr uint32_t FWD_SMEM = OFF_XM + 2048;
static_assert(OFF_BAR == 229376 && FWD_SMEM <= SMEM_LIMIT, "forward kernel shared memory");
constexpr uint32_t TM_S = 0, TM_PV = 256, TM_P = 384;

constexpr uint32_t IDESC_S = instr_desc(128, 128, 0, 0);    // Q (K-major) x [K_hi;K_lo] (K-major) -> 128 x 128
// the lo plane of the A operand meets only the hi half of the stacked tile (N = 64: the first 64 rows of every k-block): lo x lo is the
// 2^-16 term every other product of the split scheme drops too, and an N = 64 MMA takes as long as an N = 128 one but switches half the MACs
constexpr uint32_t IDESC_S64 = instr_desc(128, 64, 0, 0);
constexpr uint32_t IDESC_PV = instr_desc(128, 128, 0, 1);   // P (tensor memory) x V (MN-major) -> 128 x 128
constexpr float RESCALE_THRESHOLD = 8.0f;                   // log2 units: P stays below 2^8 between rescales

enum { B_QFULL = 0, B_KFULL = 1, B_KEMPTY = 4, B_VFULL = 7, B_VEMPTY = 9, B_SFULL = 11, B_SFREE = 13, B_PFULL = 15, B_PVFULL = 17, B_QFREE = 18 };

// Key-prefix hoisting (decoder layer 0): the first L of the 2L tokens are the instance embedding, the same for every step of a wave, so
// for a query that is itself an instance token the scores against the instance keys -- and with them the running maximum, the running
// sum and the un-normalised P.V accumulator after those key tiles -- never change.  A "save" launch (state_out) runs the key tiles
// [0, j_split) once per wave and writes that state; afterwards CTAs whose 128 queries all lie below hoist_rows start from it and stream
// the key tiles from j_split on only.  The state is exactly what the full loop holds at that point, so results are bit-identical.
// state layout: per (sample, query row) 132 floats = 128 accumulator columns | m_used | l_part of key half 0 | l_part of key half 1 | pad
constexpr int STATE_W = 132;

__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_hi128, const __grid_constant__ CUtensorMap tm_lo128,
                   const __grid_constant__ CUtensorMap tm_hi64, const __grid_constant__ CUtensorMap tm_lo64, float* __restrict__ o,
                   float* __restrict__ lse, const uint8_t* __restrict__ key_mask, int64_t mask_stride, int T, int q_begin,
                   const float* __restrict__ state_in, float* __restrict__ state_out, int j_split, int hoist_rows, int n_qx, int n_items) {
    extern __shared__ __align__(1024) uint8_t sm[];
    require_aligned_smem(sm);
    uint8_t* sQ = sm;
    uint8_t* sK = sm + OFF_K;
    uint8_t* sV = sm + OFF_V;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    volatile uint64_t* bits_ring = reinterpret_cast<volatile uint64_t*>(sm + OFF_BITS);
    volatile float* xm = reinterpret_cast<volatile float*>(sm + OFF_XM);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // PERSISTENT over work items (sample, 128-query tile): CTA c takes the items c, c + gridDim.x, ...  Every warp role walks the same
    // item sequence; ring slots and barrier phases are numbered by a running tile counter (g0 + jj), so the rings simply keep turning
    // across items: the producer fetches the next item's Q tile as soon as the last score product of the current one has completed
    // (B_QFREE) and its first key tiles as the ring frees up, i.e. the next prologue runs under the current item's last products
    // and epilogue, and there is no CTA exit / launch, TMEM allocation or descriptor fetch between items (per 32-tile item the
    // one-CTA-per-item form spent 6.6k + 6.8k of 76k cycles there, tools/ktrace_fixed.py).  With gridDim.x == n_items it degenerates
    // to one item per CTA.  Item order: the long items of a hoisted launch (queries past the instance tokens: all key tiles) first,
    // the short ones (resume at the split) after them, so every CTA gets the same mix.
    const int n_h = (state_in != nullptr) ? min(hoist_rows / BM, n_qx) : 0, n_l = n_qx - n_h;
    struct Item { int b, q0, jb, je, n_loop; bool hoisted; };
    auto item_of = [&](int idx) {
        Item it;
        int qx;
        const int n_long = n_l * (n_items / n_qx);
        if (idx < n_long) { it.b = idx / n_l; qx = n_h + idx % n_l; }
        else { const int r = idx - n_long; it.b = r / n_h; qx = r % n_h; }
        it.q0 = q_begin + qx * BM;                                  // queries q_begin .. T-1 only
        // key tiles jb .. je-1 (item-uniform): a save launch stops at the split, a hoisted item starts there
        it.hoisted = state_in != nullptr && it.q0 + BM <= hoist_rows;
        it.jb = it.hoisted ? j_split : 0;
        it.je = state_out != nullptr ? j_split : (T + BN - 1) / BN;
        it.n_loop = it.je - it.jb;
        return it;
    };
    Tracer tracer(threadIdx.x == 0 ? 0 : (lane == 0 && (warp == W_PRODUCER || warp == W_PRODUCER_V)) ? (warp == W_PRODUCER ? 1 : 4) : (lane == 0 && warp == W_MMA) ? 2 : (lane == 0 && warp == W_MMA2) ? 3 : -1);

    TR(1, 0);
    if (threadIdx.x == 0) {
        mbar_init(&bars[B_QFULL], 1);
        for (int s = 0; s < K_STAGES; ++s) { mbar_init(&bars[B_KFULL + s], 1); mbar_init(&bars[B_KEMPTY + s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars[B_VFULL + s], 1);
            mbar_init(&bars[B_VEMPTY + s], 1);
            mbar_init(&bars[B_SFULL + s], 1);
            mbar_init(&bars[B_SFREE + s], ROW_THREADS);
            mbar_init(&bars[B_PFULL + s], ROW_THREADS);
        }
        mbar_init(&bars[B_PVFULL], 1);
        mbar_init(&bars[B_QFREE], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    TR(2, 0);

    if (warp == W_PRODUCER) {
        // ------------------------------------------------------------------ TMA producer + mask ballots
        if (lane == 0) { tma_prefetch_desc(&tm_hi128); tma_prefetch_desc(&tm_lo128); tma_prefetch_desc(&tm_hi64); tma_prefetch_desc(&tm_lo64); }
        uint32_t g0 = 0;
        int n_it = 0;
        for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++n_it) {
            const Item I = item_of(idx);
            const int b = I.b, jb = I.jb;
            const uint8_t* mrow = key_mask ? key_mask + (int64_t)b * mask_stride : nullptr;
            auto load_K = [&](int j) {      // all lanes: ballot the mask of tile j; lane 0: wait for the stage, issue the 4 boxes
                const uint32_t gj = g0 + (uint32_t)(j - jb);
                const int s = (int)(gj % K_STAGES);
                const int k0 = j * BN + lane, k1 = k0 + 32;
                const bool m0 = (k0 >= T) || (mrow && mrow[k0]);
                const bool m1 = (k1 >= T) || (mrow && mrow[k1]);
                const uint32_t b0 = __ballot_sync(0xffffffffu, m0), b1 = __ballot_sync(0xffffffffu, m1);
                if (lane == 0) {
                    TR(40, j);
                    if (gj >= K_STAGES) mbar_wait(&bars[B_KEMPTY + s], ((gj / K_STAGES) - 1) & 1);
                    TR(41, j);
                    bits_ring[gj & 7] = (uint64_t)b0 | ((uint64_t)b1 << 32);     // published by the arrive below (release)
                    mbar_expect_tx(&bars[B_KFULL + s], K_STAGE);
                    uint8_t* dst = sK + s * K_STAGE;
                    for (int kb = 0; kb < 2; ++kb) {
                        tma_load_3d(dst + kb * 16384, &tm_hi64, &bars[B_KFULL + s], 128 + kb * 64, j * BN, b);           // rows 0-63 of the stacked operand
                        tma_load_3d(dst + kb * 16384 + 8192, &tm_lo64, &bars[B_KFULL + s], 128 + kb * 64, j * BN, b);    // rows 64-127
                    }
                }
                __syncwarp();
            };
            if (lane == 0) {
                if (n_it > 0) mbar_wait(&bars[B_QFREE], (n_it - 1) & 1);     // the previous item's last score product has read its Q tile
                mbar_expect_tx(&bars[B_QFULL], Q_BYTES);
                for (int kb = 0; kb < 2; ++kb) {
                    tma_load_3d(sQ + kb * 16384, &tm_hi128, &bars[B_QFULL], kb * 64, I.q0, b);
                    tma_load_3d(sQ + 32768 + kb * 16384, &tm_lo128, &bars[B_QFULL], kb * 64, I.q0, b);
                }
            }
            __syncwarp();
            // K runs up to three tiles ahead, independent of the V ring: a single producer thread alternating between the two rings
            // stalled the K loads behind the V ring's empty-waits and left the TMA queue idle a third of the time (tools/ktrace.py)
            for (int j = I.jb; j < I.je; ++j) load_K(j);
            g0 += (uint32_t)I.n_loop;
        }
    } else if (warp == W_PRODUCER_V) {
        // ------------------------------------------------------------------ TMA producer of the V ring
        if (lane == 0) {
            uint32_t g0 = 0;
            for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
                const Item I = item_of(idx);
                for (int j = I.jb; j < I.je; ++j) {
                    const uint32_t gj = g0 + (uint32_t)(j - I.jb);
                    const int s = (int)(gj & 1);
                    TR(42, j);
                    if (gj >= 2) mbar_wait(&bars[B_VEMPTY + s], ((gj >> 1) - 1) & 1);
                    TR(43, j);
                    mbar_expect_tx(&bars[B_VFULL + s], V_STAGE);
                    uint8_t* dst = sV + s * V_STAGE;
                    for (int kb = 0; kb < 2; ++kb) {
                        tma_load_3d(dst + kb * 8192, &tm_hi64, &bars[B_VFULL + s], 256 + kb * 64, j * BN, I.b);            // V hi, d-half kb
                        tma_load_3d(dst + 16384 + kb * 8192, &tm_lo64, &bars[B_VFULL + s], 256 + kb * 64, j * BN, I.b);    // V lo
                    }
                }
                g0 += (uint32_t)I.n_loop;
            }
        }
    } else if (warp == W_MMA || warp == W_MMA2) {
        // ------------------------------------------------------------------ MMA issuers: warp-uniform loops, the elected lane issues
        const uint64_t dQ = smem_desc(smem_u32(sQ), 16, 1024), dK = smem_desc(smem_u32(sK), 16, 1024), dVmn = smem_desc(smem_u32(sV), 8192, 1024);
        if (warp == W_MMA) {
          uint32_t g0 = 0;
          int n_it = 0;
          for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++n_it) {
            const int n_loop = item_of(idx).n_loop;
            mbar_wait(&bars[B_QFULL], n_it & 1);
            for (int jj = 0; jj < n_loop; ++jj) {
                const uint32_t gj = g0 + (uint32_t)jj;
                const int s = (int)(gj % K_STAGES), sb = (int)(gj & 1);
                TR(20, jj);
                mbar_wait(&bars[B_KFULL + s], (gj / K_STAGES) & 1);
                TR(21, jj);
                if (gj >= 2) mbar_wait(&bars[B_SFREE + sb], ((gj >> 1) - 1) & 1);
                TR(22, jj);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t kbase = desc_adv(dK, s * K_STAGE);
                    uint32_t acc = 0;
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl)                     // A = Q_hi, then Q_lo, into the same 128 columns
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                tc_mma(tmem + TM_S + sb * 128, desc_adv(dQ, pl * 32768 + kb * 16384 + ks * 32), desc_adv(kbase, kb * 16384 + ks * 32), pl == 0 ? IDESC_S : IDESC_S64, acc);
                                acc = 1;
                            }
                    tc_commit(&bars[B_SFULL + sb]);
                    tc_commit(&bars[B_KEMPTY + s]);
                    if (jj == n_loop - 1) tc_commit(&bars[B_QFREE]);       // the item's Q tile has been read for the last time
                }
                __syncwarp();
                TR(23, jj);
            }
            g0 += (uint32_t)n_loop;
          }
        } else {
          uint32_t g0 = 0;
          for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
            const int n_loop = item_of(idx).n_loop;
            for (int jj = 0; jj < n_loop; ++jj) {
                const uint32_t gj = g0 + (uint32_t)jj;
                const int s = (int)(gj & 1);
                TR(24, jj);
                mbar_wait(&bars[B_VFULL + s], (gj >> 1) & 1);
                TR(25, jj);
                mbar_wait(&bars[B_PFULL + s], (gj >> 1) & 1);       // P_j stored AND the previous P.V tile folded by the row warps
                TR(26, jj);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t vbase = desc_adv(dVmn, s * V_STAGE);
                    const uint32_t pbase = tmem + TM_P + s * 64;
                    uint32_t acc = 0;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            // P hi, hi, lo x V hi, lo, hi; V tile: two [64 keys x 64 d] blocks 8192 bytes apart (MN blocks), 16 keys per k-step = 2048 bytes
                            tc_mma_ts(tmem + TM_PV, pbase + (pass == 2 ? 32 : 0) + ks * 8, desc_adv(vbase, (pass == 1 ? 16384 : 0) + ks * 2048), IDESC_PV, acc);
                            acc = 1;
                        }
                    }
                    tc_commit(&bars[B_PVFULL]);
                    tc_commit(&bars[B_VEMPTY + s]);
                }
                __syncwarp();
                TR(28, jj);
            }
            g0 += (uint32_t)n_loop;
          }
        }
    } else {
        // ------------------------------------------------------------------ row warps: thread = (query row, key half)
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t g0 = 0;
        for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
        const Item I = item_of(idx);
        const int b = I.b, n_loop = I.n_loop;
        const bool hoisted = I.hoisted;
        const int row_base = b * T;
        const int t_q = I.q0 + row;
        float oacc[64];
        float m_used = -INFINITY, l_part = 0.f;
        if (hoisted) {
            // resume from the state the save launch left after the key tiles [0, j_split)
            const float* st = state_in + ((int64_t)b * hoist_rows + t_q) * STATE_W;
#pragma unroll
            for (int i = 0; i < 64; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(st + h * 64 + i);
                oacc[i] = v.x; oacc[i + 1] = v.y; oacc[i + 2] = v.z; oacc[i + 3] = v.w;
            }
            m_used = st[128];
            l_part = st[129 + h];
        } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) oacc[i] = 0.f;
        }
        for (int jj = 0; jj < n_loop; ++jj) {
            const uint32_t gj = g0 + (uint32_t)jj;
            const int s = (int)(gj & 1);
            TR(30, jj);
            mbar_wait(&bars[B_SFULL + s], (gj >> 1) & 1);
            TR(31, jj);
            tc_fence_after();
            float sv[32];
            {
                float s_lo[32];
                tmem_ld32x2(lane_addr + TM_S + s * 128 + h * 32, sv, lane_addr + TM_S + s * 128 + 64 + h * 32, s_lo);    // Q . K_hi, Q . K_lo
#pragma unroll
                for (int c = 0; c < 32; ++c) sv[c] += s_lo[c];
            }
            tc_fence_before();
            mbar_arrive(&bars[B_SFREE + s]);
            TR(32, jj);
            const uint32_t mb = (uint32_t)(bits_ring[gj & 7] >> (32 * h));
            float mx = -INFINITY;
            if (mb == 0u) {
#pragma unroll
                for (int c = 0; c < 32; ++c) mx = fmaxf(mx, sv[c]);
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if ((mb >> c) & 1u) sv[c] = -INFINITY;
                    mx = fmaxf(mx, sv[c]);
                }
            }
            // the two key halves of a row agree on the tile maximum through shared memory
            xm[(s * 2 + h) * 128 + row] = mx;
            named_bar_sync(1, ROW_THREADS);
            TR(33, jj);
            const float m_tile = fmaxf(mx, xm[(s * 2 + (h ^ 1)) * 128 + row]) * SC2;
            // lazy rescaling: the exponent base only moves when the row maximum grew by more than 2^8
            const bool need = m_tile > m_used + RESCALE_THRESHOLD;
            const float m_new = need ? m_tile : m_used;
            const float alpha = need ? ex2_approx(m_used - m_new) : 1.0f;
            const bool any_need = __any_sync(0xffffffffu, need);
            const float neg_m = (m_new == -INFINITY) ? 0.f : -m_new;
            float rowsum = 0.f;
            uint32_t phi[16], plo[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float p0 = ex2_approx(fmaf(sv[2 * c], SC2, neg_m));
                const float p1 = ex2_approx(fmaf(sv[2 * c + 1], SC2, neg_m));
                rowsum += p0 + p1;
                split2(p0, p1, phi[c], plo[c]);
            }
            l_part = l_part * alpha + rowsum;
            m_used = m_new;
            // P_j -> tensor memory (A operand of P.V): this half's 32 keys = 16 packed columns per plane.
            // Buffer s was last read by P_{j-2}.V_{j-2}, whose completion this thread observed one iteration ago.
            TR(34, jj);
            tmem_st16(lane_addr + TM_P + s * 64 + h * 16, phi);
            tmem_st16(lane_addr + TM_P + s * 64 + 32 + h * 16, plo);
            tmem_st_wait();
            TR(35, jj);
            if (jj > 0) {
                mbar_wait(&bars[B_PVFULL], (gj - 1) & 1);
                TR(36, jj);
                tc_fence_after();
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    float t[32];
                    tmem_ld32(lane_addr + TM_PV + h * 64 + ch * 32, t);
                    if (any_need) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) oacc[ch * 32 + i] = (oacc[ch * 32 + i] + t[i]) * alpha;
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) oacc[ch * 32 + i] += t[i];
                    }
                } 
                TR(37, jj);
            } else if (hoisted && any_need) {
                // first tile after a resume: no P.V product is pending, but the restored accumulator still follows the exponent base
#pragma unroll
                for (int i = 0; i < 64; ++i) oacc[i] *= alpha;
            }
            // one arrival publishes both facts the accumulate issuer needs: P_j is in tensor memory and the P.V tile is free again
            tc_fence_before();
            mbar_arrive(&bars[B_PFULL + s]);
        }
        TR(3, 0);
        mbar_wait(&bars[B_PVFULL], (g0 + (uint32_t)n_loop - 1) & 1);
        tc_fence_after();
        if (state_out != nullptr) {
            // save launch: the raw state after the key tiles [0, j_split) -- accumulator with the last product folded in, base, partial sums
            float* st = state_out + ((int64_t)b * hoist_rows + t_q) * STATE_W;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float t[32];
                tmem_ld32(lane_addr + TM_PV + h * 64 + ch * 32, t);
                if (t_q < hoist_rows) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(st + h * 64 + ch * 32 + i) = make_float4(oacc[ch * 32 + i] + t[i], oacc[ch * 32 + i + 1] + t[i + 1],
                                                                                            oacc[ch * 32 + i + 2] + t[i + 2], oacc[ch * 32 + i + 3] + t[i + 3]);
                }
            }
            if (t_q < hoist_rows) {
                if (h == 0) st[128] = m_used;
                st[129 + h] = l_part;
            }
        } else {
            // total row sum = the two halves' partial sums
            named_bar_sync(1, ROW_THREADS);
            xm[h * 128 + row] = l_part;
            named_bar_sync(1, ROW_THREADS);
            const float l_run = l_part + xm[(h ^ 1) * 128 + row];
            named_bar_sync(1, ROW_THREADS);                       // persistent: the next item's first tile reuses these exchange slots
            const float inv = 1.0f / l_run;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float t[32];
                tmem_ld32(lane_addr + TM_PV + h * 64 + ch * 32, t);
                if (t_q < T) {
                    // 256-bit stores: a thread owns its row, so a warp instruction is 32 rows x 32 bytes = whole sectors (16-byte stores
                    // were two instructions per sector)
                    float* orow = o + ((int64_t)row_base + t_q) * D + h * 64 + ch * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint32_t w[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) w[e] = __float_as_uint((oacc[ch * 32 + i + e] + t[i + e]) * inv);
                        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + i), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                                     "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                                     : "memory");
                    }
                }
            }
            if (lse && h == 0 && t_q < T) lse[(int64_t)row_base + t_q] = m_used * LN2 + logf(l_run);
        }
        g0 += (uint32_t)n_loop;
        }       // items
    }
    TR(4, 0);
    tc_fence_before();
    __syncthreads();
    TR(5, 0);
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// =============================================================================================
// backward (input gradients only, deterministic: no atomics)
// =============================================================================================
// Two kernels, each with a 128-row stationary operand pair in shared memory and 64-row streamed tiles:
//   dQ    (query-stationary): Q, dO stationary; K, V streamed;   S = Q K^T, dP = dO V^T;       dQ += dS K
//   dK/dV (key-stationary)  : K, V stationary; Q, dO streamed;   S^T = K Q^T, dP^T = V dO^T;   dV += P^T dO, dK += dS^T Q
// P = exp(S/sqrt(d) - lse[query]) and dS = P (dP - delta[query]) / sqrt(d) are formed by the row warps and written to tensor memory
// as the A operands of the accumulating GEMMs; the streamed tiles double as their MN-major B operands.  An earlier generic kernel
// streamed 32-row tiles (score MMAs at N = 64): an MMA with N = 64 costs 67 cycles, one with N = 128 only 73 (the 128-row A operand
// is fetched from shared memory either way), so the tensor pipe was < 50 % active (ncu) -- both kernels now use 64-row tiles.
constexpr uint32_t ST_PLANE = 32768;           // stationary plane: [2][128][64] bf16
constexpr uint32_t IDESC_ACC = instr_desc(128, 128, 0, 1);    // P/dS (tensor memory) x streamed tile (MN-major) -> 128 x 128

// =============================================================================================
// backward, dQ: query-stationary with 64-key tiles
// =============================================================================================
//   S = Q.[K_hi;K_lo]^T, dP = dO.[V_hi;V_lo]^T, dS = P (dP - delta) / sqrt(d), dQ += dS K
// The dQ pass needs neither P nor a second accumulator, which leaves tensor memory for more than accumulators: the dO planes
// (A operand of the dP GEMM) live THERE -- loaded once per CTA by the row threads -- and dS is written in place of the dP it was
// computed from.  That frees the 64 KB of shared memory dO would occupy: the K ring is three deep (K_j is read by S_j, issued one
// tile early, and again by dS_j.K_j), the V ring two.  One issuer warp, in the order
//      dP_j , S_{j+1} , dQ += dS_j K_j , dP_{j+1} , ...
// so that the score GEMM of the next tile runs while the row threads turn (S_j, dP_j) into dS_j, and -- the tensor pipe executing
// in issue order -- dP_{j+1} overwrites dS_j only after the product that reads it.
// shared memory: Q hi/lo (64 KB) | K ring 3 x 32 KB | V ring 2 x 32 KB; a streamed tile is [d-half][hi,lo][64 keys][64 d]
// tensor memory: [0,128) S, [128,256) dP then dS, [256,384) dQ, [384,512) dO hi | dO lo (64 packed columns each)
constexpr int QT = 64;
constexpr uint32_t Q_TILE = 32768;
constexpr uint32_t QOFF_X = 0, QOFF_Y = 2 * ST_PLANE, QOFF_U = 4 * ST_PLANE, QOFF_W = QOFF_U + 2 * Q_TILE, QOFF_BAR = QOFF_W + Q_TILE;      // dK/dV kernel
constexpr uint32_t QOFF_BITS = QOFF_BAR + 192;
constexpr uint32_t DQ_SMEM = QOFF_BAR + 256;
constexpr uint32_t QOFF_LSD = DQ_SMEM, KV_SMEM = QOFF_LSD + 8 * 256;         // dK/dV kernel: [8 row warps][32 lse*log2e | 32 delta] of the tile's queries
static_assert(KV_SMEM <= SMEM_LIMIT, "dK/dV kernel shared memory");
static_assert(QOFF_BAR == 229376 && DQ_SMEM <= SMEM_LIMIT, "backward kernels shared memory");
constexpr uint32_t DOFF_K = 2 * ST_PLANE, DOFF_V = DOFF_K + 3 * Q_TILE;                                                                      // dQ kernel
static_assert(DOFF_V + 2 * Q_TILE == QOFF_BAR, "dQ kernel shared memory");
constexpr uint32_t QTM_G1 = 0, QTM_G2 = 128, QTM_ACC = 256, QTM_DO = 384;
constexpr uint32_t IDESC_G128 = instr_desc(128, 128, 0, 0);
enum { D_XFULL = 0, D_DOFULL = 1, D_KFULL = 2, D_KEMPTY = 5, D_VFULL = 8, D_VEMPTY = 10, D_SFULL = 12, D_SFREE = 13, D_DPFULL = 14, D_DSFULL = 15, D_DONE = 16 };

__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv_hi128, const __grid_constant__ CUtensorMap tm_qkv_lo128,
                      const __grid_constant__ CUtensorMap tm_qkv_hi64, const __grid_constant__ CUtensorMap tm_qkv_lo64, const uint4* __restrict__ do_hi,
                      const uint4* __restrict__ do_lo, float* __restrict__ dq, int ld_d, const float* __restrict__ lse, const float* __restrict__ delta,
                      const uint8_t* __restrict__ key_mask, int64_t mask_stride, int T, int q_begin, int k_hi, int accumulate) {
    // k_hi: only the keys t < k_hi are streamed; accumulate != 0: the result is ADDED to dq (the keys >= k_hi were handled by the
    // streaming kernel below, which ran first on the same stream)
    extern __shared__ __align__(1024) uint8_t sm[];
    require_aligned_smem(sm);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + QOFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    volatile uint64_t* bits_ring = reinterpret_cast<volatile uint64_t*>(sm + QOFF_BITS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, r0 = q_begin + blockIdx.x * 128;      // first query of this CTA
    const int row_base = b * T;
    const int n_tiles = (k_hi + QT - 1) / QT;
    Tracer tracer(threadIdx.x == 0 ? 5 : (lane == 0 && warp == W_PRODUCER) ? 6 : (lane == 0 && warp == W_MMA) ? 7 : -1);

    if (threadIdx.x == 0) {
        mbar_init(&bars[D_XFULL], 1);
        mbar_init(&bars[D_DOFULL], ROW_THREADS);
        for (int s = 0; s < 3; ++s) { mbar_init(&bars[D_KFULL + s], 1); mbar_init(&bars[D_KEMPTY + s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&bars[D_VFULL + s], 1); mbar_init(&bars[D_VEMPTY + s], 1); }
        mbar_init(&bars[D_SFULL], 1);
        mbar_init(&bars[D_SFREE], ROW_THREADS);
        mbar_init(&bars[D_DPFULL], 1);
        mbar_init(&bars[D_DSFULL], ROW_THREADS);
        mbar_init(&bars[D_DONE], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint8_t* mrow = key_mask ? key_mask + (int64_t)b * mask_stride : nullptr;

    if (warp == W_PRODUCER) {
        if (lane == 0) {
            mbar_expect_tx(&bars[D_XFULL], 2 * ST_PLANE);
            for (int kb = 0; kb < 2; ++kb) {
                tma_load_3d(sm + QOFF_X + kb * 16384, &tm_qkv_hi128, &bars[D_XFULL], kb * 64, r0, b);
                tma_load_3d(sm + QOFF_X + ST_PLANE + kb * 16384, &tm_qkv_lo128, &bars[D_XFULL], kb * 64, r0, b);
            }
        }
        for (int j = 0; j < n_tiles; ++j) {
            const int sk = j % 3, sv = j & 1;
            const int k0 = j * QT + lane, k1 = k0 + 32;
            const bool m0 = (k0 >= k_hi) || (mrow && mrow[k0]);
            const bool m1 = (k1 >= k_hi) || (mrow && mrow[k1]);
            const uint32_t b0 = __ballot_sync(0xffffffffu, m0), b1 = __ballot_sync(0xffffffffu, m1);
            if (lane == 0) {
                TR(40, j);
                if (j >= 3) mbar_wait(&bars[D_KEMPTY + sk], ((j / 3) - 1) & 1);
                TR(41, j);
                bits_ring[j & 7] = (uint64_t)b0 | ((uint64_t)b1 << 32);
                mbar_expect_tx(&bars[D_KFULL + sk], Q_TILE);
                uint8_t* du = sm + DOFF_K + sk * Q_TILE;
                for (int kb = 0; kb < 2; ++kb) {
                    tma_load_3d(du + kb * 16384, &tm_qkv_hi64, &bars[D_KFULL + sk], 128 + kb * 64, j * QT, b);            // K hi, d-half kb
                    tma_load_3d(du + kb * 16384 + 8192, &tm_qkv_lo64, &bars[D_KFULL + sk], 128 + kb * 64, j * QT, b);     // K lo
                }
                TR(42, j);
                if (j >= 2) mbar_wait(&bars[D_VEMPTY + sv], ((j >> 1) - 1) & 1);
                TR(43, j);
                mbar_expect_tx(&bars[D_VFULL + sv], Q_TILE);
                uint8_t* dw = sm + DOFF_V + sv * Q_TILE;
                for (int kb = 0; kb < 2; ++kb) {
                    tma_load_3d(dw + kb * 16384, &tm_qkv_hi64, &bars[D_VFULL + sv], 256 + kb * 64, j * QT, b);            // V hi
                    tma_load_3d(dw + kb * 16384 + 8192, &tm_qkv_lo64, &bars[D_VFULL + sv], 256 + kb * 64, j * QT, b);     // V lo
                }
            }
            __syncwarp();
        }
    } else if (warp == W_MMA) {
        // warp-uniform control flow, the elected lane issues (see elect_one)
        const uint64_t dX = smem_desc(smem_u32(sm + QOFF_X), 16, 1024);
        const uint64_t dK = smem_desc(smem_u32(sm + DOFF_K), 16, 1024), dKmn = smem_desc(smem_u32(sm + DOFF_K), 16384, 1024);
        const uint64_t dV = smem_desc(smem_u32(sm + DOFF_V), 16, 1024);
        auto issue_S = [&](int j) {                             // S_j = Q . [K_hi;K_lo]^T, both operands in shared memory
            const int sk = j % 3;
            TR(24, j);
            mbar_wait(&bars[D_KFULL + sk], (j / 3) & 1);
            if (j >= 1) mbar_wait(&bars[D_SFREE], (j - 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t tp = desc_adv(dK, sk * Q_TILE);
                uint32_t acc = 0;
#pragma unroll
                for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            tc_mma(tmem + QTM_G1, desc_adv(dX, pl * ST_PLANE + kb * 16384 + ks * 32), desc_adv(tp, kb * 16384 + ks * 32), pl == 0 ? IDESC_G128 : IDESC_S64, acc);
                            acc = 1;
                        }
                tc_commit(&bars[D_SFULL]);
            }
            __syncwarp();
            TR(25, j);
        };
        mbar_wait(&bars[D_XFULL], 0);
        mbar_wait(&bars[D_DOFULL], 0);
        issue_S(0);
        for (int j = 0; j < n_tiles; ++j) {
            const int sk = j % 3, sv = j & 1;
            TR(20, j);
            mbar_wait(&bars[D_VFULL + sv], (j >> 1) & 1);
            TR(21, j);
            tc_fence_after();
            if (elect_one()) {                                      // dP_j = dO (tensor memory) . [V_hi;V_lo]^T
                const uint64_t tp = desc_adv(dV, sv * Q_TILE);
                uint32_t acc = 0;
#pragma unroll
                for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            tc_mma_ts(tmem + QTM_G2, tmem + QTM_DO + pl * 64 + (kb * 4 + ks) * 8, desc_adv(tp, kb * 16384 + ks * 32), pl == 0 ? IDESC_G128 : IDESC_S64, acc);
                            acc = 1;
                        }
                tc_commit(&bars[D_DPFULL]);
                tc_commit(&bars[D_VEMPTY + sv]);
            }
            __syncwarp();
            TR(22, j);
            if (j + 1 < n_tiles) issue_S(j + 1);
            TR(26, j);
            mbar_wait(&bars[D_DSFULL], j & 1);
            TR(27, j);
            tc_fence_after();
            if (elect_one()) {                                      // dQ += dS_j K_j ; dS_j sits in the dP region (see the row threads)
                const uint64_t tl = desc_adv(dKmn, sk * Q_TILE);
                uint32_t acc = j > 0 ? 1u : 0u;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        // dS hi, hi, lo x K hi, lo, hi as MN-major B: d-halves 16384 bytes apart, 16 keys per k-step
                        tc_mma_ts(tmem + QTM_ACC, tmem + QTM_G2 + (ks >> 1) * 32 + (ks & 1) * 8 + (pass == 2 ? 16 : 0),
                                  desc_adv(tl, (pass == 1 ? 8192 : 0) + ks * 2048), IDESC_ACC, acc);
                        acc = 1;
                    }
                }
                tc_commit(&bars[D_KEMPTY + sk]);
            }
            __syncwarp();
            TR(28, j);
        }
        if (elect_one()) tc_commit(&bars[D_DONE]);
        __syncwarp();
    } else if (warp < 8) {
        // ------------------------------------------------------------------ row warps: thread = (query row, 32 of the 64 keys)
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        const int r_idx = r0 + row;
        const bool row_ok = r_idx < T;
        float row_lse2 = 0.f, row_delta = 0.f;
        if (row_ok) {
            row_lse2 = lse[(int64_t)row_base + r_idx] * LOG2E;
            row_delta = delta[(int64_t)row_base + r_idx];
        }
        {
            // this thread's dO plane (h = 0: hi, h = 1: lo) of its query row -> tensor memory: 128 bf16 = 64 packed columns
            const uint4* src = (h == 0 ? do_hi : do_lo) + ((int64_t)row_base + r_idx) * 16;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t w[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 v = row_ok ? __ldg(src + c * 4 + i) : make_uint4(0u, 0u, 0u, 0u);
                    w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
                }
                tmem_st16(lane_addr + QTM_DO + h * 64 + c * 16, w);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars[D_DOFULL]);
        }
        for (int j = 0; j < n_tiles; ++j) {
            TR(30, j);
            mbar_wait(&bars[D_SFULL], j & 1);
            TR(31, j);
            tc_fence_after();
            float pr[32];
            {
                float t[32];
                tmem_ld32x2(lane_addr + QTM_G1 + h * 32, pr, lane_addr + QTM_G1 + 64 + h * 32, t);
                tc_fence_before();
                mbar_arrive(&bars[D_SFREE]);
                const uint32_t bad = (uint32_t)(bits_ring[j & 7] >> (32 * h));
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const bool ok = row_ok && !((bad >> c) & 1u);
                    pr[c] = ok ? ex2_approx(fmaf(pr[c] + t[c], SC2, -row_lse2)) * SM_SCALE : 0.f;
                }
            }
            TR(32, j);
            mbar_wait(&bars[D_DPFULL], j & 1);
            TR(33, j);
            tc_fence_after();
            {
                float dp[32], t[32];
                uint32_t dh[16], dl[16];
                tmem_ld32x2(lane_addr + QTM_G2 + h * 32, dp, lane_addr + QTM_G2 + 64 + h * 32, t);
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    split2(pr[2 * c] * (dp[2 * c] + t[2 * c] - row_delta), pr[2 * c + 1] * (dp[2 * c + 1] + t[2 * c + 1] - row_delta), dh[c], dl[c]);
                TR(34, j);
                // in place: the packed planes go into the first 32 of the 64 dP columns this thread has just read
                tmem_st16(lane_addr + QTM_G2 + h * 32, dh);
                tmem_st16(lane_addr + QTM_G2 + h * 32 + 16, dl);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars[D_DSFULL]);
            }
            TR(35, j);
        }
        mbar_wait(&bars[D_DONE], 0);
        tc_fence_after();
        float* dst = dq + ((int64_t)row_base + r_idx) * ld_d + h * 64;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            float t[32];
            tmem_ld32(lane_addr + QTM_ACC + h * 64 + ch * 32, t);
            if (row_ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float4 v = make_float4(t[i], t[i + 1], t[i + 2], t[i + 3]);
                    if (accumulate) {
                        const float4 o = *reinterpret_cast<const float4*>(dst + ch * 32 + i);
                        v = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w);
                    }
                    *reinterpret_cast<float4*>(dst + ch * 32 + i) = v;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// =============================================================================================
// backward, dQ from stored dS: streaming GEMM
// =============================================================================================
// The key-stationary kernel below forms every dS tile anyway (it needs it for dK); it also writes it to global memory as bf16 hi/lo
// planes of dS^T (row = key, column = query).  This kernel then only has to contract them with K:  dQ[128 q x 128 d] += dS K over the
// keys, 64 at a time -- no score recomputation (the query-stationary kernel above executes S, dP and dQ: three products for one).
// Both operands come through TMA from a three-stage ring: the dS^T tile is consumed as an MN-major A operand (queries contiguous),
// the K tile as an MN-major B operand; three bf16 passes (hi.hi, hi.lo, lo.hi) into one TMEM accumulator.  The kernel is bound by the
// 64 KB per stage it has to pull in (half of it from HBM: dS is 4 bytes per (query, key) pair), not by the 12 MMAs per stage.
// shared memory per stage: dS^T [plane][query half][64 keys][64 q] (32 KB) | K [d-half][hi,lo][64 keys][64 d] (32 KB)
constexpr uint32_t SQ_STAGE = 65536, SQ_DS = 0, SQ_K = 32768, SQ_BAR = 3 * SQ_STAGE, SQ_SMEM = SQ_BAR + 256;
constexpr uint32_t IDESC_SQ = instr_desc(128, 128, 1, 1);
enum { S_FULL = 0, S_EMPTY = 3, S_DONE = 6 };

__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_dq_stream_tc_kernel(const __grid_constant__ CUtensorMap tm_ds_hi, const __grid_constant__ CUtensorMap tm_ds_lo,
                             const __grid_constant__ CUtensorMap tm_qkv_hi64, const __grid_constant__ CUtensorMap tm_qkv_lo64, float* __restrict__ dq,
                             int ld_d, int T, int q_begin, int k_lo, int q_origin, int c_start) {
    // Column c of the dS^T buffer is query q_origin + c, its row r is key k_lo + r.  CTA i works on the 128 columns from c_start + 128 i:
    // c_start is a multiple of the writer's 64-query tiles (TMA needs 16-byte aligned box origins), so the first CTA may begin below
    // q_begin -- those rows are computed and dropped.  Keys k_lo .. T-1.
    extern __shared__ __align__(1024) uint8_t sm[];
    require_aligned_smem(sm);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SQ_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col0 = c_start + blockIdx.x * 128;
    const int b = blockIdx.y, r0 = q_origin + col0;
    const int n_tiles = (T - k_lo + QT - 1) / QT;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 3; ++s) { mbar_init(&bars[S_FULL + s], 1); mbar_init(&bars[S_EMPTY + s], 1); }
        mbar_init(&bars[S_DONE], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == W_PRODUCER) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_ds_hi); tma_prefetch_desc(&tm_ds_lo); tma_prefetch_desc(&tm_qkv_hi64); tma_prefetch_desc(&tm_qkv_lo64);
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j % 3;
                if (j >= 3) mbar_wait(&bars[S_EMPTY + s], ((j / 3) - 1) & 1);
                mbar_expect_tx(&bars[S_FULL + s], SQ_STAGE);
                uint8_t* st = sm + s * SQ_STAGE;
                for (int qb = 0; qb < 2; ++qb) {          // rows past the last key and columns past the last query tile read as zero
                    tma_load_3d(st + SQ_DS + qb * 8192, &tm_ds_hi, &bars[S_FULL + s], col0 + qb * 64, j * QT, b);
                    tma_load_3d(st + SQ_DS + 16384 + qb * 8192, &tm_ds_lo, &bars[S_FULL + s], col0 + qb * 64, j * QT, b);
                }
                for (int kb = 0; kb < 2; ++kb) {
                    tma_load_3d(st + SQ_K + kb * 16384, &tm_qkv_hi64, &bars[S_FULL + s], 128 + kb * 64, k_lo + j * QT, b);            // K hi, d-half kb
                    tma_load_3d(st + SQ_K + kb * 16384 + 8192, &tm_qkv_lo64, &bars[S_FULL + s], 128 + kb * 64, k_lo + j * QT, b);     // K lo
                }
            }
        }
    } else if (warp == W_MMA) {
        const uint64_t dA = smem_desc(smem_u32(sm + SQ_DS), 8192, 1024);       // MN-major A: query halves 8192 bytes apart, 16 keys per k-step
        const uint64_t dB = smem_desc(smem_u32(sm + SQ_K), 16384, 1024);       // MN-major B: d-halves 16384 bytes apart
        for (int j = 0; j < n_tiles; ++j) {
            const int s = j % 3;
            mbar_wait(&bars[S_FULL + s], (j / 3) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t a = desc_adv(dA, s * SQ_STAGE), k = desc_adv(dB, s * SQ_STAGE);
                uint32_t acc = j > 0 ? 1u : 0u;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass)                        // dS hi, hi, lo x K hi, lo, hi
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        tc_mma(tmem, desc_adv(a, (pass == 2 ? 16384 : 0) + ks * 2048), desc_adv(k, (pass == 1 ? 8192 : 0) + ks * 2048), IDESC_SQ, acc);
                        acc = 1;
                    }
                tc_commit(&bars[S_EMPTY + s]);
            }
            __syncwarp();
        }
        if (elect_one()) tc_commit(&bars[S_DONE]);
        __syncwarp();
    } else if (warp < 8) {
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        const int r_idx = r0 + row;
        mbar_wait(&bars[S_DONE], 0);
        tc_fence_after();
        float* dst = dq + ((int64_t)b * T + r_idx) * ld_d + h * 64;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            float t[32];
            tmem_ld32(lane_addr + h * 64 + ch * 32, t);
            if (r_idx < T && r_idx >= q_begin) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) st_global_v8f(dst + ch * 32 + i, t + i);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// =============================================================================================
// backward, dK / dV: key-stationary with 64-query tiles
// =============================================================================================
// Same idea as the dQ kernel, for the pass that needs two accumulators.  Tensor memory is the tight resource: dV and dK take 256
// columns, S^T and dP^T of a 64-query tile (hi- and lo-part stacked) 128 each -- nothing is left for the packed P / dS operands.
// They are written IN PLACE: a row thread reads its 32 + 32 score columns and stores the 16 + 16 packed columns of P (or dS) back
// into the first half of the columns it just read, so no other thread's unread data is touched.  One thread issues everything in
// the order
//      S^T_j , dP^T_j , dV += P_j^T dO_j , S^T_{j+1} , dK += dS_j^T Q_j , dP^T_{j+1} , ...
// and because the tensor pipe executes in issue order no barrier is needed between a product that reads P_j and the score GEMM that
// overwrites its region; every wait the issuer performs (P_j stored, dS_j stored, next tile loaded) is covered by two products of
// queued work.  Shared memory: K hi/lo, V hi/lo stationary (128 KB) | Q ring 2 x 32 KB | dO 32 KB, single-buffered.  Score MMAs are
// N = 128: 73 cycles for twice the work of the N = 64 MMAs (67 cycles) the 32-query version had to use.
// tensor memory: [0,128) S^T_j then P_j, [128,256) dP^T_j then dS_j, [256,384) dV, [384,512) dK
constexpr uint32_t KTM_G1 = 0, KTM_G2 = 128, KTM_ACC1 = 256, KTM_ACC2 = 384;
enum { E_XFULL = 0, E_UFULL = 1, E_UEMPTY = 3, E_WFULL = 5, E_WEMPTY = 6, E_G1FULL = 7, E_G2FULL = 8, E_PFULL = 9, E_DSFULL = 10, E_DONE = 11, E_YFULL = 12 };

__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_kv_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv_hi128, const __grid_constant__ CUtensorMap tm_qkv_lo128,
                      const __grid_constant__ CUtensorMap tm_qkv_hi64, const __grid_constant__ CUtensorMap tm_qkv_lo64,
                      const __grid_constant__ CUtensorMap tm_do_hi64, const __grid_constant__ CUtensorMap tm_do_lo64, float* __restrict__ dv,
                      float* __restrict__ dk, int ld_d, const float* __restrict__ lse, const float* __restrict__ delta,
                      const uint8_t* __restrict__ key_mask, int64_t mask_stride, int T, int k_begin, int q_begin,
                      uint8_t* __restrict__ ds_hi, uint8_t* __restrict__ ds_lo, int ds_tq_pad, int ds_q_from) {
    extern __shared__ __align__(1024) uint8_t sm[];
    require_aligned_smem(sm);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + QOFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
    volatile uint64_t* bits_ring = reinterpret_cast<volatile uint64_t*>(sm + QOFF_BITS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, r0 = k_begin + blockIdx.x * 128;      // first key of this CTA
    const int row_base = b * T;
    const int n_tiles = (T - q_begin + QT - 1) / QT;                // queries q_begin .. T-1 are streamed
    Tracer tracer(threadIdx.x == 0 ? 0 : (lane == 0 && warp == W_PRODUCER) ? 1 : (lane == 0 && warp == W_MMA) ? 2 : -1);

    TR(1, 0);
    if (threadIdx.x == 0) {
        mbar_init(&bars[E_XFULL], 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars[E_UFULL + s], 1);
            mbar_init(&bars[E_UEMPTY + s], 1);
        }
        mbar_init(&bars[E_WFULL], 1);
        mbar_init(&bars[E_WEMPTY], 1);
        mbar_init(&bars[E_G1FULL], 1);
        mbar_init(&bars[E_G2FULL], 1);
        mbar_init(&bars[E_PFULL], ROW_THREADS);
        mbar_init(&bars[E_DSFULL], ROW_THREADS);
        mbar_init(&bars[E_DONE], 1);
        mbar_init(&bars[E_YFULL], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint8_t* mrow = key_mask ? key_mask + (int64_t)b * mask_stride : nullptr;

    if (warp == W_PRODUCER) {
        // start-up order = order of first use: K, Q_0 (S^T_0), then V, dO_0 (dP^T_0) -- K and V on separate barriers
        auto load_V = [&]() {
            mbar_expect_tx(&bars[E_YFULL], 2 * ST_PLANE);
            for (int kb = 0; kb < 2; ++kb) {
                tma_load_3d(sm + QOFF_Y + kb * 16384, &tm_qkv_hi128, &bars[E_YFULL], 256 + kb * 64, r0, b);                 // V hi
                tma_load_3d(sm + QOFF_Y + ST_PLANE + kb * 16384, &tm_qkv_lo128, &bars[E_YFULL], 256 + kb * 64, r0, b);      // V lo
            }
        };
        if (lane == 0) {
            mbar_expect_tx(&bars[E_XFULL], 2 * ST_PLANE);
            for (int kb = 0; kb < 2; ++kb) {
                tma_load_3d(sm + QOFF_X + kb * 16384, &tm_qkv_hi128, &bars[E_XFULL], 128 + kb * 64, r0, b);                 // K hi
                tma_load_3d(sm + QOFF_X + ST_PLANE + kb * 16384, &tm_qkv_lo128, &bars[E_XFULL], 128 + kb * 64, r0, b);      // K lo
            }
        }
        for (int j = 0; j < n_tiles; ++j) {
            const int s = j & 1;
            const int q0 = q_begin + j * QT;
            const uint32_t b0 = __ballot_sync(0xffffffffu, q0 + lane >= T), b1 = __ballot_sync(0xffffffffu, q0 + 32 + lane >= T);
            if (lane == 0) {
                TR(40, j);
                if (j >= 2) mbar_wait(&bars[E_UEMPTY + s], ((j >> 1) - 1) & 1);
                TR(41, j);
                bits_ring[j & 7] = (uint64_t)b0 | ((uint64_t)b1 << 32);
                mbar_expect_tx(&bars[E_UFULL + s], Q_TILE);
                uint8_t* du = sm + QOFF_U + s * Q_TILE;
                for (int kb = 0; kb < 2; ++kb) {
                    tma_load_3d(du + kb * 16384, &tm_qkv_hi64, &bars[E_UFULL + s], kb * 64, q0, b);               // Q hi, d-half kb
                    tma_load_3d(du + kb * 16384 + 8192, &tm_qkv_lo64, &bars[E_UFULL + s], kb * 64, q0, b);        // Q lo
                }
                TR(42, j);
                if (j == 0) load_V();
                if (j >= 1) mbar_wait(&bars[E_WEMPTY], (j - 1) & 1);
                TR(43, j);
                mbar_expect_tx(&bars[E_WFULL], Q_TILE);
                uint8_t* dw = sm + QOFF_W;
                for (int kb = 0; kb < 2; ++kb) {
                    tma_load_3d(dw + kb * 16384, &tm_do_hi64, &bars[E_WFULL], kb * 64, q0, b);                    // dO hi
                    tma_load_3d(dw + kb * 16384 + 8192, &tm_do_lo64, &bars[E_WFULL], kb * 64, q0, b);             // dO lo
                }
            }
            __syncwarp();
        }
    } else if (warp == W_MMA) {
        // warp-uniform issue loop: every lane waits, the elected lane issues (see elect_one)
        const uint64_t dX = smem_desc(smem_u32(sm + QOFF_X), 16, 1024), dY = smem_desc(smem_u32(sm + QOFF_Y), 16, 1024);
        const uint64_t dU = smem_desc(smem_u32(sm + QOFF_U), 16, 1024), dW = smem_desc(smem_u32(sm + QOFF_W), 16, 1024);
        const uint64_t dUmn = smem_desc(smem_u32(sm + QOFF_U), 16384, 1024), dWmn = smem_desc(smem_u32(sm + QOFF_W), 16384, 1024);
        auto issue_G = [&](uint64_t st, uint64_t tp, uint32_t dst) {
            uint32_t acc = 0;
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)                      // A = stationary hi, then lo, into the same 128 columns
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        tc_mma(dst, desc_adv(st, pl * ST_PLANE + kb * 16384 + ks * 32), desc_adv(tp, kb * 16384 + ks * 32), pl == 0 ? IDESC_G128 : IDESC_S64, acc);
                        acc = 1;
                    }
        };
        // A operand in place of the scores: the thread that owned queries 32h .. 32h+31 left their packed hi plane in columns
        // [32h, 32h+16) and the lo plane in [32h+16, 32h+32) of the region; k-step ks covers queries 16ks .. 16ks+15
        auto issue_acc = [&](uint32_t dst, uint32_t region, uint64_t tile_mn, bool first) {
            uint32_t acc = first ? 0u : 1u;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t a = region + (ks >> 1) * 32 + (ks & 1) * 8 + (pass == 2 ? 16 : 0);       // A hi, hi, lo
                    // streamed tile as MN-major B (hi, lo, hi): d-halves 16384 bytes apart, 16 queries per k-step = 2048 bytes
                    tc_mma_ts(dst, a, desc_adv(tile_mn, (pass == 1 ? 8192 : 0) + ks * 2048), IDESC_ACC, acc);
                    acc = 1;
                }
            }
        };
        auto issue_scores1 = [&](int j) {                       // S^T_j
            const int s = j & 1;
            TR(20, j);
            mbar_wait(&bars[E_UFULL + s], (j >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                issue_G(dX, desc_adv(dU, s * Q_TILE), tmem + KTM_G1);
                tc_commit(&bars[E_G1FULL]);
            }
            __syncwarp();
            TR(22, j);
        };
        auto issue_scores2 = [&](int j) {                       // dP^T_j
            mbar_wait(&bars[E_WFULL], j & 1);
            TR(24, j);
            tc_fence_after();
            if (elect_one()) {
                issue_G(dY, dW, tmem + KTM_G2);
                tc_commit(&bars[E_G2FULL]);
            }
            __syncwarp();
            TR(25, j);
        };
        mbar_wait(&bars[E_XFULL], 0);
        issue_scores1(0);
        mbar_wait(&bars[E_YFULL], 0);
        issue_scores2(0);
        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(&bars[E_PFULL], j & 1);                                    // dV += P_j^T dO_j
            tc_fence_after();
            if (elect_one()) {
                issue_acc(tmem + KTM_ACC1, tmem + KTM_G1, dWmn, j == 0);
                tc_commit(&bars[E_WEMPTY]);
            }
            __syncwarp();
            TR(28, j);
            if (j + 1 < n_tiles) issue_scores1(j + 1);                            // S^T_{j+1} overwrites P_j: after its product, in pipe order
            mbar_wait(&bars[E_DSFULL], j & 1);                                   // dK += dS_j^T Q_j
            tc_fence_after();
            if (elect_one()) {
                issue_acc(tmem + KTM_ACC2, tmem + KTM_G2, desc_adv(dUmn, (j & 1) * Q_TILE), j == 0);
                tc_commit(&bars[E_UEMPTY + (j & 1)]);
            }
            __syncwarp();
            TR(23, j);
            if (j + 1 < n_tiles) issue_scores2(j + 1);                            // dP^T_{j+1} overwrites dS_j
        }
        if (elect_one()) tc_commit(&bars[E_DONE]);
        __syncwarp();
    } else if (warp < 8) {
        // ------------------------------------------------------------------ row warps: thread = (key row, 32 of the 64 queries)
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        const int r_idx = r0 + row;
        bool row_ok = r_idx < T;
        if (row_ok && mrow && mrow[r_idx]) row_ok = false;          // masked key: its column of P is zero
        // dS^T for the streaming dQ kernel: row = this key, 32 consecutive queries = 64 contiguous bytes per plane (the MN-major A operand
        // of dQ = dS K needs exactly this [key][query] order).  The stores of tile j are issued half an iteration later, after P_{j+1}
        // has been handed to the tensor pipe: right after dS_j they would delay P_{j+1} (every 256-bit store instruction is 32 L1
        // wavefronts -- 32 different rows) and with it the product the pipe is waiting for; here they fall into the wait for dP_{j+1}.
        uint32_t dh[16], dl[16];
        size_t ds_off = 0;
        bool ds_pending = false;
        auto flush_ds = [&]() {
            if (ds_pending) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    st_global_v8(ds_hi + ds_off + 32 * c, dh + 8 * c);
                    st_global_v8(ds_lo + ds_off + 32 * c, dl + 8 * c);
                }
                ds_pending = false;
            }
        };
        // lse and delta of this warp's 32 queries of a tile: fetched one tile ahead by lane = query, parked in a per-warp block of shared
        // memory and read back as broadcast float4 (8 + 8 loads per tile instead of 64 shuffles, and no load latency inside the tile)
        float* lsd = reinterpret_cast<float*>(sm + QOFF_LSD) + warp * 64;
        const float4* lsd4 = reinterpret_cast<const float4*>(lsd);
        const bool all_rows_ok = __all_sync(0xffffffffu, row_ok);
        float n_lse2 = 0.f, n_delta = 0.f;
        {
            const int cq = q_begin + h * 32 + lane;
            if (cq < T) {
                n_lse2 = __ldg(lse + (int64_t)row_base + cq) * LOG2E;
                n_delta = __ldg(delta + (int64_t)row_base + cq);
            }
        }
        for (int j = 0; j < n_tiles; ++j) {
            __syncwarp();                                           // every lane is done with the previous tile's values
            lsd[lane] = n_lse2;
            lsd[32 + lane] = n_delta;
            __syncwarp();
            {
                const int cq = q_begin + (j + 1) * QT + h * 32 + lane;
                n_lse2 = 0.f; n_delta = 0.f;
                if (j + 1 < n_tiles && cq < T) {
                    n_lse2 = __ldg(lse + (int64_t)row_base + cq) * LOG2E;
                    n_delta = __ldg(delta + (int64_t)row_base + cq);
                }
            }
            TR(30, j);
            // read back BEFORE the wait: with the tensor pipe streaming its operands out of shared memory an LDS takes hundreds of cycles
            float4 l4[8];
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) l4[c4] = lsd4[c4];
            const uint32_t bad = (uint32_t)(bits_ring[j & 7] >> (32 * h));
            mbar_wait(&bars[E_G1FULL], j & 1);
            TR(31, j);
            tc_fence_after();
            float pr[32];
            {
                float t[32];
                uint32_t ph[16], pl[16];
                tmem_ld32x2(lane_addr + KTM_G1 + h * 32, pr, lane_addr + KTM_G1 + 64 + h * 32, t);
                if (all_rows_ok && bad == 0u) {
                    // no masked key in this warp's rows, no query past the end of the sample: nothing to select (every tile but the last,
                    // every warp without padding keys)
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 l = l4[c4];
                        pr[4 * c4] = ex2_approx(fmaf(pr[4 * c4] + t[4 * c4], SC2, -l.x));
                        pr[4 * c4 + 1] = ex2_approx(fmaf(pr[4 * c4 + 1] + t[4 * c4 + 1], SC2, -l.y));
                        pr[4 * c4 + 2] = ex2_approx(fmaf(pr[4 * c4 + 2] + t[4 * c4 + 2], SC2, -l.z));
                        pr[4 * c4 + 3] = ex2_approx(fmaf(pr[4 * c4 + 3] + t[4 * c4 + 3], SC2, -l.w));
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float l2 = lsd[c];
                        const bool ok = row_ok && !((bad >> c) & 1u);
                        pr[c] = ok ? ex2_approx(fmaf(pr[c] + t[c], SC2, -l2)) : 0.f;
                    }
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) split2(pr[2 * c], pr[2 * c + 1], ph[c], pl[c]);
                // in place: the packed planes go into the first 32 of the 64 columns this thread has just read
                tmem_st16(lane_addr + KTM_G1 + h * 32, ph);
                tmem_st16(lane_addr + KTM_G1 + h * 32 + 16, pl);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars[E_PFULL]);
            }
            flush_ds();                                             // dS^T of the previous tile
            TR(32, j);
            float4 d4[8];
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) d4[c4] = lsd4[8 + c4];
            mbar_wait(&bars[E_G2FULL], j & 1);
            TR(33, j);
            tc_fence_after();
            {
                float g2[32], t[32];
                tmem_ld32x2(lane_addr + KTM_G2 + h * 32, g2, lane_addr + KTM_G2 + 64 + h * 32, t);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 dq4 = d4[c4];
                    const float d0 = pr[4 * c4] * (g2[4 * c4] + t[4 * c4] - dq4.x) * SM_SCALE;
                    const float d1 = pr[4 * c4 + 1] * (g2[4 * c4 + 1] + t[4 * c4 + 1] - dq4.y) * SM_SCALE;
                    const float d2 = pr[4 * c4 + 2] * (g2[4 * c4 + 2] + t[4 * c4 + 2] - dq4.z) * SM_SCALE;
                    const float d3 = pr[4 * c4 + 3] * (g2[4 * c4 + 3] + t[4 * c4 + 3] - dq4.w) * SM_SCALE;
                    split2(d0, d1, dh[2 * c4], dl[2 * c4]);
                    split2(d2, d3, dh[2 * c4 + 1], dl[2 * c4 + 1]);
                }
                TR(34, j);
                tmem_st16(lane_addr + KTM_G2 + h * 32, dh);
                tmem_st16(lane_addr + KTM_G2 + h * 32 + 16, dl);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars[E_DSFULL]);
                // query tiles entirely below ds_q_from are never read by the dQ kernel
                if (ds_hi != nullptr && r_idx < T && q_begin + j * QT + QT > ds_q_from) {
                    ds_off = (((size_t)b * (size_t)(T - k_begin) + (size_t)(r_idx - k_begin)) * (size_t)ds_tq_pad + (size_t)(j * QT + h * 32)) * 2u;
                    ds_pending = true;
                }
            }
            TR(35, j);
        }
        flush_ds();
        TR(3, 0);
        mbar_wait(&bars[E_DONE], 0);
        tc_fence_after();
        // dV, dK: 32 rows x 64 columns per warp, transposed through a private 4 KB block of the (now idle) Q ring so that every global
        // store instruction writes 4 rows x 128 contiguous bytes instead of 32 rows x 16 bytes
        float* stg = reinterpret_cast<float*>(sm + QOFF_U) + warp * 1024;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            float* dst = (g == 0 ? dv : dk) + ((int64_t)row_base + r0 + q * 32) * ld_d + h * 64;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float t[32];
                tmem_ld32(lane_addr + (g == 0 ? KTM_ACC1 : KTM_ACC2) + h * 64 + ch * 32, t);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    *reinterpret_cast<float4*>(stg + lane * 32 + ((k ^ (lane & 7)) << 2)) = make_float4(t[4 * k], t[4 * k + 1], t[4 * k + 2], t[4 * k + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + (lane >> 3);
                    const float4 x = *reinterpret_cast<const float4*>(stg + rr * 32 + (((lane & 7) ^ (rr & 7)) << 2));
                    if (r0 + q * 32 + rr < T) *reinterpret_cast<float4*>(dst + (int64_t)rr * ld_d + ch * 32 + (lane & 7) * 4) = x;
                }
                __syncwarp();
            }
        }
    }
    TR(4, 0);
    tc_fence_before();
    __syncthreads();
    TR(5, 0);
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// delta[row] = o[row] . dO[row] for the rows t >= begin of every sample (rows = B * (T - begin))
__global__ void __launch_bounds__(256) delta_kernel(float* __restrict__ delta, const float* __restrict__ o, const float* __restrict__ d_o, int64_t rows,
                                                    int T, int begin) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int w = T - begin;
    for (int64_t i = warp; i < rows; i += nwarps) {
        const int64_t r = (i / w) * T + begin + (i % w);
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

// (batch, rows, cols) bf16 tensor, row stride `row_stride_elems`, samples `rows` rows apart; box = box_rows x box_cols of one sample.
// Out-of-range rows / columns of a box read as zero.
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows,
                      uint32_t box_cols) {
    EncodeTiledFn fn = encode_fn();
    DIP_REQUIRE(fn, "cuTensorMapEncodeTiled entry point not available");
    DIP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (row_stride_elems * 2) % 16 == 0, "tensor map: base / row stride must be 16-byte aligned");
    cuuint64_t gdim[3] = {cols, rows, batch};
    cuuint64_t gstr[2] = {row_stride_elems * 2, rows * row_stride_elems * 2};
    cuuint32_t box[3] = {box_cols, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DIP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed with code %d", (int)r);
    return 0;
}

}  // namespace dip

using namespace dip;

extern "C" {

size_t dip_attn_fwd_tc_state_bytes(int B, int prefix_keys) {
    if (B <= 0 || prefix_keys < tcattn::BM) return 0;
    return (size_t)B * (size_t)(prefix_keys / tcattn::BM * tcattn::BM) * tcattn::STATE_W * sizeof(float);
}

int dip_attn_fwd_tc(float* o, float* lse, const void* qkv_hi, const void* qkv_lo, const uint8_t* key_mask, int64_t mask_stride, int B, int T,
                    int q_begin, float* prefix_state, int prefix_keys, int prefix_mode, dip_stream_t stream) {
    DIP_REQUIRE(qkv_hi && qkv_lo && B > 0 && T > 0, "dip_attn_fwd_tc: null/empty argument");
    DIP_REQUIRE(q_begin >= 0 && q_begin < T, "dip_attn_fwd_tc: q_begin %d outside [0, T)", q_begin);
    DIP_REQUIRE(B <= 65535, "dip_attn_fwd_tc: B too large for one launch");
    DIP_REQUIRE(prefix_mode >= 0 && prefix_mode <= 2, "dip_attn_fwd_tc: prefix_mode must be 0, 1 (save) or 2 (use)");
    const int hoist_rows = prefix_keys / tcattn::BM * tcattn::BM, j_split = prefix_keys / tcattn::BN;
    if (prefix_mode != 0) {
        DIP_REQUIRE(prefix_state && q_begin == 0 && prefix_keys > 0 && prefix_keys < T && hoist_rows > 0 && j_split > 0 && j_split < cdiv(T, tcattn::BN),
                    "dip_attn_fwd_tc: key-prefix hoisting needs a state buffer, q_begin == 0 and 128 <= prefix_keys < T");
        DIP_REQUIRE((reinterpret_cast<uintptr_t>(prefix_state) & 15) == 0, "dip_attn_fwd_tc: prefix_state must be 16-byte aligned");
    }
    DIP_REQUIRE(prefix_mode == 1 || o, "dip_attn_fwd_tc: null output");
    static bool attr = false;
    if (!attr) {
        DIP_CUDA(cudaFuncSetAttribute(tcattn::attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcattn::FWD_SMEM));
        attr = true;
    }
    CUtensorMap hi128, lo128, hi64, lo64;
    DIP_TRY(make_tmap_bf16_3d(&hi128, qkv_hi, B, T, 3 * D, 3 * D, 128, 64));
    DIP_TRY(make_tmap_bf16_3d(&lo128, qkv_lo, B, T, 3 * D, 3 * D, 128, 64));
    DIP_TRY(make_tmap_bf16_3d(&hi64, qkv_hi, B, T, 3 * D, 3 * D, 64, 64));
    DIP_TRY(make_tmap_bf16_3d(&lo64, qkv_lo, B, T, 3 * D, 3 * D, 64, 64));
    // work items = (sample, 128-query tile); one persistent CTA per SM walks them (DIP_FWD_PERSIST=0: one CTA per item)
    const int n_qx = prefix_mode == 1 ? hoist_rows / tcattn::BM : cdiv(T - q_begin, tcattn::BM);
    DIP_REQUIRE((int64_t)n_qx * B < (1ll << 31), "dip_attn_fwd_tc: too many work items");
    const int n_items = n_qx * B;
    static int persist = -1;
    if (persist < 0) {
        const char* e = getenv("DIP_FWD_PERSIST");
        persist = (e && e[0] == '0') ? 0 : 1;
    }
    const int grid = persist ? (n_items < sm_count() ? n_items : sm_count()) : n_items;
    DIP_PROF(DIP_PROF_ATTN_FWD, stream);
    tcattn::attn_fwd_tc_kernel<<<grid, tcattn::FWD_THREADS, tcattn::FWD_SMEM, as_stream(stream)>>>(
        hi128, lo128, hi64, lo64, o, lse, key_mask, mask_stride, T, q_begin, prefix_mode == 2 ? prefix_state : nullptr,
        prefix_mode == 1 ? prefix_state : nullptr, j_split, hoist_rows, n_qx, n_items);
    DIP_LAUNCHED();
    return 0;
}

#ifdef DIP_TRACE
int dip_debug_set_trace(unsigned long long* device_buffer) {
    DIP_CUDA(cudaMemcpyToSymbol(tcattn::g_trace, &device_buffer, sizeof(device_buffer)));
    return 0;
}
#endif

size_t dip_attn_bwd_tc_ds_bytes(int B, int T, int do_begin, int k_begin) {
    if (B <= 0 || T <= 0 || do_begin < 0 || do_begin >= T || k_begin < 0 || k_begin >= T) return 0;
    const size_t tq_pad = (size_t)cdiv(T - do_begin, tcattn::QT) * tcattn::QT;
    return 2u * (size_t)B * (size_t)(T - k_begin) * tq_pad * 2u;
}

int dip_attn_bwd_tc(float* dq, float* dk, float* dv, int ld_d, const void* qkv_hi, const void* qkv_lo, const void* do_hi, const void* do_lo,
                    const float* o, const float* d_o, const float* lse, const uint8_t* key_mask, int64_t mask_stride, float* delta_ws, int B, int T,
                    int do_begin, int q_begin, int k_begin, void* ds_ws, size_t ds_ws_bytes, dip_stream_t stream) {
    DIP_REQUIRE(dq && dk && dv && qkv_hi && qkv_lo && do_hi && do_lo && o && d_o && lse && delta_ws && B > 0 && T > 0,
                "dip_attn_bwd_tc: null/empty argument");
    DIP_REQUIRE(ld_d % 8 == 0 && B <= 65535 && ((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 31) == 0,
                "dip_attn_bwd_tc: ld_d must be a multiple of 8 and dq / dk / dv 32-byte aligned (256-bit stores)");
    DIP_REQUIRE(0 <= do_begin && do_begin <= q_begin && q_begin < T && 0 <= k_begin && k_begin < T,
                "dip_attn_bwd_tc: need 0 <= do_begin <= q_begin < T and 0 <= k_begin < T (got %d, %d, %d)", do_begin, q_begin, k_begin);
    static bool attr = false;
    if (!attr) {
        DIP_CUDA(cudaFuncSetAttribute(tcattn::attn_bwd_kv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcattn::KV_SMEM));
        DIP_CUDA(cudaFuncSetAttribute(tcattn::attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcattn::DQ_SMEM));
        DIP_CUDA(cudaFuncSetAttribute(tcattn::attn_bwd_dq_stream_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcattn::SQ_SMEM));
        attr = true;
    }
    // with a dS workspace the key-stationary kernel stores dS^T and dQ is a streaming GEMM over it (no score recomputation)
    const size_t ds_need = dip_attn_bwd_tc_ds_bytes(B, T, do_begin, k_begin);
    const bool stream_dq = ds_ws != nullptr && ds_ws_bytes >= ds_need && ds_need > 0;
    DIP_REQUIRE(ds_ws == nullptr || (reinterpret_cast<uintptr_t>(ds_ws) & 127) == 0, "dip_attn_bwd_tc: dS workspace must be 128-byte aligned");
    const int tq_pad = cdiv(T - do_begin, tcattn::QT) * tcattn::QT;
    uint8_t* ds_hi = stream_dq ? reinterpret_cast<uint8_t*>(ds_ws) : nullptr;
    uint8_t* ds_lo = stream_dq ? ds_hi + ds_need / 2 : nullptr;
    cudaStream_t st = as_stream(stream);
    CUtensorMap q_hi128, q_lo128;
    DIP_TRY(make_tmap_bf16_3d(&q_hi128, qkv_hi, B, T, 3 * D, 3 * D, 128, 64));
    DIP_TRY(make_tmap_bf16_3d(&q_lo128, qkv_lo, B, T, 3 * D, 3 * D, 128, 64));
    CUtensorMap q_hi64, q_lo64;
    DIP_TRY(make_tmap_bf16_3d(&q_hi64, qkv_hi, B, T, 3 * D, 3 * D, 64, 64));
    DIP_TRY(make_tmap_bf16_3d(&q_lo64, qkv_lo, B, T, 3 * D, 3 * D, 64, 64));
    CUtensorMap d_hi64, d_lo64;
    DIP_TRY(make_tmap_bf16_3d(&d_hi64, do_hi, B, T, D, D, 64, 64));
    DIP_TRY(make_tmap_bf16_3d(&d_lo64, do_lo, B, T, D, D, 64, 64));
    {
        DIP_PROF(DIP_PROF_OTHER, stream);
        const int64_t rows = (int64_t)B * (T - do_begin);
        int64_t blocks = (rows + 7) / 8, cap = (int64_t)sm_count() * 8;
        tcattn::delta_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(delta_ws, o, d_o, rows, T, do_begin);
        DIP_LAUNCHED();
    }
    {
        dim3 grid(cdiv(T - k_begin, 128), B);      // keys k_begin.. stationary, queries do_begin.. streamed
        DIP_PROF(DIP_PROF_ATTN_BWD_DKDV, stream);
        tcattn::attn_bwd_kv_tc_kernel<<<grid, tcattn::NTHREADS, tcattn::KV_SMEM, st>>>(q_hi128, q_lo128, q_hi64, q_lo64, d_hi64, d_lo64, dv, dk, ld_d, lse,
                                                                                       delta_ws, key_mask, mask_stride, T, k_begin, do_begin, ds_hi, ds_lo, tq_pad,
                                                                                       q_begin);
        DIP_LAUNCHED();
    }
    dim3 grid(cdiv(T - q_begin, 128), B);          // 128 stationary queries per CTA
    if (stream_dq) {
        const int c_start = (q_begin - do_begin) / tcattn::QT * tcattn::QT;
        dim3 sgrid(cdiv(T - do_begin - c_start, 128), B);
        CUtensorMap s_hi, s_lo;
        DIP_TRY(make_tmap_bf16_3d(&s_hi, ds_hi, B, T - k_begin, tq_pad, tq_pad, 64, 64));
        DIP_TRY(make_tmap_bf16_3d(&s_lo, ds_lo, B, T - k_begin, tq_pad, tq_pad, 64, 64));
        {
            DIP_PROF(DIP_PROF_ATTN_BWD_DQ, stream);    // keys k_begin .. T-1 from the stored dS
            tcattn::attn_bwd_dq_stream_tc_kernel<<<sgrid, tcattn::NTHREADS, tcattn::SQ_SMEM, st>>>(s_hi, s_lo, q_hi64, q_lo64, dq, ld_d, T, q_begin, k_begin,
                                                                                                   do_begin, c_start);
            DIP_LAUNCHED();
        }
        if (k_begin > 0) {                             // keys below k_begin have no stationary CTA (their dK / dV are not wanted): recompute path, added on top
            DIP_PROF(DIP_PROF_ATTN_BWD_DQ, stream);
            tcattn::attn_bwd_dq_tc_kernel<<<grid, tcattn::NTHREADS, tcattn::DQ_SMEM, st>>>(q_hi128, q_lo128, q_hi64, q_lo64, reinterpret_cast<const uint4*>(do_hi),
                                                                                           reinterpret_cast<const uint4*>(do_lo), dq, ld_d, lse, delta_ws, key_mask,
                                                                                           mask_stride, T, q_begin, k_begin, 1);
            DIP_LAUNCHED();
        }
    } else {
        DIP_PROF(DIP_PROF_ATTN_BWD_DQ, stream);        // queries q_begin.. stationary, all keys streamed, S and dP recomputed
        tcattn::attn_bwd_dq_tc_kernel<<<grid, tcattn::NTHREADS, tcattn::DQ_SMEM, st>>>(q_hi128, q_lo128, q_hi64, q_lo64, reinterpret_cast<const uint4*>(do_hi),
                                                                                       reinterpret_cast<const uint4*>(do_lo), dq, ld_d, lse, delta_ws, key_mask,
                                                                                       mask_stride, T, q_begin, T, 0);
        DIP_LAUNCHED();
    }
    return 0;
}

}  // extern "C"
