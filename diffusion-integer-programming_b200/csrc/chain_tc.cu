// The row-wise half of a ResidualAttentionBlock as ONE kernel (model/modules.py:293-296 after the attention itself):
//
//     forward :  x_mid = x_in + attn . W_out^T + b_out ;  u = LN2(x_mid) . W_fc^T + b_fc ;  x_out = x_mid + QuickGELU(u) . W_proj^T + b_proj
//     backward:  dU = (dx_out . W_proj) * gelu'(u) ;  dH = dU . W_fc ;  dx_mid = dx_out + LN2'(dH) ;  d_attn = dx_mid . W_out   (input gradients only)
//
// As three separate Linear launches (+ LayerNorm kernels in the backward) every stage reads a 512-byte row and writes one: 9 row passes
// forward and 12 backward, all HBM-bound (32 FLOP per byte).  With D = 128 a 128-row tile carries whole feature vectors, so the chain can
// stay on chip: here a tile makes ONE trip -- 2 rows in, the rows the backward pass needs (x_mid, u) and x_out out.
//
// Layout of the work (one persistent CTA per SM, 18 warps):
//   * the three weights (bf16 hi/lo planes, 64 KB each) are resident in shared memory (TMA, SWIZZLE_128B) -- that is 192 of the 227 KB, so
//     the A operands cannot be staged there: they live in TENSOR MEMORY (tcgen05.mma with the A operand in TMEM: lane = row, a 32-bit column
//     = a bf16 pair), written by the same threads that own the rows;
//   * two independent GROUPS of 8 row warps, each with its own tile, TMEM slot (A planes 128 columns + accumulator 128) and issuer warp:
//     while one group waits for memory or for the tensor pipe the other computes.  Row thread = (row, column half);
//   * every stage walks its 64 columns in four 16-column CHUNKS inside a rolled loop: accumulator chunk out of TMEM, bias / residual /
//     activation in registers, next A operand chunk back into TMEM.  The first version unrolled whole rows (64 values per thread): 170 KB of
//     straight-line code per kernel, 54 % of the warp stalls were instruction fetch (ncu `stall_no_inst`; the instruction cache holds
//     32 KB).  Values a later pass of the same stage needs again (x_mid for the LayerNorm statistics, dH . gamma and x_hat for LayerNorm')
//     are parked in TMEM columns that are free at that moment instead of in registers;
//   * A-plane layout: k-step j (16 features) has its hi words in columns 16j .. 16j+7 and its lo words in 16j+8 .. 16j+15, so the fp32
//     chunk a thread parks in the A region is replaced IN PLACE by the planes made from it;
//   * global rows are moved by their owner thread as 256-bit accesses (32 rows x 32 bytes per warp instruction, whole sectors); the loads
//     of the next chunk are in flight while the current one is worked on, and the first chunk of a stage is requested before the wait for
//     the product it is combined with;
//   * forward: the residual rows are not added after the product but PRELOADED into the accumulator (x_in + b_out before the first
//     product, x_mid + b_proj before the third), which the MMAs then accumulate onto: the residual loads join the stream of stage-0
//     loads instead of sitting between a product and the LayerNorm.
// Measured (B200, 148 000 rows): forward 0.14 ms, backward 0.146 ms = 2.7 / 3.1 TB/s of algorithmic row traffic; the three separate
// Linear launches they replace take 0.21 ms (+ LayerNorm kernels in the backward).  Tried and dropped: per-thread
// cp.async.bulk.prefetch.L2 of the next tile's rows (12 % slower: 1 500 small bulk operations per tile pair).
// Numerics: "bf16x2 split" like gemm_tc.cu (tc_common.cuh), every product 3 passes; LayerNorm in fp32 (shifted one-pass sums per half
// row, pairwise combination, rsqrt + one Newton step as in the GEMM loader).
#include <string.h>
#include "tc_common.cuh"

namespace dip {
namespace tcchain {

using namespace tc;

constexpr int GROUPS = 2, GROUP_WARPS = 8, GROUP_THREADS = GROUP_WARPS * 32, ROW_WARPS = GROUPS * GROUP_WARPS, NT = (ROW_WARPS + GROUPS) * 32;
constexpr uint32_t W_BLOCK = 16384, W_PLANE = 32768, W_ONE = 65536;          // [plane][k-block][128 rows][64] bf16
constexpr uint32_t OFF_XCH = 3 * W_ONE, XCH_WARP = 256, OFF_PAR = OFF_XCH + ROW_WARPS * XCH_WARP, OFF_BAR = OFF_PAR + 5 * 512;   // per-warp LayerNorm exchange
constexpr uint32_t SMEM_BYTES = OFF_BAR + 64;
static_assert(SMEM_BYTES <= 232448, "chain kernel shared memory");
constexpr uint32_t TM_A = 0, TM_ACC = 128, TM_SLOT = 256;                     // per slot: A planes (k-step j: hi 16j, lo 16j+8), accumulator 128
constexpr uint32_t IDESC = instr_desc(128, 128, 0, 0);
enum { B_W = 0, B_AFULL = 1, B_CFULL = 3, B_STAGGER = 5 };
enum { PAR_B0 = 0, PAR_B1 = 128, PAR_B2 = 256, PAR_G = 384, PAR_BETA = 512 };  // per-column parameter vectors in shared memory (float offsets)

struct FastDiv {
    uint32_t d, mul, shr;
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1u ? n : (__umulhi(n, mul) >> shr); }
};
static FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d; f.mul = 0; f.shr = 0;
    if (d > 1) {
        int lg = 0;
        while ((1ull << lg) < d) ++lg;
        const int p = 31 + lg;
        f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
        f.shr = (uint32_t)(p - 32);
    }
    return f;
}

struct Params {
    int n_live, n_tiles;                 // live rows (tokens t >= win of every sample) and 128-row tiles over them
    FastDiv live_T, full_T;              // T - win, T
    uint32_t T, win;
    // forward: a = attn, r_* = x_in (two-source);  backward: a = dx_out, u_in = u, xm_in = x_mid
    const float* a;
    const float *r_shared, *r_batch; uint32_t r_split; int64_t r_sstride;
    const float *b0, *b1, *b2;           // biases of the three Linears (forward)
    const float *ln_g, *ln_b;
    float *xmid, *mean2, *rstd2, *u, *xout;          // forward outputs (stats / u nullable)
    const float *u_in, *xm_in, *mean_in, *rstd_in;   // backward inputs
    float *dxm, *dao; void *do_hi, *do_lo;           // backward outputs (planes nullable)
    const float* dres;                               // MODE 2: residual gradient added to the result (a = dQKV rows of 384, dxm = result)
    int w_col0[3];                                   // first column of weight w inside its tensor map (MODE 2: 0, 128, 256 of W_in^T)
};

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_sigmoid(float z) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * z)); }

// live row index -> global row (b*T + t)
__device__ __forceinline__ uint32_t grow(const Params& P, uint32_t i) {
    const uint32_t b = P.live_T.div(i);
    return b * P.T + P.win + (i - b * P.live_T.d);
}
__device__ __forceinline__ const float* rsrc_row(const Params& P, uint32_t g) {
    if (P.r_split == 0) return P.r_batch + (size_t)g * D;
    const uint32_t b = P.full_T.div(g), t = g - b * P.T;
    return t < P.r_split ? P.r_shared + (int64_t)b * P.r_sstride + (size_t)t * D : P.r_batch + ((size_t)b * (P.T - P.r_split) + (t - P.r_split)) * D;
}

// ---- row I/O.  A thread owns a row (its TMEM lane), so it moves its own 16-column chunk as two 256-bit accesses: a warp instruction
// touches 32 rows x 32 bytes = 32 whole sectors.  That is as many L1 wavefronts per byte as the staged alternative (8 rows x 64 bytes per
// instruction + a shared-memory transpose: 32 + 32 wavefronts per 2 KB chunk against 2 x 32 here) without its STS -> __syncwarp -> LDS
// chain, which next to a tensor pipe that streams its B operand out of shared memory cost hundreds of cycles per chunk. --------------
__device__ __forceinline__ void ld8_stream(float* v, const float* p) {      // read-once inputs: no L1 allocation
    asm volatile("ld.global.L1::no_allocate.L2::256B.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void ld8(float* v, const float* p) {             // coherent: rows this thread wrote earlier in the kernel
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void st8(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]),
                 "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld16_stream(float (&v)[16], const float* p) { ld8_stream(v, p); ld8_stream(v + 8, p + 8); }
__device__ __forceinline__ void ld16(float (&v)[16], const float* p) { ld8(v, p); ld8(v + 8, p + 8); }
__device__ __forceinline__ void st16(float* p, const float (&v)[16], bool ok) {
    if (ok) { st8(p, v); st8(p + 8, v + 8); }
}
__device__ __forceinline__ void copy16(float (&d)[16], const float (&s)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = s[i];
}

// tcgen05.ld without the wait, and the wait tied to the registers it guards: a chunk loop asks for the next accumulator chunk before it
// works on the current one (next to a tensor pipe busy with the other group's product a tcgen05.ld takes ~500 cycles to come back)
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, float (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]),
          "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(float (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]), "+f"(v[10]),
                   "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :
                 : "memory");
}

__device__ __forceinline__ void tmem_st16f(uint32_t taddr, const float (&x)[16]) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = __float_as_uint(x[i]);
    tmem_st16(taddr, w);
}
// one fp32 chunk (16 features = one k-step) -> its packed bf16 hi words (8 columns) and lo words (8 columns) in the A region
__device__ __forceinline__ void put_chunk(uint32_t taddr, const float (&x)[16]) {
    uint32_t w[16];
#pragma unroll
    for (int c = 0; c < 8; ++c) split2(x[2 * c], x[2 * c + 1], w[c], w[8 + c]);
    tmem_st16(taddr, w);
}

// 24 MMAs of one product: A (TMEM planes of the slot) . W^T (resident weight w), fp32 accumulate into the slot's accumulator
// (acc0 = 1: on top of what the row threads preloaded into the accumulator)
__device__ __forceinline__ void issue_product(uint32_t tmem_slot_addr, uint64_t dW, int w, uint32_t acc0) {
    const uint64_t dw = desc_adv(dW, w * W_ONE);
    uint32_t acc = acc0;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass)                                 // A hi, hi, lo x W hi, lo, hi
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                tc_mma_ts(tmem_slot_addr + TM_ACC, tmem_slot_addr + TM_A + (kb * 4 + ks) * 16 + (pass == 2 ? 8 : 0),
                          desc_adv(dw, (pass == 1 ? W_PLANE : 0) + kb * W_BLOCK + ks * 32), IDESC, acc);
                acc = 1;
            }
}

// MODE 0: forward chain (weights out_w, fc_w, proj_w);  MODE 1: input-gradient chain (weights proj_wT, fc_wT, out_wT);
// MODE 2: the input gradient of LayerNorm 1 + in-projection, the last piece of a block's backward pass:
//     dx_in = dx_mid + LN1'( dQKV . W_in )          (model/modules.py:288-291 under autograd, input gradients only)
// the three 128-column blocks of dQ|dK|dV are three products into ONE accumulator (K = 384: the three resident "weights" are the
// column blocks of W_in^T), the LayerNorm backward is the epilogue -- instead of a K = 384 Linear launch, a 512-byte row written and
// read back, and a LayerNorm launch.
template <int MODE>
__global__ void __launch_bounds__(NT, 1)
chain_tc_kernel(const __grid_constant__ CUtensorMap tm_w0_hi, const __grid_constant__ CUtensorMap tm_w0_lo, const __grid_constant__ CUtensorMap tm_w1_hi,
                const __grid_constant__ CUtensorMap tm_w1_lo, const __grid_constant__ CUtensorMap tm_w2_hi, const __grid_constant__ CUtensorMap tm_w2_lo,
                const Params P) {
    constexpr bool BWD = MODE == 1;
    extern __shared__ __align__(1024) uint8_t sm[];
    if ((smem_u32(sm) & 1023u) != 0u) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
    float* par = reinterpret_cast<float*>(sm + OFF_PAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(&bars[B_W], 1);
        for (int s = 0; s < GROUPS; ++s) { mbar_init(&bars[B_AFULL + s], GROUP_THREADS); mbar_init(&bars[B_CFULL + s], 1); }
        mbar_init(&bars[B_STAGGER], GROUP_THREADS);
        mbar_fence_init();
    }
    if (warp == ROW_WARPS) tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < 5 * 128; i += NT) {
        const float* src = i < PAR_B1 ? P.b0 : i < PAR_B2 ? P.b1 : i < PAR_G ? P.b2 : i < PAR_BETA ? P.ln_g : P.ln_b;
        par[i] = src ? src[i & 127] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int n_my = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // tiles of this CTA: blockIdx.x + k * gridDim.x

    if (warp >= ROW_WARPS) {
        // ------------------------------------------------------------------ issuer of group g: weights (once), then the three products of every tile
        const int g = warp - ROW_WARPS;
        if (g == 0 && lane == 0) {
            mbar_expect_tx(&bars[B_W], 3 * W_ONE);
            const CUtensorMap* maps[6] = {&tm_w0_hi, &tm_w0_lo, &tm_w1_hi, &tm_w1_lo, &tm_w2_hi, &tm_w2_lo};
            for (int w = 0; w < 3; ++w)
                for (int pl = 0; pl < 2; ++pl)
                    for (int kb = 0; kb < 2; ++kb) tma_load_2d(sm + w * W_ONE + pl * W_PLANE + kb * W_BLOCK, maps[2 * w + pl], &bars[B_W], P.w_col0[w] + kb * 64, 0);
        }
        __syncwarp();
        const uint64_t dW = smem_desc(smem_u32(sm), 16, 1024);
        mbar_wait(&bars[B_W], 0);
        uint32_t ph = 0;
        for (int k = g; k < n_my; k += GROUPS)
            for (int w = 0; w < 3; ++w) {
                mbar_wait(&bars[B_AFULL + g], ph & 1);
                ph++;
                tc_fence_after();
                if (elect_one()) {
                    issue_product(tmem + g * TM_SLOT, dW, w, MODE == 0 ? (w != 1 ? 1u : 0u) : MODE == 2 ? (w > 0 ? 1u : 0u) : 0u);
                    tc_commit(&bars[B_CFULL + g]);
                }
                __syncwarp();
            }
    } else {
        // ------------------------------------------------------------------ row warps: thread = (row, column half) of its group's tile
        const int g = warp >> 3, wg = warp & 7, q = wg & 3, h = wg >> 2;
        float* xch = reinterpret_cast<float*>(sm + OFF_XCH + warp * XCH_WARP);
        const float* xch_partner = reinterpret_cast<const float*>(sm + OFF_XCH + (warp ^ 4) * XCH_WARP);   // same rows, other column half
        const uint32_t slot = tmem + ((uint32_t)(q * 32) << 16) + g * TM_SLOT;
        const int c0 = h * 64;                                                                     // this thread's columns: c0 .. c0 + 63
        const uint32_t a_reg = slot + TM_A + c0, acc_reg = slot + TM_ACC + c0;                      // chunk ch at + 16 ch in both regions
        const int bar_id = 1 + g;
        uint32_t cph = 0;                                                                          // CFULL phase
        // The two groups run the same sequence of memory-heavy and compute-heavy stages; started together they stay in step and the
        // stage times ADD (measured: row work 44 us + products 21 + loads 36 + stores 44 = the 140 us of a 148 000-row launch, no
        // overlap).  Group 1 therefore starts half a tile late: when group 0 has handed over its second operand.
        if (g == 1 && n_my > 0) mbar_wait(&bars[B_STAGGER], 0);
        for (int k = g; k < n_my; k += GROUPS) {
            const uint32_t live = (uint32_t)(blockIdx.x + k * gridDim.x) * 128u + q * 32 + lane;    // this thread's row
            const bool ok = live < (uint32_t)P.n_live;
            const uint32_t g_own = grow(P, min(live, (uint32_t)P.n_live - 1u));
            const size_t off = (size_t)g_own * D + c0;                                              // its 64 columns in every (B*T,128) array
            if constexpr (MODE == 2) {
                const float* ap = P.a + (size_t)g_own * (3 * D) + c0;
                float nx[16], ny[16];
#pragma unroll 1
                for (int p = 0; p < 3; ++p) {
                    ld16_stream(nx, ap + D * p);
                    ld16_stream(ny, ap + D * p + 16);
                    if (p > 0) {                                          // the A region is free once the previous product has completed
                        mbar_wait(&bars[B_CFULL + g], cph & 1);
                        cph++;
                        tc_fence_after();
                    }
#pragma unroll 1
                    for (int ch = 0; ch < 4; ch += 2) {
                        float v[16];
                        copy16(v, nx);
                        if (ch < 2) ld16_stream(nx, ap + D * p + 16 * (ch + 2));
                        put_chunk(a_reg + 16 * ch, v);
                        copy16(v, ny);
                        if (ch < 2) ld16_stream(ny, ap + D * p + 16 * (ch + 3));
                        put_chunk(a_reg + 16 * (ch + 1), v);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&bars[B_AFULL + g]);
                    if (g == 0 && k == 0 && p == 1) mbar_arrive(&bars[B_STAGGER]);
                }
                // epilogue: dx_in = dres + LN1'(dH), dH in the accumulator;  gg = dH * gamma ;  dx = rstd * (gg - mean(gg) - xhat * mean(gg * xhat))
                const float* xp = rsrc_row(P, g_own) + c0;
                ld16_stream(nx, xp);
                const float mu = P.mean_in[g_own], rs = P.rstd_in[g_own];
                mbar_wait(&bars[B_CFULL + g], cph & 1);
                cph++;
                tc_fence_after();
                float s1 = 0.f, s2 = 0.f;
                float nacc[16];
                tmem_ld16_async(acc_reg, nacc);
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    float xh[16];
                    copy16(xh, nx);
                    if (ch < 3) ld16_stream(nx, xp + 16 * (ch + 1));
                    float gg[16];
                    tmem_ld_wait16(nacc);
                    copy16(gg, nacc);
                    if (ch < 3) tmem_ld16_async(acc_reg + 16 * (ch + 1), nacc);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        xh[i] = (xh[i] - mu) * rs;
                        gg[i] *= par[PAR_G + c0 + 16 * ch + i];
                        s1 += gg[i];
                        s2 += gg[i] * xh[i];
                    }
                    tmem_st16f(acc_reg + 16 * ch, gg);
                    tmem_st16f(a_reg + 16 * ch, xh);
                }
                tmem_st_wait();
                const float* dp = P.dres + off;
                ld16_stream(nx, dp);
                xch[lane] = s1;
                xch[32 + lane] = s2;
                named_bar_sync(bar_id, GROUP_THREADS);
                const float m1 = (s1 + xch_partner[lane]) * (1.0f / D);
                const float m2 = (s2 + xch_partner[32 + lane]) * (1.0f / D);
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    float x[16];
                    copy16(x, nx);
                    if (ch < 3) ld16_stream(nx, dp + 16 * (ch + 1));
                    float gg[16], xh[16];
                    tmem_ld16_async(acc_reg + 16 * ch, gg);
                    tmem_ld16_async(a_reg + 16 * ch, xh);
                    tmem_ld_wait16(gg);
                    tmem_ld_wait16(xh);
#pragma unroll
                    for (int i = 0; i < 16; ++i) x[i] += rs * (gg[i] - m1 - xh[i] * m2);
                    st16(P.dxm + off + 16 * ch, x, ok);
                }
                tc_fence_before();
                continue;
            }
            float nx[16];
            // ---------------- stage 0: the first A operand (forward: attention output rows, backward: dx_out rows).  Forward: the accumulator
            // is preloaded with x_in + b_out, so the first product (issued on top of it) leaves x_mid itself and the residual rows are
            // fetched here, in one stream with the attention rows, instead of after the wait for the product.
            if (!BWD) {
                const float* ap = P.a + off;
                const float* rp = rsrc_row(P, g_own) + c0;
                float nr[16];
                ld16_stream(nx, ap);
                ld16_stream(nr, rp);
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    float v[16];
                    copy16(v, nx);
                    if (ch < 3) ld16_stream(nx, ap + 16 * (ch + 1));
                    put_chunk(a_reg + 16 * ch, v);
                    copy16(v, nr);
                    if (ch < 3) ld16_stream(nr, rp + 16 * (ch + 1));
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += par[PAR_B0 + c0 + 16 * ch + i];
                    tmem_st16f(acc_reg + 16 * ch, v);
                }
            } else {
                const float* ap = P.a + off;
                float ny[16];
                ld16_stream(nx, ap);
                ld16_stream(ny, ap + 16);
#pragma unroll 1
                for (int ch = 0; ch < 4; ch += 2) {
                    float v[16];
                    copy16(v, nx);
                    if (ch < 2) ld16_stream(nx, ap + 16 * (ch + 2));
                    put_chunk(a_reg + 16 * ch, v);
                    copy16(v, ny);
                    if (ch < 2) ld16_stream(ny, ap + 16 * (ch + 3));
                    put_chunk(a_reg + 16 * (ch + 1), v);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars[B_AFULL + g]);
            // ---------------- stage 1
            if (!BWD) {
                // x_mid = attn . W_out^T + (b_out + x_in) comes out of the accumulator ; LayerNorm 2 -> A
                mbar_wait(&bars[B_CFULL + g], cph & 1);
                cph++;
                tc_fence_after();
                // statistics in ONE pass over the row (the accumulator is read twice, not three times): sums of x - s and (x - s)^2 with the
                // shift s = the half-row's first element (the textbook shifted-data form: no cancellation unless |mean - s| >> std), the two
                // halves of a row combined by the pairwise update  M2 = M2_a + M2_b + (mean_a - mean_b)^2 * 64 * 64 / 128
                float s1 = 0.f, s2 = 0.f, shift = 0.f;
                float nacc[16];
                tmem_ld16_async(acc_reg, nacc);
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    float x[16];
                    tmem_ld_wait16(nacc);
                    copy16(x, nacc);
                    if (ch < 3) tmem_ld16_async(acc_reg + 16 * (ch + 1), nacc);
                    if (ch == 0) shift = x[0];
#pragma unroll
                    for (int i = 0; i < 16; ++i) { const float d = x[i] - shift; s1 += d; s2 = fmaf(d, d, s2); }
                    st16(P.xmid + off + 16 * ch, x, ok);
                }
                const float mean_h = shift + s1 * (1.0f / 64.0f), m2_h = s2 - s1 * s1 * (1.0f / 64.0f);
                xch[lane] = mean_h;
                xch[32 + lane] = m2_h;
                tmem_ld16_async(acc_reg, nacc);
                named_bar_sync(bar_id, GROUP_THREADS);
                const float mean_p = xch_partner[lane], m2_p = xch_partner[32 + lane];
                const float mu = 0.5f * (mean_h + mean_p);
                const float dm = mean_h - mean_p;
                const float var = fmaxf((m2_h + m2_p + 32.0f * dm * dm) * (1.0f / D), 0.f);
                float rs = rsqrtf(var + 1e-5f);
                rs = rs * (1.5f - 0.5f * (var + 1e-5f) * rs * rs);
                if (P.mean2 && h == 0 && ok) { P.mean2[g_own] = mu; P.rstd2[g_own] = rs; }
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    float x[16];
                    tmem_ld_wait16(nacc);
                    copy16(x, nacc);
                    if (ch < 3) tmem_ld16_async(acc_reg + 16 * (ch + 1), nacc);
#pragma unroll
                    for (int i = 0; i < 16; ++i) x[i] = (x[i] - mu) * rs * par[PAR_G + c0 + 16 * ch + i] + par[PAR_BETA + c0 + 16 * ch + i];
                    put_chunk(a_reg + 16 * ch, x);
                }
            } else {
                // dU = (dx_out . W_proj) * gelu'(u) -> A
                const float* up = P.u_in + off;
                ld16_stream(nx, up);
                mbar_wait(&bars[B_CFULL + g], cph & 1);
                cph++;
                tc_fence_after();
                float nacc[16];
                tmem_ld16_async(acc_reg, nacc);
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    float uu[16];
                    copy16(uu, nx);
                    if (ch < 3) ld16_stream(nx, up + 16 * (ch + 1));
                    float acc[16];
                    tmem_ld_wait16(nacc);
                    copy16(acc, nacc);
                    if (ch < 3) tmem_ld16_async(acc_reg + 16 * (ch + 1), nacc);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float sg = fast_sigmoid(1.702f * uu[i]);
                        acc[i] *= sg * (1.0f + 1.702f * uu[i] * (1.0f - sg));
                    }
                    put_chunk(a_reg + 16 * ch, acc);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars[B_AFULL + g]);
            if (g == 0 && k == 0) mbar_arrive(&bars[B_STAGGER]);
            // ---------------- stage 2
            if (!BWD) {
                // u = LN2(x_mid) . W_fc^T + b_fc (saved) ; QuickGELU -> A ; the accumulator chunk just read is preloaded with x_mid + b_proj
                // for the last product (x_mid re-read: this thread wrote it in stage 1, it is L2-resident and requested before the wait)
                const float* xp = P.xmid + off;
                ld16(nx, xp);
                mbar_wait(&bars[B_CFULL + g], cph & 1);
                cph++;
                tc_fence_after();
                float nacc[16];
                tmem_ld16_async(acc_reg, nacc);
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    float x[16];
                    tmem_ld_wait16(nacc);
                    copy16(x, nacc);
                    if (ch < 3) tmem_ld16_async(acc_reg + 16 * (ch + 1), nacc);
#pragma unroll
                    for (int i = 0; i < 16; ++i) x[i] += par[PAR_B1 + c0 + 16 * ch + i];
                    if (P.u) st16(P.u + off + 16 * ch, x, ok);
#pragma unroll
                    for (int i = 0; i < 16; ++i) x[i] *= fast_sigmoid(1.702f * x[i]);
                    put_chunk(a_reg + 16 * ch, x);
                    copy16(x, nx);
                    if (ch < 3) ld16(nx, xp + 16 * (ch + 1));
#pragma unroll
                    for (int i = 0; i < 16; ++i) x[i] += par[PAR_B2 + c0 + 16 * ch + i];
                    tmem_st16f(acc_reg + 16 * ch, x);
                }
            } else {
                // dx_mid = dx_out + LN2'(dH):  gg = dH * gamma ;  dx = rstd * (gg - mean(gg) - xhat * mean(gg * xhat))
                const float* xp = P.xm_in + off;
                ld16_stream(nx, xp);
                const float mu = P.mean_in[g_own], rs = P.rstd_in[g_own];
                mbar_wait(&bars[B_CFULL + g], cph & 1);
                cph++;
                tc_fence_after();
                float s1 = 0.f, s2 = 0.f;
                float nacc[16];
                tmem_ld16_async(acc_reg, nacc);
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    float xh[16];
                    copy16(xh, nx);
                    if (ch < 3) ld16_stream(nx, xp + 16 * (ch + 1));
                    float gg[16];
                    tmem_ld_wait16(nacc);
                    copy16(gg, nacc);
                    if (ch < 3) tmem_ld16_async(acc_reg + 16 * (ch + 1), nacc);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        xh[i] = (xh[i] - mu) * rs;
                        gg[i] *= par[PAR_G + c0 + 16 * ch + i];
                        s1 += gg[i];
                        s2 += gg[i] * xh[i];
                    }
                    tmem_st16f(acc_reg + 16 * ch, gg);                   // both parked for the second pass: gg in place, xhat in the (free) A region
                    tmem_st16f(a_reg + 16 * ch, xh);
                }
                tmem_st_wait();
                const float* dp = P.a + off;
                ld16_stream(nx, dp);                                     // + dx_out
                xch[lane] = s1;
                xch[32 + lane] = s2;
                named_bar_sync(bar_id, GROUP_THREADS);
                const float m1 = (s1 + xch_partner[lane]) * (1.0f / D);
                const float m2 = (s2 + xch_partner[32 + lane]) * (1.0f / D);
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    float x[16];
                    copy16(x, nx);
                    if (ch < 3) ld16_stream(nx, dp + 16 * (ch + 1));
                    float gg[16], xh[16];
                    tmem_ld16_async(acc_reg + 16 * ch, gg);
                    tmem_ld16_async(a_reg + 16 * ch, xh);
                    tmem_ld_wait16(gg);
                    tmem_ld_wait16(xh);
#pragma unroll
                    for (int i = 0; i < 16; ++i) x[i] += rs * (gg[i] - m1 - xh[i] * m2);
                    st16(P.dxm + off + 16 * ch, x, ok);
                    put_chunk(a_reg + 16 * ch, x);                        // replaces the parked xhat chunk
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars[B_AFULL + g]);
            // ---------------- stage 3
            if (!BWD) {
                // x_out = QuickGELU(u) . W_proj^T + (x_mid + b_proj)
                mbar_wait(&bars[B_CFULL + g], cph & 1);
                cph++;
                tc_fence_after();
                float nacc[16];
                tmem_ld16_async(acc_reg, nacc);
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    float x[16];
                    tmem_ld_wait16(nacc);
                    copy16(x, nacc);
                    if (ch < 3) tmem_ld16_async(acc_reg + 16 * (ch + 1), nacc);
                    st16(P.xout + off + 16 * ch, x, ok);
                }
            } else {
                // d_attn = dx_mid . W_out (+ its bf16 planes for the attention backward)
                mbar_wait(&bars[B_CFULL + g], cph & 1);
                cph++;
                tc_fence_after();
                float nacc[16];
                tmem_ld16_async(acc_reg, nacc);
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    float x[16];
                    tmem_ld_wait16(nacc);
                    copy16(x, nacc);
                    if (ch < 3) tmem_ld16_async(acc_reg + 16 * (ch + 1), nacc);
                    st16(P.dao + off + 16 * ch, x, ok);
                    if (P.do_hi && ok) {
                        float hi[8], lo[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            uint32_t wh, wl;
                            split2(x[2 * c], x[2 * c + 1], wh, wl);
                            hi[c] = __uint_as_float(wh); lo[c] = __uint_as_float(wl);
                        }
                        // 16 bf16 = 32 bytes per row and plane
                        st8(reinterpret_cast<float*>(reinterpret_cast<uint16_t*>(P.do_hi) + off + 16 * ch), hi);
                        st8(reinterpret_cast<float*>(reinterpret_cast<uint16_t*>(P.do_lo) + off + 16 * ch), lo);
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == ROW_WARPS) tmem_dealloc(tmem, 512);
}

}  // namespace tcchain
}  // namespace dip

using namespace dip;

static int chain_launch(int mode, const tcchain::Params& P, const void* const w_hi[3], const void* const w_lo[3], cudaStream_t st) {
    CUtensorMap m[6];
    const uint64_t w_cols = mode == 2 ? 3 * D : D;               // MODE 2: the three weights are column blocks of one (128, 384) matrix
    for (int i = 0; i < 3; ++i) {
        DIP_TRY(make_tmap_bf16_2d(&m[2 * i], w_hi[i], D, w_cols, w_cols, 128, 64));
        DIP_TRY(make_tmap_bf16_2d(&m[2 * i + 1], w_lo[i], D, w_cols, w_cols, 128, 64));
    }
    static bool attr = false;
    if (!attr) {
        DIP_CUDA(cudaFuncSetAttribute(tcchain::chain_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcchain::SMEM_BYTES));
        DIP_CUDA(cudaFuncSetAttribute(tcchain::chain_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcchain::SMEM_BYTES));
        DIP_CUDA(cudaFuncSetAttribute(tcchain::chain_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcchain::SMEM_BYTES));
        attr = true;
    }
    int grid = sm_count();
    if (grid > P.n_tiles) grid = P.n_tiles;
    if (mode == 1) tcchain::chain_tc_kernel<1><<<grid, tcchain::NT, tcchain::SMEM_BYTES, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], P);
    else if (mode == 2) tcchain::chain_tc_kernel<2><<<grid, tcchain::NT, tcchain::SMEM_BYTES, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], P);
    else tcchain::chain_tc_kernel<0><<<grid, tcchain::NT, tcchain::SMEM_BYTES, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], P);
    return 0;
}

static int chain_rows(tcchain::Params& P, int B, int T, int win) {
    DIP_REQUIRE(B > 0 && T > 0 && win >= 0 && win < T, "dip_block_chain: bad sizes B=%d T=%d win=%d", B, T, win);
    DIP_REQUIRE((int64_t)B * T * D < (1ll << 32), "dip_block_chain: B*T*128 must stay below 2^32");
    P.n_live = B * (T - win);
    P.n_tiles = cdiv(P.n_live, 128);
    P.live_T = tcchain::make_fastdiv((uint32_t)(T - win));
    P.full_T = tcchain::make_fastdiv((uint32_t)T);
    P.T = (uint32_t)T; P.win = (uint32_t)win;
    return 0;
}

extern "C" {

int dip_block_chain_fwd_tc(float* xout, float* xmid, float* mean2, float* rstd2, float* u, const float* attn, dip_rows xin, const void* out_w_hi,
                           const void* out_w_lo, const float* out_b, const float* ln2_g, const float* ln2_b, const void* fc_w_hi, const void* fc_w_lo,
                           const float* fc_b, const void* proj_w_hi, const void* proj_w_lo, const float* proj_b, int B, int T, int win,
                           dip_stream_t stream) {
    DIP_REQUIRE(xout && xmid && attn && xin.batch && out_w_hi && out_w_lo && out_b && ln2_g && ln2_b && fc_w_hi && fc_w_lo && fc_b && proj_w_hi &&
                    proj_w_lo && proj_b,
                "dip_block_chain_fwd_tc: null argument");
    DIP_REQUIRE((mean2 == nullptr) == (rstd2 == nullptr), "dip_block_chain_fwd_tc: mean2 and rstd2 go together");
    DIP_REQUIRE(xin.split == 0 || (xin.shared && xin.T == T && xin.split < T), "dip_block_chain_fwd_tc: bad residual row source");
    tcchain::Params P;
    memset(&P, 0, sizeof(P));
    DIP_TRY(chain_rows(P, B, T, win));
    P.a = attn;
    P.r_shared = xin.shared; P.r_batch = xin.batch; P.r_split = (uint32_t)xin.split; P.r_sstride = xin.shared ? xin.shared_stride : 0;
    P.b0 = out_b; P.b1 = fc_b; P.b2 = proj_b; P.ln_g = ln2_g; P.ln_b = ln2_b;
    P.xmid = xmid; P.mean2 = mean2; P.rstd2 = rstd2; P.u = u; P.xout = xout;
    const void* hi[3] = {out_w_hi, fc_w_hi, proj_w_hi};
    const void* lo[3] = {out_w_lo, fc_w_lo, proj_w_lo};
    DIP_PROF(DIP_PROF_GEMM, stream);
    DIP_TRY(chain_launch(0, P, hi, lo, as_stream(stream)));
    DIP_LAUNCHED();
    return 0;
}

int dip_block_chain_bwd_tc(float* dxmid, float* dattn, void* dattn_hi, void* dattn_lo, const float* dxout, const float* u, const float* xmid,
                           const float* mean2, const float* rstd2, const float* ln2_g, const void* proj_wT_hi, const void* proj_wT_lo,
                           const void* fc_wT_hi, const void* fc_wT_lo, const void* out_wT_hi, const void* out_wT_lo, int B, int T, int win,
                           dip_stream_t stream) {
    DIP_REQUIRE(dxmid && dattn && dxout && u && xmid && mean2 && rstd2 && ln2_g && proj_wT_hi && proj_wT_lo && fc_wT_hi && fc_wT_lo && out_wT_hi &&
                    out_wT_lo,
                "dip_block_chain_bwd_tc: null argument");
    DIP_REQUIRE((dattn_hi == nullptr) == (dattn_lo == nullptr), "dip_block_chain_bwd_tc: the two planes go together");
    tcchain::Params P;
    memset(&P, 0, sizeof(P));
    DIP_TRY(chain_rows(P, B, T, win));
    P.a = dxout; P.u_in = u; P.xm_in = xmid; P.mean_in = mean2; P.rstd_in = rstd2; P.ln_g = ln2_g;
    P.dxm = dxmid; P.dao = dattn; P.do_hi = dattn_hi; P.do_lo = dattn_lo;
    const void* hi[3] = {proj_wT_hi, fc_wT_hi, out_wT_hi};
    const void* lo[3] = {proj_wT_lo, fc_wT_lo, out_wT_lo};
    DIP_PROF(DIP_PROF_GEMM, stream);
    DIP_TRY(chain_launch(1, P, hi, lo, as_stream(stream)));
    DIP_LAUNCHED();
    return 0;
}

int dip_block_in_bwd_tc(float* dxin, const float* dqkv, dip_rows xin, const float* mean1, const float* rstd1, const float* ln1_g, const float* dres,
                        const void* in_wT_hi, const void* in_wT_lo, int B, int T, int win, dip_stream_t stream) {
    DIP_REQUIRE(dxin && dqkv && xin.batch && mean1 && rstd1 && ln1_g && dres && in_wT_hi && in_wT_lo, "dip_block_in_bwd_tc: null argument");
    DIP_REQUIRE(xin.split == 0 || (xin.shared && xin.T == T && xin.split < T), "dip_block_in_bwd_tc: bad row source");
    tcchain::Params P;
    memset(&P, 0, sizeof(P));
    DIP_TRY(chain_rows(P, B, T, win));
    DIP_REQUIRE((int64_t)B * T * 3 * D < (1ll << 32), "dip_block_in_bwd_tc: B*T*384 must stay below 2^32");
    P.a = dqkv;
    P.r_shared = xin.shared; P.r_batch = xin.batch; P.r_split = (uint32_t)xin.split; P.r_sstride = xin.shared ? xin.shared_stride : 0;
    P.mean_in = mean1; P.rstd_in = rstd1; P.ln_g = ln1_g; P.dres = dres; P.dxm = dxin;
    P.w_col0[0] = 0; P.w_col0[1] = D; P.w_col0[2] = 2 * D;
    const void* hi[3] = {in_wT_hi, in_wT_hi, in_wT_hi};
    const void* lo[3] = {in_wT_lo, in_wT_lo, in_wT_lo};
    DIP_PROF(DIP_PROF_GEMM, stream);
    DIP_TRY(chain_launch(2, P, hi, lo, as_stream(stream)));
    DIP_LAUNCHED();
    return 0;
}

}  // extern "C"
