// Linear layers on tcgen05:  C = act(LN?(A) . W^T + bias + addmat + residual)  with the same fused
// epilogue set as the fp32 SIMT GEMM, "bf16x2 split" numerics (tc_common.cuh), optional LayerNorm fused
// into the operand load (every CTA tile holds complete 128-feature rows, so the statistics are
// row-local).  These GEMMs have only 32 FLOP per HBM byte: the kernel is built to stream --
// persistent CTAs, warp-specialised:
//   warps 0-7  loaders : coalesced fp32 row loads (one warp per row, float4 per lane, 16 rows = 8 KB in flight per
//                        warp), optional LN by warp shuffles, split into bf16 hi/lo, swizzled K-major stores
//                        into a 2-stage ring
//   (warp 0, lane 0)   : MMA issue -- 3 passes x K/16 tcgen05.mma per stage into a double-buffered TMEM
//                        accumulator, issued while the warp's next row loads are in flight; weights (bf16 hi/lo
//                        planes, TMA, SWIZZLE_128B) stay resident
//   warps 8-15 epilogue (two per TMEM lane quarter, splitting the columns): TMEM -> registers -> (K = 128: transposed through shared memory so that every
//                        global access is 4 rows x 128 contiguous bytes) -> bias / residual / activation -> global
//                        (fp32 and/or bf16 hi/lo planes for the attention kernels)
// The kernel body is written branch-free on purpose: an earlier version skipped unused epilogue features with
// branches over large blocks in a ~100 KB kernel and spent most of its time in instruction fetch (measured:
// ~2000 cycles for ~300 executed instructions).  Here unused operands read a zero block with stride 0, unused
// outputs are predicated stores, and only the activation is skipped as one small uniform block.
// Supported shapes: K = 128 (BN = 128 columns per CTA column-block) and K = 384 (BN = 64), N % BN == 0.
#include <string.h>
#include "tc_common.cuh"

namespace dip {
namespace tcgemm {

using namespace tc;

constexpr uint32_t A_PLANE = 32768;                 // [2 kb][128 rows][64] bf16
constexpr uint32_t A_STAGE = 2 * A_PLANE;           // hi + lo
constexpr uint32_t OFF_W = 2 * A_STAGE;             // after the 2-stage A ring (128 KB)
constexpr uint32_t W_BYTES = 98304;                 // up to 96 KB of weight planes (K = 128: 64 KB + output staging)
constexpr uint32_t OFF_BAR = OFF_W + W_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256;
constexpr uint32_t STG_WARP = 32 * 32 * 4;          // per epilogue warp: a 32 x 32 fp32 block, 16-byte chunks XOR-swizzled by row (conflict-free both ways)
constexpr uint32_t BIG = 1u << 30;

enum { G_WFULL = 0, G_AFULL = 1, G_AEMPTY = 3, G_CFULL = 5, G_CEMPTY = 7 };

#ifdef DIP_TRACE
// Development aid (-DDIP_TRACE): CTA (0,0) logs (event << 48 | tile << 32 | clock) from loader warp 0, the MMA lane and
// epilogue warp 0 into three regions of a global buffer.
__device__ unsigned long long* g_trace = nullptr;
struct Tracer {
    unsigned long long* p; int n;
    __device__ Tracer(int region) : p(nullptr), n(0) { if (g_trace && blockIdx.x == 0 && blockIdx.y == 0 && region >= 0) p = g_trace + region * 16384; }
    __device__ __forceinline__ void ev(int e, int tile) { if (p && n < 16384) p[n++] = ((unsigned long long)e << 48) | ((unsigned long long)tile << 32) | (unsigned long long)(unsigned int)clock64(); }
};
#define TR(e, t) tracer.ev((e), (t))
__device__ int g_dbg = 0;      // bit 0: epilogue skips its global stores; bit 1: loaders skip their global loads
#define DBG(bit) (g_dbg & (bit))
#else
#define DBG(bit) 0
struct Tracer { __device__ Tracer(int) {} };
#define TR(e, t)
#endif

__device__ float g_zeros[4 * 128];                  // what absent operands read (stride 0)

// n / d for n < 2^31 in three instructions (multiplier and shift precomputed on the host): the row maps below need a few
// divisions per row, and plain integer division made this memory-bound kernel instruction-bound (measured).
struct FastDiv {
    uint32_t d, mul, shr;
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1u ? n : (__umulhi(n, mul) >> shr); }
};
static FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d; f.mul = 0; f.shr = 0;
    if (d > 1) {
        int lg = 0;
        while ((1ull << lg) < d) ++lg;                       // ceil(log2 d)
        const int p = 31 + lg;
        f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
        f.shr = (uint32_t)(p - 32);
    }
    return f;
}

// Every optional feature is expressed as data: absent inputs point at g_zeros with stride 0, absent outputs are null.
struct Params {
    int M, N, n_tiles, act;
    // A: K = 128 -> row r = (b, t) with b = r / a_T, t = r % a_T: t < a_split ? a_shared + t*128 : a_batch + (b*(a_T-a_split) + t-a_split)*128
    //    K = 384 -> a_batch + r*lda + kc*128
    //    row windows: GEMM row r reads input row (r / in_rows) * in_T + in_off + r % in_rows (identity: in_rows = BIG, in_T = in_off = 0)
    const float *a_shared, *a_batch;
    FastDiv a_T; uint32_t a_split;
    int64_t a_sstride, res_sstride, add_bstride;    // per-sample strides of the shared A / residual blocks and of addmat (0: one block for the wave)
    FastDiv in_rows; uint32_t in_T, in_off;
    int lda;
    const float *ln_gamma, *ln_beta;
    float *ln_mean, *ln_rstd;
    // epilogue
    FastDiv rows_in; uint32_t rows_out, row_off;    // output row = (r / rows_in) * rows_out + row_off + r % rows_in
    const float* bias;                              // N entries (or zeros)
    const float* row_bias_scale;                    // M entries or null (-> 1)
    const float* addmat; uint32_t add_ld;           // row t_in, stride add_ld (0 when absent)
    const float *res_shared, *res_batch; FastDiv res_T; uint32_t res_split, res_mul;   // residual rows like A; res_mul = 0 when absent
    const float* aux; uint32_t aux_ld;              // operand of the QuickGELU-gradient epilogue (stride 0 when absent)
    float *C, *preact; void *split_hi, *split_lo;
    int ldc;
};

__device__ __forceinline__ const float* src_row(const float* shared, const float* batch, const FastDiv& T, uint32_t split, int64_t sstride, uint32_t r) {
    const uint32_t b = T.div(r), t = r - b * T.d;
    const bool sh = t < split;
    const float* base = sh ? shared + (int64_t)b * sstride : batch;
    const uint32_t idx = sh ? t : b * (T.d - split) + (t - split);
    return base + (size_t)idx * D;
}

// sigmoid by MUFU.EX2 + MUFU.RCP (~2 ulp): the slow paths of expf / IEEE division / __frcp_rn would triple the epilogue's size
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_sigmoid(float z) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * z)); }
// ACT (template): 0 none, 1 QuickGELU, 2 multiply by the QuickGELU derivative at aux, 3 any (selected at run time: ReLU / sigmoid / ...)
template <int ACT>
__device__ __forceinline__ float act_apply(float v, int act, float aux) {
    if (ACT == 1) return v * fast_sigmoid(1.702f * v);
    if (ACT == 2) {
        const float sg = fast_sigmoid(1.702f * aux);
        return v * (sg + 1.702f * aux * sg * (1.0f - sg));
    }
    const float z = act == DIP_ACT_MUL_QUICKGELU_GRAD ? aux : v;
    const float sg = fast_sigmoid((act == DIP_ACT_SIGMOID ? 1.0f : 1.702f) * z);
    const float gelu = v * sg, grad = v * (sg + 1.702f * z * sg * (1.0f - sg));
    return act == DIP_ACT_RELU ? fmaxf(v, 0.f) : act == DIP_ACT_QUICKGELU ? gelu : act == DIP_ACT_SIGMOID ? sg : grad;
}

// Epilogue of one float4 (columns c..c+3) of one output row, in two phases so that a warp can request the global operands of
// several rows before it stores anything (stores may alias the residual: in-place residual adds).  ROWOPS = false: the GEMM has
// no per-row operand (no residual / aux / addmat / row-scaled bias) and phase one only maps the row.
struct EpiIn {
    float4 ext;        // bias * row scale + addmat + residual
    float4 aux;
    uint32_t out_row;
};
template <bool ROWOPS>
__device__ __forceinline__ EpiIn epi_load(const Params& P, uint32_t row, int c, const float4& bv) {
    EpiIn e;
    const uint32_t b = P.rows_in.div(row), t_in = row - b * P.rows_in.d;
    e.out_row = b * P.rows_out + P.row_off + t_in;
    if (ROWOPS) {
        float rbs = 1.0f;
        if (P.row_bias_scale) rbs = P.row_bias_scale[row];
        const float4 am = *reinterpret_cast<const float4*>(P.addmat + (int64_t)b * P.add_bstride + (size_t)t_in * P.add_ld + c);
        const float4 rv = *reinterpret_cast<const float4*>(src_row(P.res_shared, P.res_batch, P.res_T, P.res_split, P.res_sstride, e.out_row * P.res_mul) + c);
        e.aux = *reinterpret_cast<const float4*>(P.aux + (size_t)e.out_row * P.aux_ld + c);
        e.ext = make_float4(bv.x * rbs + am.x + rv.x, bv.y * rbs + am.y + rv.y, bv.z * rbs + am.z + rv.z, bv.w * rbs + am.w + rv.w);
    } else {
        e.ext = bv;
        e.aux = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return e;
}
// predicated global stores without a branch: the slots of a batch stay one basic block, so their shared-memory loads, adds
// and stores overlap instead of running one slot after the other
__device__ __forceinline__ void st_global_v4(float* p, float a, float b, float c, float d, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.global.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(p), "f"(a), "f"(b), "f"(c),
                 "f"(d), "r"((uint32_t)ok)
                 : "memory");
}
__device__ __forceinline__ void st_global_v2(void* p, uint32_t a, uint32_t b, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.global.v2.b32 [%0], {%1, %2};\n\t}" ::"l"(p), "r"(a), "r"(b), "r"((uint32_t)ok) : "memory");
}

template <int ACT>
__device__ __forceinline__ void epi_finish(const Params& P, const EpiIn& e, int c, float4 a4, bool ok) {
    float x[4] = {a4.x + e.ext.x, a4.y + e.ext.y, a4.z + e.ext.z, a4.w + e.ext.w};
    const size_t o = (size_t)e.out_row * P.ldc + c;
    if (P.preact) st_global_v4(P.preact + o, x[0], x[1], x[2], x[3], ok);
    if (ACT != 0) {     // compiled out of the variants without activation: a skipped block this size costs instruction-fetch stalls
        x[0] = act_apply<ACT>(x[0], P.act, e.aux.x); x[1] = act_apply<ACT>(x[1], P.act, e.aux.y);
        x[2] = act_apply<ACT>(x[2], P.act, e.aux.z); x[3] = act_apply<ACT>(x[3], P.act, e.aux.w);
    }
    if (P.C) st_global_v4(P.C + o, x[0], x[1], x[2], x[3], ok);
    if (P.split_hi) {
        uint32_t h0, l0, h1, l1;
        split2(x[0], x[1], h0, l0);
        split2(x[2], x[3], h1, l1);
        st_global_v2(reinterpret_cast<__nv_bfloat16*>(P.split_hi) + o, h0, h1, ok);
        st_global_v2(reinterpret_cast<__nv_bfloat16*>(P.split_lo) + o, l0, l1, ok);
    }
}

constexpr int N_LOADERS = 8, W_MMA = 0, W_EPI = N_LOADERS, N_EPI = 8, N_THREADS = (N_LOADERS + N_EPI) * 32;      // 16 warps: 128 registers each
constexpr int RPW = 128 / N_LOADERS;      // rows of a tile per loader warp, loaded in batches of 16

// APL = true: A arrives as ready-made bf16 hi/lo planes (dip_layernorm_split) and is fetched by TMA straight into the MMA layout --
// no loader warps.  Used by the q|k|v in-projection, whose three column CTAs per row tile would otherwise each repeat the row
// loads, the LayerNorm and the split.
template <int KCH, int BN, bool LN, bool ROWOPS, int ACT, bool APL>
__global__ void __launch_bounds__(N_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo, const __grid_constant__ CUtensorMap tm_a_hi,
               const __grid_constant__ CUtensorMap tm_a_lo, const Params P) {
    extern __shared__ __align__(1024) uint8_t sm[];
    if ((smem_u32(sm) & 1023u) != 0u) __trap();
    uint8_t* sW = sm + OFF_W;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    constexpr uint32_t W_BLOCK = BN * 128;                 // one 64-column k-block of BN weight rows
    constexpr uint32_t W_PLANE = 2 * KCH * W_BLOCK;
    constexpr uint32_t IDESC = instr_desc(128, BN, 0, 0);
    constexpr bool STAGED = (2 * W_PLANE + N_EPI * STG_WARP <= W_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * BN;

    if (threadIdx.x == 0) {
        mbar_init(&bars[G_WFULL], 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars[G_AFULL + s], APL ? 1 : N_LOADERS * 32);
            mbar_init(&bars[G_AEMPTY + s], 1);
            mbar_init(&bars[G_CFULL + s], 1);
            mbar_init(&bars[G_CEMPTY + s], N_EPI * 32);
        }
        mbar_fence_init();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    Tracer tracer(lane != 0 ? -1 : warp == 0 ? 0 : warp == 1 ? 1 : warp == W_EPI ? 2 : -1);
    if (warp < N_LOADERS) {
        // ------------------------------------------------------------------ loaders: warp w owns rows 16w .. 16w+15 of every tile
        float4 gam = make_float4(1.f, 1.f, 1.f, 1.f), bet = make_float4(0.f, 0.f, 0.f, 0.f);
        if (LN) {
            gam = reinterpret_cast<const float4*>(P.ln_gamma)[lane];
            bet = reinterpret_cast<const float4*>(P.ln_beta)[lane];
        }
        // lane l holds columns 4l..4l+3 of its row: k-block l/16, 16-byte chunk (l%16)/2, half l%2
        const uint32_t lane_off = (uint32_t)(lane >> 4) * 16384u + (uint32_t)(lane & 1) * 8u;
        const int chunk = (lane & 15) >> 1;
        const bool stats = LN && P.ln_mean != nullptr && blockIdx.y == 0 && lane == 0;
        // Loader warp 0 doubles as the MMA issuer: lane 0 issues the MMAs of stage it-1 right after the warp has requested the
        // rows of stage it, i.e. in the shadow of its own global-load latency (a 17th warp would cut the register budget to 96).
        const bool issuer = warp == W_MMA && lane == 0;
        const uint32_t aA = smem_u32(sm), aW = smem_u32(sW);
        if (issuer) {
            mbar_expect_tx(&bars[G_WFULL], 2 * W_PLANE);
            for (int kb = 0; kb < 2 * KCH; ++kb) {
                tma_load_2d(sW + kb * W_BLOCK, &tm_w_hi, &bars[G_WFULL], kb * 64, n0);
                tma_load_2d(sW + W_PLANE + kb * W_BLOCK, &tm_w_lo, &bars[G_WFULL], kb * 64, n0);
            }
        }
        const uint64_t dA = smem_desc(aA, 16, 1024), dWd = smem_desc(aW, 16, 1024);
        auto mma_issue = [&](int itx) {                     // all lanes of warp 0 (uniform): they wait, the elected lane issues
            const int ti = itx / KCH, kc = itx - ti * KCH, cb = ti & 1, s = itx & 1;
            if (itx == 0) mbar_wait(&bars[G_WFULL], 0);
            if (kc == 0 && ti >= 2) mbar_wait(&bars[G_CEMPTY + cb], ((ti >> 1) - 1) & 1);
            TR(10, ti);
            mbar_wait(&bars[G_AFULL + s], (itx >> 1) & 1);
            TR(11, ti);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t da = desc_adv(dA, s * A_STAGE), dw = desc_adv(dWd, kc * 2 * W_BLOCK);
                uint32_t acc = kc > 0 ? 1u : 0u;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass)                                 // A hi, hi, lo x W hi, lo, hi
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            tc_mma(tmem + cb * BN, desc_adv(da, (pass == 2 ? A_PLANE : 0) + kb * 16384 + ks * 32),
                                   desc_adv(dw, (pass == 1 ? W_PLANE : 0) + kb * W_BLOCK + ks * 32), IDESC, acc);
                            acc = 1;
                        }
                tc_commit(&bars[G_AEMPTY + s]);
                if (kc == KCH - 1) tc_commit(&bars[G_CFULL + cb]);
            }
            __syncwarp();
            TR(12, ti);
        };
        int it = 0;
        if (APL) {
            if (warp == W_MMA) {
                const int n_it = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
                for (int itx = 0; itx <= n_it; ++itx) {
                    if (itx < n_it) {
                        const int s = itx & 1, tile = blockIdx.x + itx * gridDim.x;
                        if (itx >= 2) mbar_wait(&bars[G_AEMPTY + s], ((itx >> 1) - 1) & 1);
                        if (issuer) {
                            mbar_expect_tx(&bars[G_AFULL + s], A_STAGE);
                            for (int kb = 0; kb < 2; ++kb) {        // rows past M are zero-filled by the tensor map
                                tma_load_2d(sm + s * A_STAGE + kb * 16384, &tm_a_hi, &bars[G_AFULL + s], kb * 64, tile * 128);
                                tma_load_2d(sm + s * A_STAGE + A_PLANE + kb * 16384, &tm_a_lo, &bars[G_AFULL + s], kb * 64, tile * 128);
                            }
                        }
                        __syncwarp();
                    }
                    if (itx >= 1) mma_issue(itx - 1);
                }
            }
        } else {
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
#pragma unroll 1
            for (int kc = 0; kc < KCH; ++kc, ++it) {
                const int s = it & 1;
                TR(1, tile);
                uint8_t* dst = sm + s * A_STAGE + lane_off;
#pragma unroll 1
                for (int bt = 0; bt < RPW; bt += 16) {
                    float4 v[16];
                    // 16 independent 512-byte row loads in flight per warp before any is consumed; rows past M re-read row M-1
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const uint32_t gr = min((uint32_t)tile * 128u + warp * RPW + bt + u, (uint32_t)P.M - 1u);
                        const uint32_t gb = P.in_rows.div(gr);
                        const uint32_t row = gb * P.in_T + P.in_off + (gr - gb * P.in_rows.d);
                        const float* src = KCH == 1 ? src_row(P.a_shared, P.a_batch, P.a_T, P.a_split, P.a_sstride, row) : P.a_batch + (size_t)row * P.lda + kc * 128;
                        v[u] = DBG(2) ? make_float4(1.f, 2.f, 3.f, 4.f) : reinterpret_cast<const float4*>(src)[lane];
                    }
                    if (bt == 0) {
                        if (warp == W_MMA && it > 0) mma_issue(it - 1);
                        if (it >= 2) mbar_wait(&bars[G_AEMPTY + s], ((it >> 1) - 1) & 1);      // the stage's previous MMAs have read it
                        TR(2, tile);
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int r = warp * RPW + bt + u;
                        const uint32_t row = (uint32_t)tile * 128u + r;
                        float4 x = v[u];
                        if (LN) {
                            const float mu = warp_sum(x.x + x.y + x.z + x.w) * (1.0f / D);
                            const float dx = x.x - mu, dy = x.y - mu, dz = x.z - mu, dw = x.w - mu;
                            const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / D);
                            float rs = rsqrtf(var + 1e-5f);
                            rs = rs * (1.5f - 0.5f * (var + 1e-5f) * rs * rs);
                            x = make_float4(dx * rs * gam.x + bet.x, dy * rs * gam.y + bet.y, dz * rs * gam.z + bet.z, dw * rs * gam.w + bet.w);
                            if (stats && row < (uint32_t)P.M) {
                                const uint32_t gb = P.in_rows.div(row);
                                const uint32_t vrow = gb * P.in_T + P.in_off + (row - gb * P.in_rows.d);     // statistics are indexed like the input rows
                                P.ln_mean[vrow] = mu;
                                P.ln_rstd[vrow] = rs;
                            }
                        }
                        uint32_t h0, l0, h1, l1;
                        split2(x.x, x.y, h0, l0);
                        split2(x.z, x.w, h1, l1);
                        const uint32_t off = sw128_off(r, chunk);
                        *reinterpret_cast<uint2*>(dst + off) = make_uint2(h0, h1);
                        *reinterpret_cast<uint2*>(dst + A_PLANE + off) = make_uint2(l0, l1);
                    }
                }
                fence_proxy_async();
                mbar_arrive(&bars[G_AFULL + s]);
                TR(3, tile);
            }
        }
        if (warp == W_MMA && it > 0) mma_issue(it - 1);
        }
    } else {
        // ------------------------------------------------------------------ epilogue: TMEM lane quarter = warp % 4; the two warps of a
        // quarter split the tile's 32-column chunks between them
        const int q = warp & 3, half = (warp - W_EPI) >> 2;
        constexpr int CPW = BN / 32 / 2;                    // chunks per warp
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        float* stg = reinterpret_cast<float*>(sW + 2 * W_PLANE + (warp - W_EPI) * STG_WARP);
        const int cc = (lane & 7) * 4;
        float4 bias4[CPW];                                  // this lane's bias columns, fixed for the whole launch
#pragma unroll
        for (int k = 0; k < CPW; ++k) bias4[k] = *reinterpret_cast<const float4*>(P.bias + n0 + (half * CPW + k) * 32 + (STAGED ? cc : 0));
        int ti = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++ti) {
            const int cb = ti & 1;
            TR(20, tile);
            mbar_wait(&bars[G_CFULL + cb], (ti >> 1) & 1);
            TR(21, tile);
            tc_fence_after();
            const uint32_t row0 = (uint32_t)tile * 128u + q * 32;
#pragma unroll
            for (int k = 0; k < CPW; ++k) {
                const int ch = half * CPW + k;
                float v[32];
                TR(22, tile);
                tmem_ld32(lane_addr + cb * BN + ch * 32, v);
                TR(23, tile);
                if (k == CPW - 1) {
                    tc_fence_before();
                    mbar_arrive(&bars[G_CEMPTY + cb]);           // accumulator drained: the MMA warp may reuse it
                }
                const int c0 = n0 + ch * 32;
                if (STAGED) {
                    // transpose through shared memory: afterwards every warp instruction touches 4 rows x 128 contiguous bytes
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *reinterpret_cast<float4*>(stg + lane * 32 + ((i ^ (lane & 7)) << 2)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    __syncwarp();
                    TR(24, tile);
                    const uint32_t last = (uint32_t)P.M - 1u;
                    constexpr int NB = ROWOPS ? 4 : 8;           // rows whose global operands are requested before any store
#pragma unroll
                    for (int h = 0; h < 8 / NB; ++h) {
                        EpiIn in[NB];
#pragma unroll
                        for (int u = 0; u < NB; ++u) in[u] = epi_load<ROWOPS>(P, min(row0 + (h * NB + u) * 4 + (lane >> 3), last), c0 + cc, bias4[k]);
#pragma unroll
                        for (int u = 0; u < NB; ++u) {
                            const int rr = (h * NB + u) * 4 + (lane >> 3);
                            const float4 a4 = *reinterpret_cast<const float4*>(stg + rr * 32 + (((lane & 7) ^ (rr & 7)) << 2));
                            epi_finish<ACT>(P, in[u], c0 + cc, a4, row0 + rr <= last);
                        }
                        TR(25, tile);
                    }
                    __syncwarp();
                } else {
                    const bool ok = row0 + lane < (uint32_t)P.M;
                    const uint32_t row = min(row0 + lane, (uint32_t)P.M - 1u);
#pragma unroll 2
                    for (int i = 0; i < 32; i += 4) {
                        const float4 bv = *reinterpret_cast<const float4*>(P.bias + c0 + i);
                        const EpiIn in = epi_load<ROWOPS>(P, row, c0 + i, bv);
                        epi_finish<ACT>(P, in, c0 + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]), ok);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tmem, 2 * BN);
}

template <int KCH, int BN, bool LN, bool ROWOPS, int ACT, bool APL = false>
static int launch(const CUtensorMap& whi, const CUtensorMap& wlo, const CUtensorMap& ahi, const CUtensorMap& alo, const Params& P, int n_blocks_y, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        DIP_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<KCH, BN, LN, ROWOPS, ACT, APL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr = true;
    }
    int gx = sm_count() / n_blocks_y;
    if (gx < 1) gx = 1;
    if (gx > P.n_tiles) gx = P.n_tiles;
    dim3 grid(gx, n_blocks_y);
    gemm_tc_kernel<KCH, BN, LN, ROWOPS, ACT, APL><<<grid, N_THREADS, SMEM_BYTES, st>>>(whi, wlo, ahi, alo, P);
    return 0;
}

}  // namespace tcgemm
}  // namespace dip

using namespace dip;

extern "C" {

#ifdef DIP_TRACE
int dip_debug_set_gemm_knobs(int knobs) {
    DIP_CUDA(cudaMemcpyToSymbol(tcgemm::g_dbg, &knobs, sizeof(knobs)));
    return 0;
}
int dip_debug_set_trace_gemm(unsigned long long* device_buffer) {
    DIP_CUDA(cudaMemcpyToSymbol(tcgemm::g_trace, &device_buffer, sizeof(device_buffer)));
    return 0;
}
#endif

int dip_gemm_tc(const dip_gemm_args* args, const void* w_hi, const void* w_lo, const dip_rows* a_rows, const float* ln_gamma,
                const float* ln_beta, float* ln_mean, float* ln_rstd, dip_stream_t stream) {
    DIP_REQUIRE(args && w_hi && w_lo, "dip_gemm_tc: null argument");
    const dip_gemm_args& g = *args;
    const bool rows_a = a_rows && a_rows->batch;
    DIP_REQUIRE((g.A || rows_a) && (g.C || g.split_hi || g.preact), "dip_gemm_tc: no input / no output");
    DIP_REQUIRE(g.M > 0 && (g.K == 128 || g.K == 384), "dip_gemm_tc: K must be 128 or 384 (K=%d)", g.K);
    const int BN = g.K == 128 ? 128 : 64;
    DIP_REQUIRE(g.N % BN == 0, "dip_gemm_tc: N must be a multiple of %d (N=%d)", BN, g.N);
    DIP_REQUIRE(g.ldc % 4 == 0 && (g.A == nullptr || g.lda % 4 == 0), "dip_gemm_tc: lda/ldc must be multiples of 4");
    DIP_REQUIRE(g.res.batch == nullptr || g.N == D, "dip_gemm_tc: residual needs N == 128");
    DIP_REQUIRE(g.act != DIP_ACT_MUL_QUICKGELU_GRAD || g.aux, "dip_gemm_tc: aux missing");
    DIP_REQUIRE((g.split_hi == nullptr) == (g.split_lo == nullptr), "dip_gemm_tc: split_hi and split_lo go together");
    DIP_REQUIRE(!rows_a || g.K == 128, "dip_gemm_tc: row sources need K == 128");
    DIP_REQUIRE(g.K == 384 || rows_a || g.lda == 128, "dip_gemm_tc: K == 128 needs contiguous rows (lda == 128) or a row source");
    DIP_REQUIRE(ln_gamma == nullptr || (g.K == 128 && ln_beta), "dip_gemm_tc: LayerNorm prologue needs K == 128 and beta");
    DIP_REQUIRE((ln_mean == nullptr) == (ln_rstd == nullptr), "dip_gemm_tc: mean and rstd go together");
    DIP_REQUIRE(g.rows_in == 0 || (g.rows_out >= g.rows_in + g.row_off), "dip_gemm_tc: bad row remap");
    DIP_REQUIRE(g.bias || g.N <= 4 * 128, "dip_gemm_tc: N too large for the zero block");
    DIP_REQUIRE(g.in_T == 0 || (g.rows_in > 0 && g.in_off >= 0 && g.in_T >= g.rows_in + g.in_off), "dip_gemm_tc: bad input row window");
    float* zeros = nullptr;
    DIP_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&zeros), tcgemm::g_zeros));
    tcgemm::Params P;
    memset(&P, 0, sizeof(P));
    P.M = g.M; P.N = g.N; P.n_tiles = cdiv(g.M, 128); P.act = g.act;
    if (rows_a) {
        P.a_shared = a_rows->shared ? a_rows->shared : a_rows->batch; P.a_batch = a_rows->batch;
        P.a_split = (uint32_t)a_rows->split;
        P.a_sstride = a_rows->shared ? a_rows->shared_stride : 0;
        P.a_T = tcgemm::make_fastdiv(a_rows->split > 0 ? (uint32_t)a_rows->T : tcgemm::BIG);
    } else {
        P.a_shared = g.A; P.a_batch = g.A; P.a_T = tcgemm::make_fastdiv(tcgemm::BIG); P.a_split = 0;
    }
    P.lda = g.lda;
    if (g.in_T > 0) { P.in_rows = tcgemm::make_fastdiv((uint32_t)g.rows_in); P.in_T = (uint32_t)g.in_T; P.in_off = (uint32_t)g.in_off; }
    else { P.in_rows = tcgemm::make_fastdiv(tcgemm::BIG); P.in_T = 0; P.in_off = 0; }
    P.ln_gamma = ln_gamma; P.ln_beta = ln_beta; P.ln_mean = ln_mean; P.ln_rstd = ln_rstd;
    if (g.rows_in > 0) { P.rows_in = tcgemm::make_fastdiv((uint32_t)g.rows_in); P.rows_out = (uint32_t)g.rows_out; P.row_off = (uint32_t)g.row_off; }
    else { P.rows_in = tcgemm::make_fastdiv(tcgemm::BIG); P.rows_out = 0; P.row_off = 0; }
    P.bias = g.bias ? g.bias : zeros;
    P.row_bias_scale = g.row_bias_scale;
    P.addmat = g.addmat ? g.addmat : zeros; P.add_ld = g.addmat ? (uint32_t)g.N : 0u; P.add_bstride = g.addmat ? g.addmat_stride : 0;
    if (g.res.batch) {
        P.res_shared = g.res.shared ? g.res.shared : g.res.batch; P.res_batch = g.res.batch;
        P.res_split = (uint32_t)g.res.split; P.res_T = tcgemm::make_fastdiv(g.res.split > 0 ? (uint32_t)g.res.T : tcgemm::BIG); P.res_mul = 1;
        P.res_sstride = g.res.shared ? g.res.shared_stride : 0;
    } else {
        P.res_shared = zeros; P.res_batch = zeros; P.res_split = 0; P.res_T = tcgemm::make_fastdiv(tcgemm::BIG); P.res_mul = 0;
    }
    P.aux = g.aux ? g.aux : zeros; P.aux_ld = g.aux ? (uint32_t)g.ldc : 0u;
    P.C = g.C; P.preact = g.preact; P.split_hi = g.split_hi; P.split_lo = g.split_lo; P.ldc = g.ldc;
    CUtensorMap whi, wlo;
    DIP_TRY(make_tmap_bf16_2d(&whi, w_hi, g.N, g.K, g.K, BN, 64));
    DIP_TRY(make_tmap_bf16_2d(&wlo, w_lo, g.N, g.K, g.K, BN, 64));
    DIP_PROF(DIP_PROF_GEMM, stream);
    cudaStream_t st = as_stream(stream);
    const bool rowops = g.row_bias_scale || g.addmat || g.res.batch || g.aux;
    const int actk = g.act == DIP_ACT_NONE ? 0 : g.act == DIP_ACT_QUICKGELU ? 1 : g.act == DIP_ACT_MUL_QUICKGELU_GRAD ? 2 : 3;
    const int shape = g.K == 128 ? (ln_gamma ? 1 : 0) : 2;
    const int variant = (shape * 2 + (rowops ? 1 : 0)) * 4 + actk;
    const int ny = g.N / BN;
    switch (variant) {
#define DIP_GEMM_CASE(v, KCH, BNv, LNv, RO, AC) case v: DIP_TRY((tcgemm::launch<KCH, BNv, LNv, RO, AC>(whi, wlo, whi, wlo, P, ny, st))); break;
#define DIP_GEMM_ACTS(v0, KCH, BNv, LNv, RO) DIP_GEMM_CASE(v0, KCH, BNv, LNv, RO, 0) DIP_GEMM_CASE(v0 + 1, KCH, BNv, LNv, RO, 1) \
        DIP_GEMM_CASE(v0 + 2, KCH, BNv, LNv, RO, 2) DIP_GEMM_CASE(v0 + 3, KCH, BNv, LNv, RO, 3)
        DIP_GEMM_ACTS(0, 1, 128, false, false) DIP_GEMM_ACTS(4, 1, 128, false, true)
        DIP_GEMM_ACTS(8, 1, 128, true, false)  DIP_GEMM_ACTS(12, 1, 128, true, true)
        DIP_GEMM_ACTS(16, 3, 64, false, false) DIP_GEMM_ACTS(20, 3, 64, false, true)
#undef DIP_GEMM_ACTS
#undef DIP_GEMM_CASE
        default: DIP_FAIL("dip_gemm_tc: no kernel variant %d", variant);
    }
    DIP_LAUNCHED();
    return 0;
}

/* C = A . W^T + bias with A given as bf16 hi/lo planes (M,128) -- see include/dip_b200.h. */
int dip_gemm_tc_planes(const dip_gemm_args* args, const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, dip_stream_t stream) {
    DIP_REQUIRE(args && a_hi && a_lo && w_hi && w_lo, "dip_gemm_tc_planes: null argument");
    const dip_gemm_args& g = *args;
    DIP_REQUIRE(g.M > 0 && g.K == 128 && g.N % 128 == 0, "dip_gemm_tc_planes: needs K == 128 and N %% 128 == 0 (K=%d N=%d)", g.K, g.N);
    DIP_REQUIRE(g.C || g.split_hi || g.preact, "dip_gemm_tc_planes: no output");
    DIP_REQUIRE(g.ldc % 4 == 0 && (g.split_hi == nullptr) == (g.split_lo == nullptr), "dip_gemm_tc_planes: bad ldc / split planes");
    DIP_REQUIRE(!g.row_bias_scale && !g.addmat && !g.res.batch && !g.aux && g.act == DIP_ACT_NONE && g.in_T == 0,
                "dip_gemm_tc_planes: bias-only epilogue (no row operands, activation, input windows)");
    DIP_REQUIRE(g.rows_in == 0 || (g.rows_out >= g.rows_in + g.row_off), "dip_gemm_tc_planes: bad row remap");
    float* zeros = nullptr;
    DIP_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&zeros), tcgemm::g_zeros));
    DIP_REQUIRE(g.bias || g.N <= 4 * 128, "dip_gemm_tc_planes: N too large for the zero block");
    tcgemm::Params P;
    memset(&P, 0, sizeof(P));
    P.M = g.M; P.N = g.N; P.n_tiles = cdiv(g.M, 128); P.act = DIP_ACT_NONE;
    P.a_T = tcgemm::make_fastdiv(tcgemm::BIG); P.in_rows = tcgemm::make_fastdiv(tcgemm::BIG);
    P.rows_in = tcgemm::make_fastdiv(tcgemm::BIG); P.res_T = tcgemm::make_fastdiv(tcgemm::BIG);
    if (g.rows_in > 0) { P.rows_in = tcgemm::make_fastdiv((uint32_t)g.rows_in); P.rows_out = (uint32_t)g.rows_out; P.row_off = (uint32_t)g.row_off; }
    P.bias = g.bias ? g.bias : zeros;
    P.addmat = zeros; P.res_shared = zeros; P.res_batch = zeros; P.aux = zeros;
    P.C = g.C; P.preact = g.preact; P.split_hi = g.split_hi; P.split_lo = g.split_lo; P.ldc = g.ldc;
    CUtensorMap whi, wlo, ahi, alo;
    DIP_TRY(make_tmap_bf16_2d(&whi, w_hi, g.N, 128, 128, 128, 64));
    DIP_TRY(make_tmap_bf16_2d(&wlo, w_lo, g.N, 128, 128, 128, 64));
    DIP_TRY(make_tmap_bf16_2d(&ahi, a_hi, g.M, 128, 128, 128, 64));
    DIP_TRY(make_tmap_bf16_2d(&alo, a_lo, g.M, 128, 128, 128, 64));
    DIP_PROF(DIP_PROF_GEMM, stream);
    DIP_TRY((tcgemm::launch<1, 128, false, false, 0, true>(whi, wlo, ahi, alo, P, g.N / 128, as_stream(stream))));
    DIP_LAUNCHED();
    return 0;
}

}  // extern "C"
