// Host-side orchestration of one wave (B samples of one instance) of the guided sampler:
// denoiser forward, decoder forward with saved activations, K2 guidance, decoder input-gradient
// backward, fused DDIM/DDPM update -- model/diffusion.py:594-644 / 429-454 as a fixed sequence of
// kernel launches on the caller's stream.  No device synchronisation, no allocation in the step.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"

namespace dip {

constexpr int MAX_LAYERS = 8;

// a weight in the three forms the kernels consume: fp32 (SIMT GEMM) and bf16 hi / lo planes (tcgen05 GEMM)
struct WRef {
    float* w = nullptr;
    void* hi = nullptr;
    void* lo = nullptr;
};

struct BlockW {
    float *ln1_w, *ln1_b, *in_b, *out_b, *ln2_w, *ln2_b, *fc_b, *proj_b;
    WRef in_w, out_w, fc_w, proj_w;
    WRef in_wT, out_wT, fc_wT, proj_wT;   // transposed copies for the input-gradient GEMMs
};

__global__ void transpose_kernel(float* __restrict__ out, const float* __restrict__ in, int rows, int cols) {
    int64_t total = (int64_t)rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
        out[(int64_t)c * rows + r] = in[i];
    }
}

struct Pool {
    std::vector<void*> ptrs;
    int alloc(float** p, size_t n_floats) {
        void* q = nullptr;
        DIP_CUDA(cudaMalloc(&q, n_floats * sizeof(float)));
        ptrs.push_back(q);
        *p = reinterpret_cast<float*>(q);
        return 0;
    }
    void release() {
        for (void* q : ptrs) cudaFree(q);
        ptrs.clear();
    }
};

}  // namespace dip

using namespace dip;

struct dip_engine {
    int n_den = 0, n_dec = 0;
    BlockW den[MAX_LAYERS], dec[MAX_LAYERS];
    float *w1 = nullptr, *w1m = nullptr, *b1 = nullptr, *b2 = nullptr;
    WRef w1x, w2;
    float *proj_w = nullptr, *proj_b = nullptr;
    Pool den_pool, dec_pool, inst_pool;
    bool has_den = false, has_dec = false, has_inst = false, mproj_valid = false;
    dip_instance inst;                      // instance 0 (the only one unless dip_engine_set_instances was given several)
    float* mproj = nullptr;
    uint8_t *mask_den = nullptr, *mask_dec = nullptr;
    int inst_cap_L = 0;
    int mode = DIP_PRECISION_TC_SPLIT;
    // decoder layer 0, tensor-core mode: the q|k|v planes of the L mip tokens are constants of the wave (they depend on the instance and
    // the decoder weights only); they are written by the first decode into a given workspace and kept -- later steps normalise and
    // project the x tokens only.  Valid for (workspace, B) below; reset by anything that changes the instance, the map or the weights.
    const void* dec0_ws = nullptr;
    int dec0_B = 0;
    // ---- waves that mix instances (dip_engine_set_instances + dip_engine_set_wave_map) ----
    std::vector<dip_instance> insts;        // host copies of the resident instances
    dip_instance* table = nullptr;          // device copy (guidance / rounding kernels index it by inst_of)
    int table_cap = 0;
    int n_max = 0, M_max = 0;
    Pool wave_pool;                         // per-sample gathers for the current map
    int wave_B = 0, wave_cap = 0;           // wave_B > 0: a map is set
    int32_t* inst_of = nullptr;             // device (wave_B)
    float *wave_mip = nullptr, *wave_mproj = nullptr;      // (B, L, 128) each
    uint8_t *wave_kpm = nullptr, *wave_mask_den = nullptr, *wave_mask_dec = nullptr;     // (B, L), (B, L+1), (B, 2L)
};

namespace dip {

// allocate the bf16 planes of an (n-element) fp32 weight already on the device
static int make_planes(Pool& pool, WRef& w, size_t n, cudaStream_t st) {
    float *h = nullptr, *l = nullptr;
    DIP_TRY(pool.alloc(&h, (n + 1) / 2));
    DIP_TRY(pool.alloc(&l, (n + 1) / 2));
    w.hi = h; w.lo = l;
    return dip_split_bf16(w.hi, w.lo, w.w, (int64_t)n, st);
}

static int copy_weight(Pool& pool, WRef& w, const float* src, size_t n, cudaStream_t st) {
    DIP_REQUIRE(src, "weights: null pointer");
    DIP_TRY(pool.alloc(&w.w, n));
    DIP_CUDA(cudaMemcpyAsync(w.w, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return make_planes(pool, w, n, st);
}

static int transposed_weight(Pool& pool, WRef& wT, const WRef& w, int rows, int cols, cudaStream_t st) {
    DIP_TRY(pool.alloc(&wT.w, (size_t)rows * cols));
    transpose_kernel<<<64, 256, 0, st>>>(wT.w, w.w, rows, cols);
    DIP_LAUNCHED();
    return make_planes(pool, wT, (size_t)rows * cols, st);
}

static int copy_block(Pool& pool, BlockW& w, const dip_block_weights& s, cudaStream_t st) {
    struct Item { float** dst; const float* src; size_t n; };
    Item items[] = {
        {&w.ln1_w, s.ln1_w, D}, {&w.ln1_b, s.ln1_b, D}, {&w.in_b, s.in_proj_b, 3 * D}, {&w.out_b, s.out_proj_b, D},
        {&w.ln2_w, s.ln2_w, D}, {&w.ln2_b, s.ln2_b, D}, {&w.fc_b, s.fc_b, D}, {&w.proj_b, s.proj_b, D},
    };
    for (auto& it : items) {
        DIP_REQUIRE(it.src, "block weights: null pointer");
        DIP_TRY(pool.alloc(it.dst, it.n));
        DIP_CUDA(cudaMemcpyAsync(*it.dst, it.src, it.n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    DIP_TRY(copy_weight(pool, w.in_w, s.in_proj_w, 3 * D * D, st));
    DIP_TRY(copy_weight(pool, w.out_w, s.out_proj_w, D * D, st));
    DIP_TRY(copy_weight(pool, w.fc_w, s.fc_w, D * D, st));
    DIP_TRY(copy_weight(pool, w.proj_w, s.proj_w, D * D, st));
    DIP_TRY(transposed_weight(pool, w.in_wT, w.in_w, 3 * D, D, st));
    DIP_TRY(transposed_weight(pool, w.out_wT, w.out_w, D, D, st));
    DIP_TRY(transposed_weight(pool, w.fc_wT, w.fc_w, D, D, st));
    DIP_TRY(transposed_weight(pool, w.proj_wT, w.proj_w, D, D, st));
    return 0;
}

// ---- workspace carving -------------------------------------------------------------------------
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* b) : base(reinterpret_cast<char*>(b)) {}
    template <typename T>
    T* take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

struct LayerSave {
    float *mean1, *rstd1, *qkv, *ao, *lse, *xmid, *mean2, *rstd2, *u, *xout;
    void *qkv_hi, *qkv_lo;   // bf16 planes of qkv (tensor-core mode)
};

struct Workspace {
    // denoiser
    float *hd, *tmp, *tmp2, *qkv1, *ao1;
    void *qkv1_hi, *qkv1_lo, *do_hi, *do_lo;
    void *ln_hi, *ln_lo;     // bf16 planes of LayerNorm(x), the A operand of the in-projection (tensor-core mode)
    // decoder
    LayerSave ls[MAX_LAYERS];
    float *h, *h2;
    float *dxa, *dxb, *dxm, *dt1, *dt2, *dao, *dqkv, *delta;
    float *p_full, *dz, *viol, *obj, *mean_buf;
    void* ds;             // dS^T planes of one attention backward (tensor-core mode), see dip_attn_bwd_tc
    size_t ds_bytes;
    float* fstate;        // softmax state of decoder layer 0's instance-token queries after the instance keys (dip_attn_fwd_tc, key-prefix hoisting)
    size_t bytes;
};

static Workspace carve(const dip_engine* e, int B, void* base) {
    Workspace w;
    Carver c(base);
    const size_t L = e->inst.L, T1 = L + 1, T2 = 2 * L;
    const size_t R1 = (size_t)B * T1, R2 = (size_t)B * T2;
    w.hd = c.take<float>(R1 * D);
    w.tmp = c.take<float>(R1 * D);
    w.tmp2 = c.take<float>(R1 * D);
    w.qkv1 = c.take<float>(R1 * 3 * D);
    w.ao1 = c.take<float>(R1 * D);
    w.qkv1_hi = c.take<uint16_t>(R1 * 3 * D);
    w.qkv1_lo = c.take<uint16_t>(R1 * 3 * D);
    w.ln_hi = c.take<uint16_t>((R1 > R2 ? R1 : R2) * D);
    w.ln_lo = c.take<uint16_t>((R1 > R2 ? R1 : R2) * D);
    for (int l = 0; l < e->n_dec; ++l) {
        LayerSave& s = w.ls[l];
        s.mean1 = c.take<float>(R2); s.rstd1 = c.take<float>(R2);
        s.qkv = c.take<float>(R2 * 3 * D);
        s.ao = c.take<float>(R2 * D);
        s.lse = c.take<float>(R2);
        s.xmid = c.take<float>(R2 * D);
        s.mean2 = c.take<float>(R2); s.rstd2 = c.take<float>(R2);
        s.u = c.take<float>(R2 * D);
        s.xout = c.take<float>(R2 * D);
        s.qkv_hi = c.take<uint16_t>(R2 * 3 * D);
        s.qkv_lo = c.take<uint16_t>(R2 * 3 * D);
    }
    w.h = c.take<float>(R2 * D);
    w.h2 = c.take<float>(R2 * D);
    w.dxa = c.take<float>(R2 * D);
    w.dxb = c.take<float>(R2 * D);
    w.dxm = c.take<float>(R2 * D);
    w.dt1 = c.take<float>(R2 * D);
    w.dt2 = c.take<float>(R2 * D);
    w.dao = c.take<float>(R2 * D);
    w.dqkv = c.take<float>(R2 * 3 * D);
    w.delta = c.take<float>(R2);
    w.do_hi = c.take<uint16_t>(R2 * D);
    w.do_lo = c.take<uint16_t>(R2 * D);
    w.p_full = c.take<float>((size_t)B * L);
    w.dz = c.take<float>((size_t)B * L);
    w.viol = c.take<float>(B);
    w.obj = c.take<float>(B);
    w.mean_buf = c.take<float>((size_t)B * L * D);
    // dS^T scratch: the largest of the decoder layers' backward launches (layer 0: keys >= L, all queries unless it is also the last
    // layer; last layer: all keys, queries >= L)
    w.ds_bytes = 0;
    for (int l = 0; l < e->n_dec; ++l) {
        const size_t need = dip_attn_bwd_tc_ds_bytes(B, (int)T2, l == e->n_dec - 1 ? (int)L : 0, l == 0 ? (int)L : 0);
        w.ds_bytes = need > w.ds_bytes ? need : w.ds_bytes;
    }
    w.ds = c.take<uint8_t>(w.ds_bytes);
    w.fstate = c.take<float>(dip_attn_fwd_tc_state_bytes(B, (int)L) / sizeof(float));
    w.bytes = (c.off + 255) & ~(size_t)255;
    return w;
}

static dip_gemm_args gemm(const float* A, int lda, float* C, int ldc, int64_t M, int N, int K, const float* bias, int act) {
    dip_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.A = A; g.lda = lda; g.bias = bias; g.C = C; g.ldc = ldc;
    g.M = (int)M; g.N = N; g.K = K; g.act = act;
    return g;
}

// DIP_NO_CHAIN=1 falls back to one launch per Linear / LayerNorm for the row-wise half of a block (A/B comparisons, bisecting)
static bool use_chain_env() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DIP_NO_CHAIN");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}
// the chain kernels address rows with 32-bit element offsets (B*T*384 < 2^32): larger waves take the one-launch-per-Linear path
static bool use_chain(int B, int T) { return use_chain_env() && (int64_t)B * T * 3 * D < (1ll << 32); }

static dip_rows plain_rows(const float* p, int T) {
    dip_rows r;
    r.shared = nullptr; r.batch = p; r.T = T; r.split = 0; r.shared_stride = 0;
    return r;
}

// One Linear, optionally preceded by LayerNorm of its input rows `a_rows` (then g.A is ignored):
//   tensor-core mode: one tcgen05 GEMM launch, LayerNorm fused into the operand load;
//   fp32 mode       : LayerNorm kernel into `ln_tmp`, then the SIMT GEMM.
static int linear(int mode, dip_gemm_args g, const WRef& w, const dip_rows* a_rows, const float* ln_g, const float* ln_b, float* mean, float* rstd,
                  float* ln_tmp, dip_stream_t st) {
    if (mode == DIP_PRECISION_TC_SPLIT) return dip_gemm_tc(&g, w.hi, w.lo, a_rows, ln_g, ln_b, mean, rstd, st);
    if (a_rows) {
        DIP_REQUIRE(ln_g && ln_tmp, "linear: row-source inputs go through LayerNorm");
        DIP_TRY(dip_layernorm_fwd(ln_tmp, mean, rstd, *a_rows, ln_g, ln_b, g.M, st));
        g.A = ln_tmp; g.lda = D;
    }
    g.W = w.w;
    if (g.C == nullptr) {
        // the SIMT GEMM always writes fp32 C; callers in fp32 mode pass one
        DIP_FAIL("linear: fp32 mode needs an fp32 output");
    }
    return dip_gemm_nt(&g, st);
}

// Restrict a Linear to the tokens t >= w of every sample; it keeps reading and writing the full-length (B*T) buffers in place.
static void window(dip_gemm_args& g, int B, int T, int w) {
    if (w <= 0) return;
    g.M = B * (T - w);
    g.rows_in = T - w; g.rows_out = T; g.row_off = w;
    g.in_T = T; g.in_off = w;
}

// One ResidualAttentionBlock forward (model/modules.py:293-296).  xin -> xmid -> xout; h, h2 are
// transients.  Statistics / lse / u are saved when non-null.
// win > 0 (tensor-core mode only): the caller will use the output tokens t >= win only (the decoder keeps the last L of
// its 2L tokens, model/decoder.py:45), so in the last layer just those tokens act as queries and go through the
// out-projection and the MLP; keys and values still cover the whole sequence.
static int block_fwd(int mode, const BlockW& w, dip_rows xin, float* xmid, float* xout, float* h, float* h2, void* ln_hi, void* ln_lo, float* qkv,
                     void* qkv_hi, void* qkv_lo, float* ao, float* lse, float* mean1, float* rstd1, float* mean2, float* rstd2, float* u, const uint8_t* mask,
                     int64_t mask_stride, int B, int T, int win, int in_begin, float* fstate, int prefix_keys, dip_stream_t st) {
    const int64_t R = (int64_t)B * T;
    const bool tc = mode == DIP_PRECISION_TC_SPLIT;
    DIP_REQUIRE(R < (1ll << 31), "wave too large: B*T = %lld rows", (long long)R);
    if (!tc) { win = 0; in_begin = 0; }
    // q|k|v = LN1(x) . W_in^T + b_in   (tensor-core mode: only the bf16 planes are written)
    dip_gemm_args g = gemm(nullptr, D, tc ? nullptr : qkv, 3 * D, R, 3 * D, D, w.in_b, DIP_ACT_NONE);
    if (tc) {
        // LayerNorm once into bf16 planes, then a TMA-fed GEMM: the three column blocks share the normalised rows
        // in_begin > 0: the q|k|v planes of the tokens t < in_begin are already in place (constants of the wave, see dip_engine::dec0_ws)
        g.split_hi = qkv_hi; g.split_lo = qkv_lo;
        if (in_begin > 0) { g.M = B * (T - in_begin); g.rows_in = T - in_begin; g.rows_out = T; g.row_off = in_begin; }
        DIP_TRY(dip_layernorm_split(ln_hi, ln_lo, mean1, rstd1, xin, w.ln1_w, w.ln1_b, R, in_begin, st));
        DIP_TRY(dip_gemm_tc_planes(&g, ln_hi, ln_lo, w.in_w.hi, w.in_w.lo, st));
    } else {
        DIP_TRY(linear(mode, g, w.in_w, &xin, w.ln1_w, w.ln1_b, mean1, rstd1, h, st));
    }
    if (tc) {
        // prefix_keys > 0 (decoder layer 0, not also the last layer): the first prefix_keys tokens are wave constants.  in_begin > 0 means
        // their planes -- and the softmax state of their queries against them -- are already in place: resume; otherwise compute in full
        // and save the state for the following calls of the wave.
        const bool hoist = prefix_keys >= 128 && win == 0 && fstate != nullptr;
        DIP_TRY(dip_attn_fwd_tc(ao, lse, qkv_hi, qkv_lo, mask, mask_stride, B, T, win, hoist ? fstate : nullptr, hoist ? prefix_keys : 0,
                                hoist && in_begin > 0 ? 2 : 0, st));
        if (hoist && in_begin == 0)
            DIP_TRY(dip_attn_fwd_tc(nullptr, nullptr, qkv_hi, qkv_lo, mask, mask_stride, B, T, 0, fstate, prefix_keys, 1, st));
    } else {
        DIP_TRY(dip_attn_fwd(ao, lse, qkv, qkv + D, qkv + 2 * D, 3 * D, mask, mask_stride, B, T, st));
    }
    if (tc && use_chain(B, T)) {
        // out-projection + residual, LayerNorm 2, c_fc, QuickGELU, c_proj + residual in one launch (chain_tc.cu)
        DIP_TRY(dip_block_chain_fwd_tc(xout, xmid, mean2, rstd2, u, ao, xin, w.out_w.hi, w.out_w.lo, w.out_b, w.ln2_w, w.ln2_b, w.fc_w.hi, w.fc_w.lo, w.fc_b,
                                       w.proj_w.hi, w.proj_w.lo, w.proj_b, B, T, win, st));
        return 0;
    }
    // x_mid = x + attn . W_out^T + b_out
    g = gemm(ao, D, xmid, D, R, D, D, w.out_b, DIP_ACT_NONE);
    g.res = xin;
    window(g, B, T, win);
    DIP_TRY(linear(mode, g, w.out_w, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st));
    // g = QuickGELU(LN2(x_mid) . W_fc^T + b_fc)   (pre-activation saved for the backward pass)
    dip_rows xm = plain_rows(xmid, T);
    g = gemm(nullptr, D, h2, D, R, D, D, w.fc_b, DIP_ACT_QUICKGELU);
    g.preact = u;
    window(g, B, T, win);
    DIP_TRY(linear(mode, g, w.fc_w, &xm, w.ln2_w, w.ln2_b, mean2, rstd2, h, st));
    // x_out = x_mid + g . W_proj^T + b_proj
    g = gemm(h2, D, xout, D, R, D, D, w.proj_b, DIP_ACT_NONE);
    g.res = xm;
    window(g, B, T, win);
    DIP_TRY(linear(mode, g, w.proj_w, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st));
    return 0;
}

// Input-gradient backward of one block: dxout -> dxin.
//   t_min    : dxin is wanted for the tokens t >= t_min only (layer 0: the mip tokens are constants);
//   do_begin : (tensor-core mode) dxout is identically zero for t < do_begin and those rows of dxout are not read
//              (last layer: only the kept tokens carry a gradient).
// Tokens outside these windows are skipped by every kernel that does not need them: the row-wise chain (MLP, LN2,
// out-projection) runs on t >= do_begin, the attention backward streams queries t >= do_begin only and produces
// dK, dV for t >= t_min and dQ for t >= max(t_min, do_begin), the in-projection backward runs on t >= t_min.
static int block_bwd(int mode, const BlockW& w, const float* dxout, dip_rows xin, const LayerSave& s, const uint8_t* mask, int64_t mask_stride, int B, int T, float* dt1,
                     float* dt2, float* dxm, float* dao, void* do_hi, void* do_lo, float* dqkv, float* delta, void* ds, size_t ds_bytes, float* dxin, int t_min,
                     int do_begin, dip_stream_t st) {
    const int64_t R = (int64_t)B * T;
    const bool tc = mode == DIP_PRECISION_TC_SPLIT;
    if (!tc) do_begin = 0;
    cudaStream_t cs = as_stream(st);
    if (tc && use_chain(B, T)) {
        // c_proj^T, gelu', c_fc^T, LayerNorm-2 backward + residual, out-projection^T (+ dO planes) in one launch (chain_tc.cu)
        DIP_TRY(dip_block_chain_bwd_tc(dxm, dao, do_hi, do_lo, dxout, s.u, s.xmid, s.mean2, s.rstd2, w.ln2_w, w.proj_wT.hi, w.proj_wT.lo, w.fc_wT.hi, w.fc_wT.lo,
                                       w.out_wT.hi, w.out_wT.lo, B, T, do_begin, st));
    } else {
    // dU = (dxout . W_proj) * gelu'(u)
        dip_gemm_args g = gemm(dxout, D, dt1, D, R, D, D, nullptr, DIP_ACT_MUL_QUICKGELU_GRAD);
        g.aux = s.u;
        window(g, B, T, do_begin);
        DIP_TRY(linear(mode, g, w.proj_wT, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st));
        // dH2 = dU . W_fc
        g = gemm(dt1, D, dt2, D, R, D, D, nullptr, DIP_ACT_NONE);
        window(g, B, T, do_begin);
        DIP_TRY(linear(mode, g, w.fc_wT, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st));
        // dxmid = dxout + LN2'(dH2)
        DIP_TRY(dip_layernorm_bwd(dxm, dt2, plain_rows(s.xmid, T), s.mean2, s.rstd2, w.ln2_w, dxout, R, do_begin, st));
        // dAO = dxmid . W_out   (+ bf16 planes for the tensor-core attention backward)
        g = gemm(dxm, D, dao, D, R, D, D, nullptr, DIP_ACT_NONE);
        if (tc) { g.split_hi = do_hi; g.split_lo = do_lo; }
        window(g, B, T, do_begin);
        DIP_TRY(linear(mode, g, w.out_wT, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st));
    }
    if (tc) {
        const int q_begin = t_min > do_begin ? t_min : do_begin;
        if (t_min < do_begin) {
            // tokens t_min .. do_begin-1 still receive a gradient through their keys / values: their dQ and dxmid are exactly zero
            DIP_CUDA(cudaMemset2DAsync(dqkv, 3 * D * sizeof(float), 0, D * sizeof(float), (size_t)R, cs));
            DIP_CUDA(cudaMemset2DAsync(dxm, (size_t)T * D * sizeof(float), 0, (size_t)do_begin * D * sizeof(float), (size_t)B, cs));
        }
        DIP_TRY(dip_attn_bwd_tc(dqkv, dqkv + D, dqkv + 2 * D, 3 * D, s.qkv_hi, s.qkv_lo, do_hi, do_lo, s.ao, dao, s.lse, mask, mask_stride, delta, B, T,
                                do_begin, q_begin, t_min, ds, ds_bytes, st));
    } else {
        DIP_TRY(dip_attn_bwd(dqkv, dqkv + D, dqkv + 2 * D, 3 * D, s.qkv, s.qkv + D, s.qkv + 2 * D, 3 * D, s.ao, dao, s.lse, mask, mask_stride, delta, B, T, st));
    }
    if (tc && use_chain(B, T)) {
        // dxin = dxmid + LN1'(dQKV . W_in) in one launch (chain_tc.cu, mode 2)
        DIP_TRY(dip_block_in_bwd_tc(dxin, dqkv, xin, s.mean1, s.rstd1, w.ln1_w, dxm, w.in_wT.hi, w.in_wT.lo, B, T, t_min, st));
        return 0;
    }
    // dH1 = dQKV . W_in
    dip_gemm_args g = gemm(dqkv, 3 * D, dt1, D, R, D, 3 * D, nullptr, DIP_ACT_NONE);
    if (tc) window(g, B, T, t_min);
    DIP_TRY(linear(mode, g, w.in_wT, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st));
    // dxin = dxmid + LN1'(dH1)
    DIP_TRY(dip_layernorm_bwd(dxin, dt1, xin, s.mean1, s.rstd1, w.ln1_w, dxm, R, t_min, st));
    return 0;
}

// ---- how the kernels see the instance(s) of the current wave: one shared block (stride 0) or per-sample gathers ----------------
struct WaveView {
    const float* mip; int64_t mip_stride;           // shared source of the decoder rows / residual
    const float* mproj; int64_t mproj_stride;       // hoisted mip half of mlp_xt.linear_1
    const uint8_t *kpm; int64_t kpm_stride;
    const uint8_t *mask_den; int64_t mask_den_stride;
    const uint8_t *mask_dec; int64_t mask_dec_stride;
    const dip_instance* table; const int32_t* inst_of;
};

static WaveView view(const dip_engine* e) {
    WaveView v;
    const int64_t L = e->inst.L;
    if (e->wave_B > 0) {
        v.mip = e->wave_mip; v.mip_stride = L * D;
        v.mproj = e->wave_mproj; v.mproj_stride = L * D;
        v.kpm = e->wave_kpm; v.kpm_stride = L;
        v.mask_den = e->wave_mask_den; v.mask_den_stride = L + 1;
        v.mask_dec = e->wave_mask_dec; v.mask_dec_stride = 2 * L;
        v.table = e->table; v.inst_of = e->inst_of;
    } else {
        v.mip = e->inst.mip; v.mip_stride = 0;
        v.mproj = e->mproj; v.mproj_stride = 0;
        v.kpm = e->inst.kpm; v.kpm_stride = 0;
        v.mask_den = e->mask_den; v.mask_den_stride = 0;
        v.mask_dec = e->mask_dec; v.mask_dec_stride = 0;
        v.table = nullptr; v.inst_of = nullptr;
    }
    return v;
}

static dip_rows mip_rows(const WaveView& v, const float* x, int T2, int L) {
    dip_rows r;
    r.shared = v.mip; r.batch = x; r.T = T2; r.split = L; r.shared_stride = v.mip_stride;
    return r;
}

static int check_ready(const dip_engine* e, bool need_den, bool need_dec, int B, size_t ws_bytes, void* ws) {
    DIP_REQUIRE(e, "null engine");
    DIP_REQUIRE(e->has_inst, "dip_engine_set_instance has not been called");
    DIP_REQUIRE(e->wave_B == 0 || e->wave_B == B, "the wave map was set for %d samples, this call has %d", e->wave_B, B);
    DIP_REQUIRE(e->insts.size() <= 1 || e->wave_B > 0, "several instances are resident: dip_engine_set_wave_map first");
    DIP_REQUIRE(!need_den || e->has_den, "dip_engine_set_denoiser has not been called");
    DIP_REQUIRE(!need_dec || e->has_dec, "dip_engine_set_decoder has not been called");
    DIP_REQUIRE(B > 0, "empty wave");
    DIP_REQUIRE(ws != nullptr, "null workspace");
    size_t need = carve(e, B, nullptr).bytes;
    DIP_REQUIRE(ws_bytes >= need, "workspace too small: %zu < %zu bytes", ws_bytes, need);
    DIP_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace must be 256-byte aligned");
    return 0;
}

// denoiser forward into ws.hd (B, L+1, 128); the model output is rows 1..L of every sample
static int denoise(dip_engine* e, const Workspace& w, const float* x, int B, float t, dip_stream_t st) {
    const int L = e->inst.L, T1 = L + 1;
    const int64_t R1 = (int64_t)B * T1;
    const WaveView v = view(e);
    if (!e->mproj_valid) {
        // mip . W1[:, :128]^T + b1 -- constant over samples and steps, hoisted out of the loop (per sample of the map when instances mix)
        const int64_t rows = e->wave_B > 0 ? (int64_t)e->wave_B * L : L;
        dip_gemm_args g = gemm(v.mip, D, const_cast<float*>(v.mproj), D, rows, D, D, e->b1, DIP_ACT_NONE);
        g.W = e->w1m;
        DIP_TRY(dip_gemm_nt(&g, st));
        e->mproj_valid = true;
    }
    DIP_TRY(dip_time_token(w.tmp, B, T1, t, e->w1, e->b1, st));
    dip_gemm_args g = gemm(x, D, w.tmp, D, (int64_t)B * L, D, D, nullptr, DIP_ACT_QUICKGELU);
    g.rows_in = L; g.rows_out = T1; g.row_off = 1; g.addmat = v.mproj; g.addmat_stride = v.mproj_stride;
    DIP_TRY(linear(e->mode, g, e->w1x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st));
    // NO activation after linear_2: the reference's OrderedDict names both QuickGELUs "geru", so the
    // second entry replaces the first and mlp_xt is linear_1 -> QuickGELU -> linear_2 (diffusion.py:145-150)
    g = gemm(w.tmp, D, w.hd, D, R1, D, D, e->b2, DIP_ACT_NONE);
    DIP_TRY(linear(e->mode, g, e->w2, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st));
    for (int l = 0; l < e->n_den; ++l)
        DIP_TRY(block_fwd(e->mode, e->den[l], plain_rows(w.hd, T1), w.hd, w.hd, w.tmp, w.tmp2, w.ln_hi, w.ln_lo, w.qkv1, w.qkv1_hi, w.qkv1_lo, w.ao1, nullptr, nullptr,
                          nullptr, nullptr, nullptr, nullptr, v.mask_den, v.mask_den_stride, B, T1, 0, 0, nullptr, 0, st));
    return 0;
}

// decoder forward with saved activations; final hidden state in ws.ls[n_dec-1].xout, p in ws.p_full
static int decode(dip_engine* e, const Workspace& w, const float* x, int B, dip_stream_t st) {
    const int L = e->inst.L, T2 = 2 * L;
    const WaveView v = view(e);
    for (int l = 0; l < e->n_dec; ++l) {
        const dip_rows xin = l == 0 ? mip_rows(v, x, T2, L) : plain_rows(w.ls[l - 1].xout, T2);
        const LayerSave& s = w.ls[l];
        // layer 0: the mip tokens' planes are in place after the first decode into this workspace
        const bool hoisted = l == 0 && e->mode == DIP_PRECISION_TC_SPLIT && e->dec0_ws == (const void*)s.qkv_hi && e->dec0_B == B;
        DIP_TRY(block_fwd(e->mode, e->dec[l], xin, s.xmid, s.xout, w.h, w.h2, w.ln_hi, w.ln_lo, s.qkv, s.qkv_hi, s.qkv_lo, s.ao, s.lse, s.mean1, s.rstd1, s.mean2,
                          s.rstd2, s.u, v.mask_dec, v.mask_dec_stride, B, T2, l == e->n_dec - 1 ? L : 0, hoisted ? L : 0, l == 0 ? w.fstate : nullptr,
                          l == 0 ? L : 0, st));
        if (l == 0 && e->mode == DIP_PRECISION_TC_SPLIT) { e->dec0_ws = s.qkv_hi; e->dec0_B = B; }
    }
    DIP_TRY(dip_decoder_head_fwd(w.p_full, w.ls[e->n_dec - 1].xout, B, T2, L, e->proj_w, e->proj_b, v.kpm, v.kpm_stride, st));
    return 0;
}

// guidance gradient: returns pointer to d loss / d x rows (sample stride 2L*128)
static int guidance_grad(dip_engine* e, const Workspace& w, const float* x, int B, float lam, const float** grad, int64_t* grad_stride,
                         dip_stream_t st) {
    const int L = e->inst.L, T2 = 2 * L;
    DIP_TRY(decode(e, w, x, B, st));
    const dip_instance& in = e->inst;
    const WaveView v = view(e);
    if (v.table)
        DIP_TRY(dip_guidance_multi(w.dz, w.viol, w.obj, w.p_full, v.table, v.inst_of, lam, B, L, e->n_max, e->M_max, st));
    else
        DIP_TRY(dip_guidance(w.dz, w.viol, w.obj, nullptr, nullptr, w.p_full, in.pos_of_col, in.csr_rowptr, in.csr_col, in.csr_val, in.csc_colptr,
                             in.csc_row, in.csc_val, in.b, in.c, lam, B, L, in.n, in.M, in.nnz, st));
    float* dcur = w.dxa;
    float* dnext = w.dxb;
    DIP_TRY(dip_decoder_head_bwd(dcur, w.dz, B, T2, L, e->proj_w, st));
    for (int l = e->n_dec - 1; l >= 0; --l) {
        const dip_rows xin = l == 0 ? mip_rows(v, x, T2, L) : plain_rows(w.ls[l - 1].xout, T2);
        // the gradient w.r.t. the (constant) mip tokens of layer 0 is never used, and only the last L tokens of the last
        // layer's output carry a gradient (head_bwd writes zeros elsewhere; in tensor-core mode those rows are not even read)
        DIP_TRY(block_bwd(e->mode, e->dec[l], dcur, xin, w.ls[l], v.mask_dec, v.mask_dec_stride, B, T2, w.dt1, w.dt2, w.dxm, w.dao, w.do_hi, w.do_lo, w.dqkv, w.delta, w.ds, w.ds_bytes,
                          dnext, l == 0 ? L : 0, l == e->n_dec - 1 ? L : 0, st));
        float* t = dcur; dcur = dnext; dnext = t;
    }
    *grad = dcur + (int64_t)L * D;
    *grad_stride = (int64_t)T2 * D;
    return 0;
}

static int final_decode(dip_engine* e, const Workspace& w, const float* x, int B, float* p_full, uint8_t* xbits, float* penalty,
                        int64_t* viol_int, int64_t* obj_int, dip_stream_t st) {
    const dip_instance& in = e->inst;
    DIP_TRY(decode(e, w, x, B, st));
    if (p_full) DIP_CUDA(cudaMemcpyAsync(p_full, w.p_full, (size_t)B * in.L * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(st)));
    if (xbits || penalty || viol_int || obj_int) {
        const WaveView v = view(e);
        if (v.table)
            DIP_TRY(dip_round_feasibility_multi(xbits, penalty, viol_int, obj_int, w.p_full, v.table, v.inst_of, B, in.L, e->n_max, st));
        else
            DIP_TRY(dip_round_feasibility(xbits, penalty, viol_int, obj_int, w.p_full, in.pos_of_col, in.csr_rowptr, in.csr_col, in.csr_val,
                                          in.csr_val_int, in.b, in.b_int, in.c_int, B, in.L, in.n, in.M, in.nnz, st));
    }
    return 0;
}

}  // namespace dip

extern "C" {

int dip_engine_create(dip_engine** out) {
    DIP_REQUIRE(out, "dip_engine_create: null out");
    int dev = 0;
    DIP_CUDA(cudaGetDevice(&dev));
    DIP_TRY(dip_check_device(dev, nullptr, nullptr, nullptr));
    *out = new dip_engine();
    memset(&(*out)->inst, 0, sizeof(dip_instance));
    return 0;
}

void dip_engine_destroy(dip_engine* e) {
    if (!e) return;
    e->den_pool.release();
    e->dec_pool.release();
    e->inst_pool.release();
    e->wave_pool.release();
    if (e->table) cudaFree(e->table);
    delete e;
}

int dip_engine_set_denoiser(dip_engine* e, const float* w1, const float* b1, const float* w2, const float* b2,
                            const dip_block_weights* blocks, int n_layers, dip_stream_t stream) {
    DIP_REQUIRE(e && w1 && b1 && w2 && b2 && blocks, "dip_engine_set_denoiser: null argument");
    DIP_REQUIRE(n_layers >= 0 && n_layers <= MAX_LAYERS, "dip_engine_set_denoiser: n_layers %d out of range", n_layers);
    cudaStream_t st = as_stream(stream);
    DIP_CUDA(cudaStreamSynchronize(st));
    e->den_pool.release();
    e->has_den = false;
    e->mproj_valid = false;
    DIP_TRY(e->den_pool.alloc(&e->w1, 2 * D * D));
    DIP_TRY(e->den_pool.alloc(&e->w1m, D * D));
    DIP_TRY(e->den_pool.alloc(&e->w1x.w, D * D));
    DIP_TRY(e->den_pool.alloc(&e->b1, D));
    DIP_TRY(e->den_pool.alloc(&e->b2, D));
    DIP_CUDA(cudaMemcpyAsync(e->w1, w1, 2 * D * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // split linear_1 (128, 256) into its mip half [:, :128] and its x half [:, 128:] (diffusion.py:156)
    DIP_CUDA(cudaMemcpy2DAsync(e->w1m, D * sizeof(float), w1, 2 * D * sizeof(float), D * sizeof(float), D, cudaMemcpyDeviceToDevice, st));
    DIP_CUDA(cudaMemcpy2DAsync(e->w1x.w, D * sizeof(float), w1 + D, 2 * D * sizeof(float), D * sizeof(float), D, cudaMemcpyDeviceToDevice, st));
    DIP_TRY(make_planes(e->den_pool, e->w1x, D * D, st));
    DIP_CUDA(cudaMemcpyAsync(e->b1, b1, D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    DIP_TRY(copy_weight(e->den_pool, e->w2, w2, D * D, st));
    DIP_CUDA(cudaMemcpyAsync(e->b2, b2, D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    for (int l = 0; l < n_layers; ++l) DIP_TRY(copy_block(e->den_pool, e->den[l], blocks[l], st));
    e->n_den = n_layers;
    e->has_den = true;
    return 0;
}

int dip_engine_set_decoder(dip_engine* e, const dip_block_weights* blocks, int n_layers, const float* proj_w, const float* proj_b,
                           dip_stream_t stream) {
    DIP_REQUIRE(e && blocks && proj_w && proj_b, "dip_engine_set_decoder: null argument");
    DIP_REQUIRE(n_layers >= 1 && n_layers <= MAX_LAYERS, "dip_engine_set_decoder: n_layers %d out of range", n_layers);
    cudaStream_t st = as_stream(stream);
    DIP_CUDA(cudaStreamSynchronize(st));
    e->dec_pool.release();
    e->has_dec = false;
    e->dec0_ws = nullptr;
    DIP_TRY(e->dec_pool.alloc(&e->proj_w, D));
    DIP_TRY(e->dec_pool.alloc(&e->proj_b, 1));
    DIP_CUDA(cudaMemcpyAsync(e->proj_w, proj_w, D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    DIP_CUDA(cudaMemcpyAsync(e->proj_b, proj_b, sizeof(float), cudaMemcpyDeviceToDevice, st));
    for (int l = 0; l < n_layers; ++l) DIP_TRY(copy_block(e->dec_pool, e->dec[l], blocks[l], st));
    e->n_dec = n_layers;
    e->has_dec = true;
    return 0;
}

static int validate_instance(const dip_instance* inst) {
    DIP_REQUIRE(inst->L > 0 && inst->n > 0 && inst->n <= inst->L && inst->M > 0 && inst->nnz >= 0, "dip_engine_set_instance: bad sizes L=%d n=%d M=%d",
                inst->L, inst->n, inst->M);
    DIP_REQUIRE(inst->mip && inst->kpm && inst->pos_of_col && inst->csr_rowptr && inst->csr_col && inst->csr_val && inst->csc_colptr &&
                    inst->csc_row && inst->csc_val && inst->b && inst->c,
                "dip_engine_set_instance: null array");
    return 0;
}

int dip_engine_set_instances(dip_engine* e, const dip_instance* insts, int n_inst, dip_stream_t stream) {
    DIP_REQUIRE(e && insts && n_inst >= 1, "dip_engine_set_instances: null / empty argument");
    for (int i = 0; i < n_inst; ++i) {
        DIP_TRY(validate_instance(&insts[i]));
        DIP_REQUIRE(insts[i].L == insts[0].L, "dip_engine_set_instances: instance %d has padding_len %d, instance 0 has %d", i, insts[i].L, insts[0].L);
    }
    const dip_instance* inst = &insts[0];
    cudaStream_t st = as_stream(stream);
    if (inst->L > e->inst_cap_L) {
        DIP_CUDA(cudaStreamSynchronize(st));
        e->inst_pool.release();
        float* m1 = nullptr; float* m2 = nullptr;
        DIP_TRY(e->inst_pool.alloc(&e->mproj, (size_t)inst->L * D));
        DIP_TRY(e->inst_pool.alloc(&m1, (inst->L + 1 + 3) / 4 + 64));
        DIP_TRY(e->inst_pool.alloc(&m2, (2 * inst->L + 3) / 4 + 64));
        e->mask_den = reinterpret_cast<uint8_t*>(m1);
        e->mask_dec = reinterpret_cast<uint8_t*>(m2);
        e->inst_cap_L = inst->L;
    }
    e->inst = *inst;
    e->insts.assign(insts, insts + n_inst);
    e->n_max = 0; e->M_max = 0;
    for (int i = 0; i < n_inst; ++i) {
        e->n_max = insts[i].n > e->n_max ? insts[i].n : e->n_max;
        e->M_max = insts[i].M > e->M_max ? insts[i].M : e->M_max;
    }
    // key masks: denoiser [False | kpm] (diffusion.py:162-165), decoder [kpm | kpm] (decoder.py:42)
    DIP_CUDA(cudaMemsetAsync(e->mask_den, 0, 1, st));
    DIP_CUDA(cudaMemcpyAsync(e->mask_den + 1, inst->kpm, inst->L, cudaMemcpyDeviceToDevice, st));
    DIP_CUDA(cudaMemcpyAsync(e->mask_dec, inst->kpm, inst->L, cudaMemcpyDeviceToDevice, st));
    DIP_CUDA(cudaMemcpyAsync(e->mask_dec + inst->L, inst->kpm, inst->L, cudaMemcpyDeviceToDevice, st));
    if (n_inst > 1) {
        if (n_inst > e->table_cap) {
            DIP_CUDA(cudaStreamSynchronize(st));
            if (e->table) cudaFree(e->table);
            e->table = nullptr;
            DIP_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->table), (size_t)n_inst * sizeof(dip_instance)));
            e->table_cap = n_inst;
        }
        // pageable host source: the copy is staged before the call returns, the vector may change afterwards
        DIP_CUDA(cudaMemcpyAsync(e->table, e->insts.data(), (size_t)n_inst * sizeof(dip_instance), cudaMemcpyHostToDevice, st));
    }
    e->wave_B = 0;                          // a map refers to the previous set of instances
    e->has_inst = true;
    e->mproj_valid = false;
    e->dec0_ws = nullptr;
    return 0;
}

int dip_engine_set_instance(dip_engine* e, const dip_instance* inst, dip_stream_t stream) {
    DIP_REQUIRE(e && inst, "dip_engine_set_instance: null argument");
    return dip_engine_set_instances(e, inst, 1, stream);
}

int dip_engine_set_wave_map(dip_engine* e, const int32_t* inst_of, int B, dip_stream_t stream) {
    DIP_REQUIRE(e && e->has_inst, "dip_engine_set_wave_map: set the instances first");
    cudaStream_t st = as_stream(stream);
    if (inst_of == nullptr || B <= 0) {
        DIP_REQUIRE(e->insts.size() == 1, "dip_engine_set_wave_map: several instances are resident, a map is required");
        e->wave_B = 0;
        e->mproj_valid = false;
        e->dec0_ws = nullptr;
        return 0;
    }
    const int n_inst = (int)e->insts.size();
    for (int s = 0; s < B; ++s) DIP_REQUIRE(inst_of[s] >= 0 && inst_of[s] < n_inst, "dip_engine_set_wave_map: sample %d -> instance %d of %d", s, inst_of[s], n_inst);
    const int64_t L = e->inst.L;
    if (B > e->wave_cap) {
        DIP_CUDA(cudaStreamSynchronize(st));
        e->wave_pool.release();
        float *io = nullptr, *k = nullptr, *m1 = nullptr, *m2 = nullptr;
        DIP_TRY(e->wave_pool.alloc(&e->wave_mip, (size_t)B * L * D));
        DIP_TRY(e->wave_pool.alloc(&e->wave_mproj, (size_t)B * L * D));
        DIP_TRY(e->wave_pool.alloc(&io, (size_t)B));
        DIP_TRY(e->wave_pool.alloc(&k, (size_t)B * L / 4 + 64));
        DIP_TRY(e->wave_pool.alloc(&m1, (size_t)B * (L + 1) / 4 + 64));
        DIP_TRY(e->wave_pool.alloc(&m2, (size_t)B * 2 * L / 4 + 64));
        e->inst_of = reinterpret_cast<int32_t*>(io);
        e->wave_kpm = reinterpret_cast<uint8_t*>(k);
        e->wave_mask_den = reinterpret_cast<uint8_t*>(m1);
        e->wave_mask_dec = reinterpret_cast<uint8_t*>(m2);
        e->wave_cap = B;
    }
    if (e->table == nullptr) {              // one resident instance used through a map: still needs the device table
        DIP_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->table), sizeof(dip_instance)));
        e->table_cap = 1;
    }
    if (n_inst == 1) DIP_CUDA(cudaMemcpyAsync(e->table, e->insts.data(), sizeof(dip_instance), cudaMemcpyHostToDevice, st));
    DIP_CUDA(cudaMemcpyAsync(e->inst_of, inst_of, (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    // per-sample gathers, once per map (the step loop then reads plain strided tensors; the reference builds the same with .repeat())
    DIP_CUDA(cudaMemsetAsync(e->wave_mask_den, 0, (size_t)B * (L + 1), st));
    for (int s = 0; s < B; ++s) {
        const dip_instance& in = e->insts[inst_of[s]];
        DIP_CUDA(cudaMemcpyAsync(e->wave_mip + (size_t)s * L * D, in.mip, (size_t)L * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
        DIP_CUDA(cudaMemcpyAsync(e->wave_kpm + (size_t)s * L, in.kpm, (size_t)L, cudaMemcpyDeviceToDevice, st));
        DIP_CUDA(cudaMemcpyAsync(e->wave_mask_den + (size_t)s * (L + 1) + 1, in.kpm, (size_t)L, cudaMemcpyDeviceToDevice, st));
        DIP_CUDA(cudaMemcpyAsync(e->wave_mask_dec + (size_t)s * 2 * L, in.kpm, (size_t)L, cudaMemcpyDeviceToDevice, st));
        DIP_CUDA(cudaMemcpyAsync(e->wave_mask_dec + (size_t)s * 2 * L + L, in.kpm, (size_t)L, cudaMemcpyDeviceToDevice, st));
    }
    e->wave_B = B;
    e->mproj_valid = false;
    e->dec0_ws = nullptr;
    return 0;
}

int dip_engine_n_max(const dip_engine* e) { return e && e->has_inst ? e->n_max : 0; }

int dip_engine_set_precision(dip_engine* e, int mode) {
    DIP_REQUIRE(e, "null engine");
    DIP_REQUIRE(mode == DIP_PRECISION_FP32_SIMT || mode == DIP_PRECISION_TC_SPLIT, "unknown precision mode %d", mode);
    e->mode = mode;
    e->dec0_ws = nullptr;
    return 0;
}

int dip_engine_reset_cache(dip_engine* e) {
    DIP_REQUIRE(e, "null engine");
    e->dec0_ws = nullptr;
    return 0;
}

size_t dip_engine_workspace_bytes(const dip_engine* e, int B) {
    if (!e || !e->has_inst || B <= 0) return 0;
    return carve(e, B, nullptr).bytes;
}

int dip_engine_denoise(dip_engine* e, float* out, const float* x, int B, float t, void* workspace, size_t workspace_bytes,
                       dip_stream_t stream) {
    DIP_TRY(check_ready(e, true, false, B, workspace_bytes, workspace));
    DIP_REQUIRE(out && x, "dip_engine_denoise: null argument");
    Workspace w = carve(e, B, workspace);
    DIP_TRY(denoise(e, w, x, B, t, stream));
    const int L = e->inst.L;
    DIP_CUDA(cudaMemcpy2DAsync(out, (size_t)L * D * sizeof(float), w.hd + D, (size_t)(L + 1) * D * sizeof(float), (size_t)L * D * sizeof(float), B,
                               cudaMemcpyDeviceToDevice, as_stream(stream)));
    return 0;
}

int dip_engine_decode(dip_engine* e, float* p_full, float* y_last, const float* x, int B, void* workspace, size_t workspace_bytes,
                      dip_stream_t stream) {
    DIP_TRY(check_ready(e, false, true, B, workspace_bytes, workspace));
    DIP_REQUIRE(p_full && x, "dip_engine_decode: null argument");
    Workspace w = carve(e, B, workspace);
    DIP_TRY(final_decode(e, w, x, B, p_full, nullptr, nullptr, nullptr, nullptr, stream));
    if (y_last) {
        const int L = e->inst.L;
        DIP_CUDA(cudaMemcpy2DAsync(y_last, (size_t)L * D * sizeof(float), w.ls[e->n_dec - 1].xout + (size_t)L * D, (size_t)2 * L * D * sizeof(float),
                                   (size_t)L * D * sizeof(float), B, cudaMemcpyDeviceToDevice, as_stream(stream)));
    }
    return 0;
}

int dip_engine_guidance_grad(dip_engine* e, float* grad, float* p_full, float* viol, float* obj, const float* x, int B, float lam,
                             void* workspace, size_t workspace_bytes, dip_stream_t stream) {
    DIP_TRY(check_ready(e, false, true, B, workspace_bytes, workspace));
    DIP_REQUIRE(grad && x, "dip_engine_guidance_grad: null argument");
    Workspace w = carve(e, B, workspace);
    const float* g = nullptr;
    int64_t gs = 0;
    DIP_TRY(guidance_grad(e, w, x, B, lam, &g, &gs, stream));
    const int L = e->inst.L;
    cudaStream_t st = as_stream(stream);
    DIP_CUDA(cudaMemcpy2DAsync(grad, (size_t)L * D * sizeof(float), g, (size_t)gs * sizeof(float), (size_t)L * D * sizeof(float), B,
                               cudaMemcpyDeviceToDevice, st));
    if (p_full) DIP_CUDA(cudaMemcpyAsync(p_full, w.p_full, (size_t)B * L * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (viol) DIP_CUDA(cudaMemcpyAsync(viol, w.viol, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (obj) DIP_CUDA(cudaMemcpyAsync(obj, w.obj, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int dip_engine_guided_ddim_step(dip_engine* e, float* x, float* x0_out, int B, const dip_ddim_coeffs* co, float gs, float lam,
                                int eps_param, const float* noise, void* workspace, size_t workspace_bytes, dip_stream_t stream) {
    DIP_TRY(check_ready(e, true, true, B, workspace_bytes, workspace));
    DIP_REQUIRE(x && co, "dip_engine_guided_ddim_step: null argument");
    DIP_REQUIRE(noise || co->sigma == 0.0f, "dip_engine_guided_ddim_step: sigma != 0 needs noise");
    Workspace w = carve(e, B, workspace);
    const int L = e->inst.L;
    DIP_TRY(denoise(e, w, x, B, co->t, stream));
    const float* g = nullptr;
    int64_t gstride = 0;
    DIP_TRY(guidance_grad(e, w, x, B, lam, &g, &gstride, stream));
    DIP_TRY(dip_ddim_step(x, x0_out, x, w.hd + D, (int64_t)(L + 1) * D, g, gstride, co->sigma != 0.0f ? noise : nullptr, B, L, co->sqrt_at,
                          co->s1m, co->c_dir, co->sqrt_aprev, co->sigma, gs, eps_param, stream));
    return 0;
}

int dip_engine_ddim_step(dip_engine* e, float* x, float* x0_out, int B, const dip_ddim_coeffs* co, int eps_param, const float* noise,
                         void* workspace, size_t workspace_bytes, dip_stream_t stream) {
    DIP_TRY(check_ready(e, true, false, B, workspace_bytes, workspace));
    DIP_REQUIRE(x && co, "dip_engine_ddim_step: null argument");
    DIP_REQUIRE(noise || co->sigma == 0.0f, "dip_engine_ddim_step: sigma != 0 needs noise");
    Workspace w = carve(e, B, workspace);
    const int L = e->inst.L;
    DIP_TRY(denoise(e, w, x, B, co->t, stream));
    DIP_TRY(dip_ddim_step(x, x0_out, x, w.hd + D, (int64_t)(L + 1) * D, nullptr, 0, co->sigma != 0.0f ? noise : nullptr, B, L, co->sqrt_at,
                          co->s1m, co->c_dir, co->sqrt_aprev, co->sigma, 0.0f, eps_param, stream));
    return 0;
}

int dip_engine_final_decode(dip_engine* e, const float* x, int B, float* p_full, uint8_t* xbits, float* penalty, int64_t* viol_int,
                            int64_t* obj_int, void* workspace, size_t workspace_bytes, dip_stream_t stream) {
    DIP_TRY(check_ready(e, false, true, B, workspace_bytes, workspace));
    DIP_REQUIRE(x, "dip_engine_final_decode: null argument");
    Workspace w = carve(e, B, workspace);
    return final_decode(e, w, x, B, p_full, xbits, penalty, viol_int, obj_int, stream);
}

int dip_engine_sample_guided_ddim(dip_engine* e, float* x, int B, const dip_ddim_coeffs* steps, int S, float gs, float lam, int eps_param,
                                  uint8_t* xbits, float* penalty, int64_t* viol_int, int64_t* obj_int, void* workspace,
                                  size_t workspace_bytes, dip_stream_t stream) {
    DIP_REQUIRE(steps && S > 0, "dip_engine_sample_guided_ddim: empty schedule");
    for (int s = 0; s < S; ++s) {
        DIP_REQUIRE(steps[s].sigma == 0.0f, "dip_engine_sample_guided_ddim: eta > 0 needs the per-step entry point with explicit noise");
        DIP_TRY(dip_engine_guided_ddim_step(e, x, nullptr, B, &steps[s], gs, lam, eps_param, nullptr, workspace, workspace_bytes, stream));
    }
    return dip_engine_final_decode(e, x, B, nullptr, xbits, penalty, viol_int, obj_int, workspace, workspace_bytes, stream);
}

int dip_engine_guided_ddpm_step(dip_engine* e, float* x, int B, const dip_ddpm_coeffs* co, float gs, float lam, int eps_param,
                                const float* noise, void* workspace, size_t workspace_bytes, dip_stream_t stream) {
    DIP_TRY(check_ready(e, true, true, B, workspace_bytes, workspace));
    DIP_REQUIRE(x && co, "dip_engine_guided_ddpm_step: null argument");
    DIP_REQUIRE(noise || co->nonzero == 0.0f, "dip_engine_guided_ddpm_step: noise required for t != 0");
    Workspace w = carve(e, B, workspace);
    const int L = e->inst.L;
    DIP_TRY(denoise(e, w, x, B, co->t, stream));
    DIP_TRY(dip_ddpm_mean(w.mean_buf, x, w.hd + D, (int64_t)(L + 1) * D, B, L, co->c1, co->c2, eps_param, co->sqrt_recip, co->sqrt_recipm1, stream));
    const float* g = nullptr;
    int64_t gstride = 0;
    DIP_TRY(guidance_grad(e, w, w.mean_buf, B, lam, &g, &gstride, stream));   // guidance at the posterior mean (diffusion.py:433-435)
    DIP_TRY(dip_ddpm_update(x, w.mean_buf, g, gstride, noise, B, L, co->std, gs, co->nonzero, stream));
    return 0;
}

int dip_engine_ddpm_step(dip_engine* e, float* x, int B, const dip_ddpm_coeffs* co, int eps_param, const float* noise, void* workspace,
                         size_t workspace_bytes, dip_stream_t stream) {
    DIP_TRY(check_ready(e, true, false, B, workspace_bytes, workspace));
    DIP_REQUIRE(x && co, "dip_engine_ddpm_step: null argument");
    DIP_REQUIRE(noise || co->nonzero == 0.0f, "dip_engine_ddpm_step: noise required for t != 0");
    Workspace w = carve(e, B, workspace);
    const int L = e->inst.L;
    DIP_TRY(denoise(e, w, x, B, co->t, stream));
    DIP_TRY(dip_ddpm_mean(w.mean_buf, x, w.hd + D, (int64_t)(L + 1) * D, B, L, co->c1, co->c2, eps_param, co->sqrt_recip, co->sqrt_recipm1, stream));
    DIP_TRY(dip_ddpm_update(x, w.mean_buf, nullptr, 0, noise, B, L, co->std, 0.0f, co->nonzero, stream));
    return 0;
}

}  // extern "C"
