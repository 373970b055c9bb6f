// K1: fused DDIM / DDPM sampler updates, the counter-based normal generator, the time token and the
// decoder head.  All HBM-bound: float4-vectorised, one pass over every operand, grid-stride loops
// sized as a multiple of the SM count.
#include <stdarg.h>
#include <math.h>
#include <vector>
#include <cuda_bf16.h>
#include "common.cuh"

namespace dip {

thread_local char g_err[1024] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- profiler ----------------------------------------------------------------------------------
struct ProfRec { cudaEvent_t a, b; int cls; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static size_t g_prof_used = 0;

ProfScope::ProfScope(int cls, cudaStream_t s) : slot(-1), st(s) {
    if (!g_prof_on) return;
    if (g_prof_used == g_prof.size()) {
        ProfRec r;
        if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
        g_prof.push_back(r);
    }
    slot = (int)g_prof_used++;
    g_prof[slot].cls = cls;
    cudaEventRecord(g_prof[slot].a, st);
}
ProfScope::~ProfScope() {
    if (slot >= 0) cudaEventRecord(g_prof[slot].b, st);
}

static int g_sm_count = 0;
int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

// ---------------------------------------------------------------------------------------------
// DDIM step.  Explicit __f*_rn intrinsics keep the compiler from contracting a*b+c into FMA so that
// every intermediate rounds exactly like the reference's separate torch kernels.
// ---------------------------------------------------------------------------------------------
template <bool EPS, bool HAS_GRAD, bool HAS_NOISE, bool HAS_X0OUT>
__global__ void __launch_bounds__(256) ddim_step_kernel(
    float4* __restrict__ x_prev, float4* __restrict__ x0_out, const float4* __restrict__ x,
    const float* __restrict__ model_out, int64_t out_stride, const float* __restrict__ grad, int64_t grad_stride,
    const float4* __restrict__ noise, int64_t per_sample4, int64_t total4,
    float sqrt_at, float s1m, float c_dir, float sqrt_aprev, float sigma, float gs) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t b = i / per_sample4;
        int64_t j = i - b * per_sample4;
        float4 xv = x[i];
        float4 ov = *reinterpret_cast<const float4*>(model_out + b * out_stride + j * 4);
        float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (HAS_GRAD) gv = *reinterpret_cast<const float4*>(grad + b * grad_stride + j * 4);
        float4 nv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (HAS_NOISE) nv = noise[i];
        float xs[4] = {xv.x, xv.y, xv.z, xv.w}, os[4] = {ov.x, ov.y, ov.z, ov.w};
        float gsv[4] = {gv.x, gv.y, gv.z, gv.w}, ns[4] = {nv.x, nv.y, nv.z, nv.w};
        float rs[4], x0s[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float e, x0;
            if (!EPS) {
                x0 = os[c];
                e = __fdiv_rn(__fsub_rn(xs[c], __fmul_rn(x0, sqrt_at)), s1m);       // diffusion.py:610
            } else {
                e = os[c];
                x0 = __fdiv_rn(__fsub_rn(xs[c], __fmul_rn(s1m, e)), sqrt_at);       // diffusion.py:616
            }
            if (HAS_GRAD) e = __fsub_rn(e, __fmul_rn(gs, gsv[c]));                  // diffusion.py:638
            float dir = __fmul_rn(c_dir, e);                                        // diffusion.py:641
            float r = __fadd_rn(__fmul_rn(sqrt_aprev, x0), dir);                    // diffusion.py:643
            if (HAS_NOISE) r = __fadd_rn(r, __fmul_rn(sigma, ns[c]));
            rs[c] = r;
            x0s[c] = x0;
        }
        x_prev[i] = make_float4(rs[0], rs[1], rs[2], rs[3]);
        if (HAS_X0OUT) x0_out[i] = make_float4(x0s[0], x0s[1], x0s[2], x0s[3]);
    }
}

template <bool EPS, bool G, bool N>
static void launch_ddim(bool x0o, int grid, cudaStream_t st, float4* xp, float4* x0_out, const float4* x,
                        const float* mo, int64_t os, const float* g, int64_t gstr, const float4* nz, int64_t ps4,
                        int64_t tot4, float a, float b, float c, float d, float e, float f) {
    if (x0o) ddim_step_kernel<EPS, G, N, true><<<grid, 256, 0, st>>>(xp, x0_out, x, mo, os, g, gstr, nz, ps4, tot4, a, b, c, d, e, f);
    else ddim_step_kernel<EPS, G, N, false><<<grid, 256, 0, st>>>(xp, x0_out, x, mo, os, g, gstr, nz, ps4, tot4, a, b, c, d, e, f);
}

__global__ void __launch_bounds__(256) ddpm_mean_kernel(
    float4* __restrict__ mean, const float4* __restrict__ x, const float* __restrict__ model_out, int64_t out_stride,
    int64_t per_sample4, int64_t total4, float c1, float c2, int eps_param, float sqrt_recip, float sqrt_recipm1) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t b = i / per_sample4;
        int64_t j = i - b * per_sample4;
        float4 xv = x[i];
        float4 ov = *reinterpret_cast<const float4*>(model_out + b * out_stride + j * 4);
        float xs[4] = {xv.x, xv.y, xv.z, xv.w}, os[4] = {ov.x, ov.y, ov.z, ov.w}, rs[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float x0 = os[c];
            if (eps_param) x0 = __fsub_rn(__fmul_rn(sqrt_recip, xs[c]), __fmul_rn(sqrt_recipm1, os[c]));  // diffusion.py:487-491
            rs[c] = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xs[c]));                                   // diffusion.py:479-482
        }
        mean[i] = make_float4(rs[0], rs[1], rs[2], rs[3]);
    }
}

__global__ void __launch_bounds__(256) ddpm_update_kernel(
    float4* __restrict__ x_prev, const float4* __restrict__ mean, const float* __restrict__ grad, int64_t grad_stride,
    const float4* __restrict__ noise, int64_t per_sample4, int64_t total4, float stdv, float gs, float nonzero) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t b = i / per_sample4;
        int64_t j = i - b * per_sample4;
        float4 mv = mean[i];
        float ms[4] = {mv.x, mv.y, mv.z, mv.w}, rs[4];
        float gsv[4] = {0.f, 0.f, 0.f, 0.f}, ns[4] = {0.f, 0.f, 0.f, 0.f};
        if (grad) {
            float4 gv = *reinterpret_cast<const float4*>(grad + b * grad_stride + j * 4);
            gsv[0] = gv.x; gsv[1] = gv.y; gsv[2] = gv.z; gsv[3] = gv.w;
        }
        if (noise) {
            float4 nv = noise[i];
            ns[0] = nv.x; ns[1] = nv.y; ns[2] = nv.z; ns[3] = nv.w;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float m = ms[c];
            if (grad) m = __fsub_rn(m, __fmul_rn(__fmul_rn(gs, stdv), gsv[c]));     // diffusion.py:446
            float r = m;
            if (noise) r = __fadd_rn(m, __fmul_rn(__fmul_rn(nonzero, stdv), ns[c]));  // diffusion.py:454
            rs[c] = r;
        }
        x_prev[i] = make_float4(rs[0], rs[1], rs[2], rs[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller.  Thread i produces elements 4i..4i+3 from counter (i, subseq).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__global__ void __launch_bounds__(256) randn_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint64_t subseq) {
    int64_t n4 = (n + 3) / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t c[4] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)subseq, (uint32_t)(subseq >> 32)};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            philox_round(c, k0, k1);
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        float z[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            // u1 in (0,1], u2 in [0,1)
            float u1 = ((float)(c[2 * h] >> 8) + 1.0f) * (1.0f / 16777216.0f);
            float u2 = (float)(c[2 * h + 1] >> 8) * (1.0f / 16777216.0f);
            float rad = sqrtf(-2.0f * logf(u1));
            float sn, cs;
            sincospif(2.0f * u2, &sn, &cs);
            z[2 * h] = rad * cs;
            z[2 * h + 1] = rad * sn;
        }
        int64_t base = i * 4;
        if (base + 3 < n) {
            *reinterpret_cast<float4*>(out + base) = make_float4(z[0], z[1], z[2], z[3]);
        } else {
            for (int c2 = 0; c2 < 4 && base + c2 < n; ++c2) out[base + c2] = z[c2];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// time token: out[b, 0, :] = QuickGELU(W1 . [sin(t f) | cos(t f)] + b1), f_k = exp(-ln(1e4) k/127)
// (model/diffusion.py:55-67,154,160).  One CTA of 128 threads; thread j computes feature j.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) time_token_kernel(float* __restrict__ out, int B, int T, float t,
                                                         const float* __restrict__ W1, const float* __restrict__ b1) {
    __shared__ float emb[2 * D];
    int j = threadIdx.x;
    {
        float exponent = (-9.210340371976184f * (float)j) / 127.0f;   // -log(10000) * k / (half_dim - 1)
        float f = expf(exponent);
        float a = t * f;
        emb[j] = sinf(a);
        emb[D + j] = cosf(a);
    }
    __syncthreads();
    float acc = 0.f;
    // same sequential accumulation order as before; the row is fetched as 64 independent 16-byte loads (the scalar loop spent 23 us per
    // step waiting for one uncoalesced 4-byte load after the other)
    const float4* w4 = reinterpret_cast<const float4*>(W1 + (int64_t)j * 2 * D);
#pragma unroll 16
    for (int k4 = 0; k4 < 2 * D / 4; ++k4) {
        const float4 x = __ldg(w4 + k4);
        acc = fmaf(emb[4 * k4], x.x, acc);
        acc = fmaf(emb[4 * k4 + 1], x.y, acc);
        acc = fmaf(emb[4 * k4 + 2], x.z, acc);
        acc = fmaf(emb[4 * k4 + 3], x.w, acc);
    }
    float v = quick_gelu(acc + b1[j]);
    for (int b = 0; b < B; ++b) out[(int64_t)b * T * D + j] = v;
}

// ---------------------------------------------------------------------------------------------
// decoder head: one warp per (b, t) row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_fwd_kernel(float* __restrict__ p_full, const float* __restrict__ y, int B, int T,
                                                       int L, const float* __restrict__ w, const float* __restrict__ beta,
                                                       const uint8_t* __restrict__ kpm, int64_t kpm_stride) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float4 wv = reinterpret_cast<const float4*>(w)[lane];
    float bb = beta[0];
    for (int64_t r = warp; r < (int64_t)B * L; r += nwarps) {
        int64_t b = r / L;
        int t = (int)(r - b * L);
        float4 yv = reinterpret_cast<const float4*>(y + (b * T + (T - L) + t) * D)[lane];
        float s = yv.x * wv.x + yv.y * wv.y + yv.z * wv.z + yv.w * wv.w;
        s = warp_sum(s);
        if (lane == 0) p_full[r] = kpm[b * kpm_stride + t] ? 0.0f : sigmoidf_(s + bb);
    }
}

__global__ void __launch_bounds__(256) head_bwd_kernel(float4* __restrict__ dy, const float* __restrict__ dz, int B, int T, int L,
                                                       const float* __restrict__ w) {
    int64_t total4 = (int64_t)B * T * (D / 4);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = i / (D / 4);
        int c4 = (int)(i - row * (D / 4));
        int64_t b = row / T;
        int t = (int)(row - b * T);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= T - L) {
            float g = dz[b * L + (t - (T - L))];
            float4 wv = reinterpret_cast<const float4*>(w)[c4];
            o = make_float4(g * wv.x, g * wv.y, g * wv.z, g * wv.w);
        }
        dy[i] = o;
    }
}

__global__ void __launch_bounds__(256) split_bf16_kernel(uint2* __restrict__ hi, uint2* __restrict__ lo, const float4* __restrict__ x, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = x[i];
        float vs[4] = {v.x, v.y, v.z, v.w};
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            __nv_bfloat16 hb = __float2bfloat16_rn(vs[e]);
            __nv_bfloat16 lb = __float2bfloat16_rn(vs[e] - __bfloat162float(hb));
            h[e] = __bfloat16_as_ushort(hb);
            l[e] = __bfloat16_as_ushort(lb);
        }
        hi[i] = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
        lo[i] = make_uint2(l[0] | (l[1] << 16), l[2] | (l[3] << 16));
    }
}

__global__ void __launch_bounds__(256) gather_pad_kernel(float4* __restrict__ out, const float* __restrict__ src,
                                                         const int64_t* __restrict__ idx, int n, int L) {
    int64_t total4 = (int64_t)L * (D / 4);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        int t = (int)(i / (D / 4));
        int c4 = (int)(i - (int64_t)t * (D / 4));
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < n) v = reinterpret_cast<const float4*>(src + idx[t] * D)[c4];
        out[i] = v;
    }
}

static int ew_grid(int64_t work_items) {
    int64_t blocks = (work_items + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace dip

using namespace dip;

extern "C" {

int dip_abi_version(void) { return DIP_ABI_VERSION; }
const char* dip_last_error(void) { return dip::g_err; }
uint64_t dip_launch_count(void) { return dip::g_launches.load(); }

void dip_prof_enable(int on) {
    dip::g_prof_on = on != 0;
    if (on) dip::g_prof_used = 0;
}

int dip_prof_collect(double* ms_per_class, uint64_t* launches_per_class, int n_classes) {
    DIP_REQUIRE(ms_per_class && launches_per_class && n_classes > 0, "dip_prof_collect: bad argument");
    for (int i = 0; i < n_classes; ++i) { ms_per_class[i] = 0.0; launches_per_class[i] = 0; }
    for (size_t i = 0; i < dip::g_prof_used; ++i) {
        dip::ProfRec& r = dip::g_prof[i];
        DIP_CUDA(cudaEventSynchronize(r.b));
        float ms = 0.f;
        DIP_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        if (r.cls >= 0 && r.cls < n_classes) { ms_per_class[r.cls] += ms; launches_per_class[r.cls] += 1; }
    }
    dip::g_prof_used = 0;
    return 0;
}

int dip_check_device(int dev, int* sm, int* major, int* minor) {
    cudaDeviceProp prop;
    DIP_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm) *sm = prop.multiProcessorCount;
    if (major) *major = prop.major;
    if (minor) *minor = prop.minor;
    DIP_REQUIRE(prop.major == 10, "libdip_b200 holds sm_100a code only; device %d is sm_%d%d", dev, prop.major, prop.minor);
    return 0;
}

int dip_ddim_step(float* x_prev, float* x0_out, const float* x, const float* model_out, int64_t out_stride,
                  const float* grad, int64_t grad_stride, const float* noise, int B, int L, float sqrt_at, float s1m,
                  float c_dir, float sqrt_aprev, float sigma, float gs, int eps_param, dip_stream_t stream) {
    DIP_REQUIRE(x_prev && x && model_out && B > 0 && L > 0, "dip_ddim_step: null/empty argument");
    DIP_REQUIRE(out_stride % 4 == 0 && grad_stride % 4 == 0, "dip_ddim_step: strides must be multiples of 4 floats");
    int64_t ps4 = (int64_t)L * D / 4, tot4 = ps4 * B;
    int grid = ew_grid(tot4);
    cudaStream_t st = as_stream(stream);
    DIP_PROF(DIP_PROF_UPDATE, stream);
    bool g = grad != nullptr, nz = noise != nullptr, x0o = x0_out != nullptr;
    auto xp = reinterpret_cast<float4*>(x_prev);
    auto x0p = reinterpret_cast<float4*>(x0_out);
    auto xi = reinterpret_cast<const float4*>(x);
    auto np = reinterpret_cast<const float4*>(noise);
#define DIP_DDIM(E, G, N) launch_ddim<E, G, N>(x0o, grid, st, xp, x0p, xi, model_out, out_stride, grad, grad_stride, np, ps4, tot4, sqrt_at, s1m, c_dir, sqrt_aprev, sigma, gs)
    if (eps_param) {
        if (g && nz) DIP_DDIM(true, true, true); else if (g) DIP_DDIM(true, true, false);
        else if (nz) DIP_DDIM(true, false, true); else DIP_DDIM(true, false, false);
    } else {
        if (g && nz) DIP_DDIM(false, true, true); else if (g) DIP_DDIM(false, true, false);
        else if (nz) DIP_DDIM(false, false, true); else DIP_DDIM(false, false, false);
    }
#undef DIP_DDIM
    DIP_LAUNCHED();
    return 0;
}

int dip_ddpm_mean(float* mean, const float* x, const float* model_out, int64_t out_stride, int B, int L, float c1, float c2,
                  int eps_param, float sqrt_recip, float sqrt_recipm1, dip_stream_t stream) {
    DIP_REQUIRE(mean && x && model_out && B > 0 && L > 0, "dip_ddpm_mean: null/empty argument");
    int64_t ps4 = (int64_t)L * D / 4, tot4 = ps4 * B;
    DIP_PROF(DIP_PROF_UPDATE, stream);
    ddpm_mean_kernel<<<ew_grid(tot4), 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(mean), reinterpret_cast<const float4*>(x),
                                                                   model_out, out_stride, ps4, tot4, c1, c2, eps_param, sqrt_recip, sqrt_recipm1);
    DIP_LAUNCHED();
    return 0;
}

int dip_ddpm_update(float* x_prev, const float* mean, const float* grad, int64_t grad_stride, const float* noise, int B, int L,
                    float stdv, float gs, float nonzero, dip_stream_t stream) {
    DIP_REQUIRE(x_prev && mean && B > 0 && L > 0, "dip_ddpm_update: null/empty argument");
    int64_t ps4 = (int64_t)L * D / 4, tot4 = ps4 * B;
    DIP_PROF(DIP_PROF_UPDATE, stream);
    ddpm_update_kernel<<<ew_grid(tot4), 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(x_prev), reinterpret_cast<const float4*>(mean),
                                                                     grad, grad_stride, reinterpret_cast<const float4*>(noise), ps4, tot4, stdv, gs, nonzero);
    DIP_LAUNCHED();
    return 0;
}

int dip_randn(float* out, int64_t n, uint64_t seed, uint64_t subseq, dip_stream_t stream) {
    DIP_REQUIRE(out && n > 0, "dip_randn: null/empty argument");
    DIP_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "dip_randn: output must be 16-byte aligned");
    randn_kernel<<<ew_grid((n + 3) / 4), 256, 0, as_stream(stream)>>>(out, n, seed, subseq);
    DIP_LAUNCHED();
    return 0;
}

int dip_time_token(float* out, int B, int T, float t, const float* W1, const float* b1, dip_stream_t stream) {
    DIP_REQUIRE(out && W1 && b1 && B > 0 && T > 0, "dip_time_token: null/empty argument");
    DIP_PROF(DIP_PROF_OTHER, stream);
    time_token_kernel<<<1, 128, 0, as_stream(stream)>>>(out, B, T, t, W1, b1);
    DIP_LAUNCHED();
    return 0;
}

int dip_decoder_head_fwd(float* p_full, const float* y, int B, int T, int L, const float* w, const float* beta,
                         const uint8_t* kpm, int64_t kpm_stride, dip_stream_t stream) {
    DIP_REQUIRE(p_full && y && w && beta && kpm && B > 0 && L > 0 && T >= L && kpm_stride >= 0, "dip_decoder_head_fwd: bad argument");
    DIP_PROF(DIP_PROF_HEAD, stream);
    head_fwd_kernel<<<ew_grid((int64_t)B * L * 32), 256, 0, as_stream(stream)>>>(p_full, y, B, T, L, w, beta, kpm, kpm_stride);
    DIP_LAUNCHED();
    return 0;
}

int dip_decoder_head_bwd(float* dy, const float* dz, int B, int T, int L, const float* w, dip_stream_t stream) {
    DIP_REQUIRE(dy && dz && w && B > 0 && L > 0 && T >= L, "dip_decoder_head_bwd: bad argument");
    DIP_PROF(DIP_PROF_HEAD, stream);
    head_bwd_kernel<<<ew_grid((int64_t)B * T * (D / 4)), 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(dy), dz, B, T, L, w);
    DIP_LAUNCHED();
    return 0;
}

int dip_split_bf16(void* hi, void* lo, const float* x, int64_t n, dip_stream_t stream) {
    DIP_REQUIRE(hi && lo && x && n > 0 && n % 4 == 0, "dip_split_bf16: bad argument");
    DIP_PROF(DIP_PROF_OTHER, stream);
    split_bf16_kernel<<<ew_grid(n / 4), 256, 0, as_stream(stream)>>>(reinterpret_cast<uint2*>(hi), reinterpret_cast<uint2*>(lo),
                                                                     reinterpret_cast<const float4*>(x), n / 4);
    DIP_LAUNCHED();
    return 0;
}

int dip_gather_pad(float* out, const float* src, const int64_t* idx, int n, int L, dip_stream_t stream) {
    DIP_REQUIRE(out && src && idx && n >= 0 && L >= n, "dip_gather_pad: bad argument");
    gather_pad_kernel<<<ew_grid((int64_t)L * (D / 4)), 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(out), src, idx, n, L);
    DIP_LAUNCHED();
    return 0;
}

}  // extern "C"
