// LayerNorm over 128 features (model/modules.py:280,288), forward and input-gradient backward.
// One warp per row, one float4 per lane; HBM-bound streaming kernels.
#include "tc_common.cuh"

namespace dip {

int sm_count();

__global__ void __launch_bounds__(256) ln_fwd_kernel(float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     const dip_rows x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     int64_t rows) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float4 gv = reinterpret_cast<const float4*>(gamma)[lane];
    float4 bv = reinterpret_cast<const float4*>(beta)[lane];
    for (int64_t r = warp; r < rows; r += nwarps) {
        float4 v = reinterpret_cast<const float4*>(row_ptr(x, r))[lane];
        float mu = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / D);
        float dx = v.x - mu, dy = v.y - mu, dz = v.z - mu, dw = v.w - mu;
        float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / D);
        float rs = rsqrtf(var + 1e-5f);
        // refine rsqrt (approximation, 2 ulp) with one Newton step for fp32-grade accuracy
        rs = rs * (1.5f - 0.5f * (var + 1e-5f) * rs * rs);
        float4 o = make_float4(dx * rs * gv.x + bv.x, dy * rs * gv.y + bv.y, dz * rs * gv.z + bv.z, dw * rs * gv.w + bv.w);
        reinterpret_cast<float4*>(y + r * D)[lane] = o;
        if (mean_out && lane == 0) {
            mean_out[r] = mu;
            rstd_out[r] = rs;
        }
    }
}

// dx = dres + rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy * gamma, xhat = (x-mu)*rstd
__global__ void __launch_bounds__(256) ln_bwd_kernel(float* __restrict__ dx, const float* __restrict__ dy, const dip_rows x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const float* __restrict__ dres, int64_t rows,
                                                     int T, int t_min) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float4 gv = reinterpret_cast<const float4*>(gamma)[lane];
    // only the rows t >= t_min of every sample: `rows` counts those, i -> full row index r
    const int w = t_min > 0 ? T - t_min : 1;
    for (int64_t i = warp; i < rows; i += nwarps) {
        const int64_t r = t_min > 0 ? (i / w) * T + t_min + (i % w) : i;
        float4 xv = reinterpret_cast<const float4*>(row_ptr(x, r))[lane];
        float4 dv = reinterpret_cast<const float4*>(dy + r * D)[lane];
        float mu = mean[r], rs = rstd[r];
        float h0 = (xv.x - mu) * rs, h1 = (xv.y - mu) * rs, h2 = (xv.z - mu) * rs, h3 = (xv.w - mu) * rs;
        float g0 = dv.x * gv.x, g1 = dv.y * gv.y, g2 = dv.z * gv.z, g3 = dv.w * gv.w;
        float m1 = warp_sum(g0 + g1 + g2 + g3) * (1.0f / D);
        float m2 = warp_sum(g0 * h0 + g1 * h1 + g2 * h2 + g3 * h3) * (1.0f / D);
        float4 o = make_float4(rs * (g0 - m1 - h0 * m2), rs * (g1 - m1 - h1 * m2), rs * (g2 - m1 - h2 * m2), rs * (g3 - m1 - h3 * m2));
        if (dres) {
            float4 rr = reinterpret_cast<const float4*>(dres + r * D)[lane];
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
        }
        reinterpret_cast<float4*>(dx + r * D)[lane] = o;
    }
}

// y = LayerNorm(x) written as bf16 hi/lo planes (rows, 128): the A operand of dip_gemm_tc_planes
__global__ void __launch_bounds__(256) ln_split_kernel(uint2* __restrict__ hi, uint2* __restrict__ lo, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                       const dip_rows x, const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows,
                                                       int T, int t_begin) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float4 g = reinterpret_cast<const float4*>(gamma)[lane], b = reinterpret_cast<const float4*>(beta)[lane];
    const int w = T - t_begin;
    // `rows` counts the live rows (t >= t_begin of every sample); the planes are compact over them, the statistics keep the
    // input's row numbering b*T + t
    for (int64_t i = warp; i < rows; i += nwarps) {
        const int64_t r = t_begin > 0 ? (i / w) * T + t_begin + (i % w) : i;
        const float4 v = reinterpret_cast<const float4*>(row_ptr(x, r))[lane];
        const float mu = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / D);
        const float dx = v.x - mu, dy = v.y - mu, dz = v.z - mu, dw = v.w - mu;
        const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / D);
        float rs = rsqrtf(var + 1e-5f);
        rs = rs * (1.5f - 0.5f * (var + 1e-5f) * rs * rs);
        const float y0 = dx * rs * g.x + b.x, y1 = dy * rs * g.y + b.y, y2 = dz * rs * g.z + b.z, y3 = dw * rs * g.w + b.w;
        uint32_t h0, l0, h1, l1;
        tc::split2(y0, y1, h0, l0);
        tc::split2(y2, y3, h1, l1);
        hi[i * 32 + lane] = make_uint2(h0, h1);
        lo[i * 32 + lane] = make_uint2(l0, l1);
        if (mean_out && lane == 0) { mean_out[r] = mu; rstd_out[r] = rs; }
    }
}

static int ln_grid(int64_t rows) {
    int64_t blocks = (rows + 7) / 8;   // 8 warps per CTA
    int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace dip

using namespace dip;

extern "C" {

int dip_layernorm_fwd(float* y, float* mean, float* rstd, dip_rows x, const float* gamma, const float* beta, int64_t rows,
                      dip_stream_t stream) {
    DIP_REQUIRE(y && x.batch && gamma && beta && rows > 0, "dip_layernorm_fwd: bad argument");
    DIP_REQUIRE((mean == nullptr) == (rstd == nullptr), "dip_layernorm_fwd: mean and rstd go together");
    DIP_REQUIRE(x.split == 0 || (x.shared && x.T > x.split), "dip_layernorm_fwd: bad row source");
    DIP_PROF(DIP_PROF_LN, stream);
    ln_fwd_kernel<<<ln_grid(rows), 256, 0, as_stream(stream)>>>(y, mean, rstd, x, gamma, beta, rows);
    DIP_LAUNCHED();
    return 0;
}

int dip_layernorm_split(void* hi, void* lo, float* mean, float* rstd, dip_rows x, const float* gamma, const float* beta, int64_t rows,
                        int t_begin, dip_stream_t stream) {
    DIP_REQUIRE(hi && lo && x.batch && gamma && beta && rows > 0, "dip_layernorm_split: bad argument");
    DIP_REQUIRE((mean == nullptr) == (rstd == nullptr), "dip_layernorm_split: mean and rstd go together");
    DIP_REQUIRE(x.split == 0 || (x.shared && x.T > x.split), "dip_layernorm_split: bad row source");
    DIP_REQUIRE(t_begin == 0 || (x.T > t_begin && t_begin > 0 && rows % x.T == 0), "dip_layernorm_split: t_begin needs whole samples of T rows");
    DIP_PROF(DIP_PROF_LN, stream);
    const int64_t live = t_begin > 0 ? rows / x.T * (x.T - t_begin) : rows;
    ln_split_kernel<<<ln_grid(live), 256, 0, as_stream(stream)>>>(reinterpret_cast<uint2*>(hi), reinterpret_cast<uint2*>(lo), mean, rstd, x, gamma, beta, live,
                                                                  x.T > 0 ? x.T : 1, t_begin);
    DIP_LAUNCHED();
    return 0;
}

int dip_layernorm_bwd(float* dx, const float* dy, dip_rows x, const float* mean, const float* rstd, const float* gamma,
                      const float* dres, int64_t rows, int t_min, dip_stream_t stream) {
    DIP_REQUIRE(dx && dy && x.batch && mean && rstd && gamma && rows > 0, "dip_layernorm_bwd: bad argument");
    DIP_REQUIRE(t_min == 0 || x.T > 0, "dip_layernorm_bwd: t_min needs rows.T");
    DIP_REQUIRE(t_min == 0 || (t_min < x.T && rows % x.T == 0), "dip_layernorm_bwd: t_min needs whole samples of T rows");
    DIP_PROF(DIP_PROF_LN, stream);
    const int64_t live = t_min > 0 ? rows / x.T * (x.T - t_min) : rows;       // rows with t >= t_min
    ln_bwd_kernel<<<ln_grid(live), 256, 0, as_stream(stream)>>>(dx, dy, x, mean, rstd, gamma, dres, live, x.T > 0 ? x.T : 1, t_min);
    DIP_LAUNCHED();
    return 0;
}

}  // extern "C"
