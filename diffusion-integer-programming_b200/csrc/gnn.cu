// K3: bipartite constraint<->variable message passing of the CMSP instance encoder
// (model/modules.py:104-268) as deterministic sorted-CSR gather / segment-reduce.  One warp per
// target node, one float4 (4 of the 128 features) per lane; the neighbour rows are gathered from
// L2/HBM with fully coalesced 512-byte reads and accumulated in edge (CSR) order -- no atomics.
#include "common.cuh"

namespace dip {

int sm_count();

// acc[i,:] = sum_{e in seg(i)} relu(s * (PL[i,:] + we[:] * a_e + PR[nbr_e,:])) ; deg[i] = |seg(i)|
__global__ void __launch_bounds__(256) conv_reduce_kernel(float* __restrict__ acc, float* __restrict__ deg, const float* __restrict__ PL,
                                                          const float* __restrict__ PR, const float* __restrict__ we, float s,
                                                          float a_shift, float a_scale, const int32_t* __restrict__ ptr, const int32_t* __restrict__ nbr,
                                                          const float* __restrict__ a, int n_target) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float4 wv = reinterpret_cast<const float4*>(we)[lane];
    for (int64_t i = warp; i < n_target; i += nwarps) {
        float4 pl = reinterpret_cast<const float4*>(PL + i * D)[lane];
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        int e0 = ptr[i], e1 = ptr[i + 1];
        for (int e = e0; e < e1; ++e) {
            float ae = (a[e] + a_shift) * a_scale;
            float4 pr = reinterpret_cast<const float4*>(PR + (int64_t)nbr[e] * D)[lane];
            sum.x += fmaxf(s * (pl.x + wv.x * ae + pr.x), 0.f);
            sum.y += fmaxf(s * (pl.y + wv.y * ae + pr.y), 0.f);
            sum.z += fmaxf(s * (pl.z + wv.z * ae + pr.z), 0.f);
            sum.w += fmaxf(s * (pl.w + wv.w * ae + pr.w), 0.f);
        }
        reinterpret_cast<float4*>(acc + i * D)[lane] = sum;
        if (lane == 0) deg[i] = (float)(e1 - e0);
    }
}

// y[i,:] = sum_e val_e * X[nbr_e,:] + add[i,:]
__global__ void __launch_bounds__(256) spmm_csr_kernel(float* __restrict__ y, const float* __restrict__ X, const float* __restrict__ add,
                                                       const int32_t* __restrict__ ptr, const int32_t* __restrict__ nbr,
                                                       const float* __restrict__ val, int n_rows) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n_rows; i += nwarps) {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        int e0 = ptr[i], e1 = ptr[i + 1];
        for (int e = e0; e < e1; ++e) {
            float ve = val[e];
            float4 xv = reinterpret_cast<const float4*>(X + (int64_t)nbr[e] * D)[lane];
            sum.x = fmaf(ve, xv.x, sum.x); sum.y = fmaf(ve, xv.y, sum.y);
            sum.z = fmaf(ve, xv.z, sum.z); sum.w = fmaf(ve, xv.w, sum.w);
        }
        if (add) {
            float4 av = reinterpret_cast<const float4*>(add + i * D)[lane];
            sum.x += av.x; sum.y += av.y; sum.z += av.z; sum.w += av.w;
        }
        reinterpret_cast<float4*>(y + i * D)[lane] = sum;
    }
}

static int warp_grid(int64_t rows) {
    int64_t blocks = (rows + 7) / 8, cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace dip

using namespace dip;

extern "C" {

int dip_conv_reduce(float* acc, float* deg, const float* PL, const float* PR, const float* we, float s, float a_shift, float a_scale,
                    const int32_t* ptr, const int32_t* nbr, const float* a, int n_target, dip_stream_t stream) {
    DIP_REQUIRE(acc && deg && PL && PR && we && ptr && nbr && a && n_target > 0, "dip_conv_reduce: bad argument");
    DIP_PROF(DIP_PROF_GNN, stream);
    conv_reduce_kernel<<<warp_grid(n_target), 256, 0, as_stream(stream)>>>(acc, deg, PL, PR, we, s, a_shift, a_scale, ptr, nbr, a, n_target);
    DIP_LAUNCHED();
    return 0;
}

int dip_spmm_csr(float* y, const float* X, const float* add, const int32_t* ptr, const int32_t* nbr, const float* val, int n_rows,
                 dip_stream_t stream) {
    DIP_REQUIRE(y && X && ptr && nbr && val && n_rows > 0, "dip_spmm_csr: bad argument");
    DIP_PROF(DIP_PROF_GNN, stream);
    spmm_csr_kernel<<<warp_grid(n_rows), 256, 0, as_stream(stream)>>>(y, X, add, ptr, nbr, val, n_rows);
    DIP_LAUNCHED();
    return 0;
}

}  // extern "C"
