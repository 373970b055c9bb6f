// K2: constraint-and-objective guidance (model/diffusion.py:623-633) and the round + feasibility
// check of the driver tail (sample.py:164-174).  One CTA per sample; the probability vector p and
// the hinge indicator h live in shared memory, the CSR/CSC of A is streamed from L2 (it is shared
// by every sample of the wave).  Rows / columns are reduced by fixed-width lane groups with a
// shuffle tree, so the summation order -- and with it the hinge active set -- is reproducible
// run to run and for any sharding of the samples.
#include <algorithm>
#include <numeric>
#include <vector>
#include <math.h>
#include <string.h>
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace dip {

template <typename T>
__device__ __forceinline__ T group_sum(T v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum in a fixed order; result valid in thread 0
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    T t = 0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += scratch[i];
    return t;
}

// A cluster of GCL CTAs works on one sample: CTA `rank` owns a contiguous slice of the rows (pass 1) and of the columns
// (pass 2); the hinge indicators of the other slices are fetched through distributed shared memory after a cluster
// barrier, and the per-CTA partial sums of viol / obj are added by rank 0 in rank order (fixed order -> deterministic).
constexpr int GCL = 8, GTHREADS = 512;   // 8 x 37 = 296 CTAs, two per SM: the passes are gather-latency bound, more CTAs = more loads in flight (4: 86 us per step)

// lane-group width for a segment pass: the power of two (<= 32) next to the average segment length
__host__ __device__ inline int pick_group(int64_t nnz, int segments) {
    int64_t avg = segments > 0 ? (nnz + segments - 1) / segments : 1;
    int g = 1;
    while (g < 32 && g < avg) g <<= 1;
    return g;
}

// The instance of sample b: `single` (one instance for the whole wave, passed by value) or table[inst_of[b]] (a wave mixing instances).
__global__ void __launch_bounds__(GTHREADS) guidance_kernel(
    float* __restrict__ dz, float* __restrict__ viol, float* __restrict__ obj, float* __restrict__ r_out, float* __restrict__ dp_out,
    const float* __restrict__ p_full, const __grid_constant__ dip_instance single, const dip_instance* __restrict__ table,
    const int32_t* __restrict__ inst_of, float one_minus_lam, float lam, int L, int n_max) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) float sm[];
    float* p_s = sm;
    float* h_s = sm + n_max;
    __shared__ float red[GTHREADS / 32];
    __shared__ float part[2];                       // this CTA's share of viol and obj
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / GCL, tid = threadIdx.x;
    const dip_instance& in = table ? table[inst_of[b]] : single;
    const int n = in.n, M = in.M;
    const int32_t* __restrict__ pos_of_col = in.pos_of_col;
    const int32_t* __restrict__ rowptr = in.csr_rowptr; const int32_t* __restrict__ col = in.csr_col; const float* __restrict__ val = in.csr_val;
    const int32_t* __restrict__ colptr = in.csc_colptr; const int32_t* __restrict__ crow = in.csc_row; const float* __restrict__ cval = in.csc_val;
    const float* __restrict__ bvec = in.b; const float* __restrict__ cvec = in.c;
    const int Gr = pick_group(in.nnz, M), Gc = pick_group(in.nnz, n);
    const int Mq = (M + GCL - 1) / GCL, nq = (n + GCL - 1) / GCL, Lq = (L + GCL - 1) / GCL;
    const int i_lo = rank * Mq, i_hi = min(M, i_lo + Mq);
    const int j_lo = rank * nq, j_hi = min(n, j_lo + nq);

    for (int j = tid; j < n; j += GTHREADS) p_s[j] = p_full[(int64_t)b * L + pos_of_col[j]];
    for (int t = rank * Lq + tid; t < min(L, (rank + 1) * Lq); t += GTHREADS) dz[(int64_t)b * L + t] = 0.f;      // padding positions stay zero
    __syncthreads();

    // r = A p - b ; h ; viol   (rows i_lo .. i_hi-1)
    float v_part = 0.f;
    {
        const int groups = GTHREADS / Gr, gid = tid / Gr, gl = tid % Gr;
        for (int i0 = i_lo; i0 < i_hi; i0 += groups) {
            int i = i0 + gid;
            float acc = 0.f;
            if (i < i_hi)
                for (int e = rowptr[i] + gl; e < rowptr[i + 1]; e += Gr) acc = fmaf(val[e], p_s[col[e]], acc);
            acc = group_sum(acc, Gr);
            if (i < i_hi && gl == 0) {
                float r = acc - bvec[i];
                h_s[i] = r > 0.f ? 1.0f : (r == 0.f ? 0.5f : 0.f);
                v_part += fmaxf(r, 0.f);
                if (r_out) r_out[(int64_t)b * M + i] = r;
            }
        }
    }
    float vt = block_sum(v_part, red);
    if (tid == 0) part[0] = vt;
    cluster.sync();                                  // every slice of h (and the zeroed dz) is published
    for (int q = 0; q < GCL; ++q) {
        if (q == rank) continue;
        const float* remote = cluster.map_shared_rank(h_s, q);
        for (int i = q * Mq + tid; i < min(M, (q + 1) * Mq); i += GTHREADS) h_s[i] = remote[i];
    }
    if (rank == 0 && tid == 0 && viol) {
        float t = 0.f;
        for (int q = 0; q < GCL; ++q) t += *cluster.map_shared_rank(&part[0], q);
        viol[b] = t;
    }
    __syncthreads();

    // dp = (1-lam) A^T h + lam c ; dz = dp * p (1-p) ; obj = p.c   (columns j_lo .. j_hi-1)
    float o_part = 0.f;
    {
        const int groups = GTHREADS / Gc, gid = tid / Gc, gl = tid % Gc;
        for (int j0 = j_lo; j0 < j_hi; j0 += groups) {
            int j = j0 + gid;
            float acc = 0.f;
            if (j < j_hi)
                for (int e = colptr[j] + gl; e < colptr[j + 1]; e += Gc) acc = fmaf(cval[e], h_s[crow[e]], acc);
            acc = group_sum(acc, Gc);
            if (j < j_hi && gl == 0) {
                float pj = p_s[j], cj = cvec[j];
                float dp = one_minus_lam * acc + lam * cj;
                dz[(int64_t)b * L + pos_of_col[j]] = (dp * (1.0f - pj)) * pj;   // sigmoid backward, torch order
                if (dp_out) dp_out[(int64_t)b * n + j] = dp;
                o_part += pj * cj;
            }
        }
    }
    float ot = block_sum(o_part, red);
    if (tid == 0) part[1] = ot;
    cluster.sync();
    if (rank == 0 && tid == 0 && obj) {
        float t = 0.f;
        for (int q = 0; q < GCL; ++q) t += *cluster.map_shared_rank(&part[1], q);
        obj[b] = t;
    }
    cluster.sync();                                  // no CTA leaves while its shared memory may still be read
}

__global__ void __launch_bounds__(256) round_feas_kernel(
    uint8_t* __restrict__ xbits, float* __restrict__ penalty, long long* __restrict__ viol_int, long long* __restrict__ obj_int,
    const float* __restrict__ p_full, const __grid_constant__ dip_instance single, const dip_instance* __restrict__ table,
    const int32_t* __restrict__ inst_of, int L, int n_max) {
    extern __shared__ __align__(16) float sm[];
    float* x_s = sm;
    __shared__ float redf[8];
    __shared__ long long redi[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    const dip_instance& in = table ? table[inst_of[b]] : single;
    const int n = in.n, M = in.M;
    const int32_t* __restrict__ pos_of_col = in.pos_of_col;
    const int32_t* __restrict__ rowptr = in.csr_rowptr; const int32_t* __restrict__ col = in.csr_col; const float* __restrict__ val = in.csr_val;
    const int32_t* __restrict__ val_int = in.csr_val_int; const float* __restrict__ bvec = in.b;
    const int32_t* __restrict__ b_int = in.b_int; const int32_t* __restrict__ c_int = in.c_int;
    const int Gr = pick_group(in.nnz, M);
    long long o_part = 0;
    for (int j = tid; j < n_max; j += blockDim.x) {
        float x = 0.f;
        if (j < n) {
            x = rintf(p_full[(int64_t)b * L + pos_of_col[j]]);     // torch.round: half to even (sample.py:164)
            x_s[j] = x;
            if (c_int) o_part += (long long)c_int[j] * (long long)x;
        }
        if (xbits) xbits[(int64_t)b * n_max + j] = (uint8_t)x;       // rows are n_max wide, zero-padded past the instance's n
    }
    __syncthreads();
    float pen = 0.f;
    long long vi = 0;
    const int groups = blockDim.x / Gr, gid = tid / Gr, gl = tid % Gr;
    for (int i0 = 0; i0 < M; i0 += groups) {
        int i = i0 + gid;
        float acc = 0.f;
        long long acci = 0;
        if (i < M)
            for (int e = rowptr[i] + gl; e < rowptr[i + 1]; e += Gr) {
                float x = x_s[col[e]];
                acc = fmaf(val[e], x, acc);
                if (val_int) acci += (long long)val_int[e] * (long long)x;
            }
        acc = group_sum(acc, Gr);
        acci = group_sum(acci, Gr);
        if (i < M && gl == 0) {
            pen += fmaxf(acc - bvec[i], 0.f);
            if (val_int) {
                long long d = acci - (long long)b_int[i];
                vi += d > 0 ? d : 0;
            }
        }
    }
    float pt = block_sum(pen, redf);
    long long vt = block_sum(vi, redi);
    long long ot = block_sum(o_part, redi);
    if (tid == 0) {
        if (penalty) penalty[b] = pt;
        if (viol_int) viol_int[b] = val_int ? vt : -1;
        if (obj_int) obj_int[b] = c_int ? ot : 0;
    }
}

static int launch_guidance(float* dz, float* viol, float* obj, float* r_out, float* dp_out, const float* p_full, const dip_instance& single,
                           const dip_instance* table, const int32_t* inst_of, float lam, int B, int L, int n_max, int M_max, dip_stream_t stream) {
    size_t smem = (size_t)(n_max + M_max) * sizeof(float);
    DIP_REQUIRE(smem <= 200 * 1024, "dip_guidance: n + M = %d exceeds the shared-memory staging limit", n_max + M_max);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        DIP_CUDA(cudaFuncSetAttribute(guidance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = 200 * 1024;
    }
    // (1 - lam) is formed in double like the reference's Python float and rounded once (diffusion.py:633)
    float oml = (float)(1.0 - (double)lam);
    DIP_PROF(DIP_PROF_GUIDANCE, stream);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)B * GCL);
    cfg.blockDim = dim3(GTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = GCL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    DIP_CUDA(cudaLaunchKernelEx(&cfg, guidance_kernel, dz, viol, obj, r_out, dp_out, p_full, single, table, inst_of, oml, lam, L, n_max));
    DIP_LAUNCHED();
    return 0;
}

static int launch_round_feas(uint8_t* xbits, float* penalty, int64_t* viol_int, int64_t* obj_int, const float* p_full, const dip_instance& single,
                             const dip_instance* table, const int32_t* inst_of, int B, int L, int n_max, dip_stream_t stream) {
    size_t smem = (size_t)n_max * sizeof(float);
    DIP_REQUIRE(smem <= 200 * 1024, "dip_round_feasibility: n too large");
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        DIP_CUDA(cudaFuncSetAttribute(round_feas_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = 200 * 1024;
    }
    DIP_PROF(DIP_PROF_GUIDANCE, stream);
    round_feas_kernel<<<B, 256, smem, as_stream(stream)>>>(xbits, penalty, reinterpret_cast<long long*>(viol_int),
                                                           reinterpret_cast<long long*>(obj_int), p_full, single, table, inst_of, L, n_max);
    DIP_LAUNCHED();
    return 0;
}

}  // namespace dip

using namespace dip;

extern "C" {

int dip_guidance(float* dz, float* viol, float* obj, float* r_out, float* dp_out, const float* p_full, const int32_t* pos_of_col,
                 const int32_t* csr_rowptr, const int32_t* csr_col, const float* csr_val, const int32_t* csc_colptr,
                 const int32_t* csc_row, const float* csc_val, const float* b, const float* c, float lam, int B, int L, int n, int M,
                 int64_t nnz, dip_stream_t stream) {
    DIP_REQUIRE(dz && p_full && pos_of_col && csr_rowptr && csr_col && csr_val && csc_colptr && csc_row && csc_val && b && c,
                "dip_guidance: null argument");
    DIP_REQUIRE(B > 0 && L > 0 && n > 0 && n <= L && M > 0, "dip_guidance: bad sizes B=%d L=%d n=%d M=%d", B, L, n, M);
    dip_instance in;
    memset(&in, 0, sizeof(in));
    in.L = L; in.n = n; in.M = M; in.nnz = nnz; in.pos_of_col = pos_of_col;
    in.csr_rowptr = csr_rowptr; in.csr_col = csr_col; in.csr_val = csr_val;
    in.csc_colptr = csc_colptr; in.csc_row = csc_row; in.csc_val = csc_val; in.b = b; in.c = c;
    return launch_guidance(dz, viol, obj, r_out, dp_out, p_full, in, nullptr, nullptr, lam, B, L, n, M, stream);
}

int dip_guidance_multi(float* dz, float* viol, float* obj, const float* p_full, const dip_instance* table, const int32_t* inst_of, float lam,
                       int B, int L, int n_max, int M_max, dip_stream_t stream) {
    DIP_REQUIRE(dz && p_full && table && inst_of, "dip_guidance_multi: null argument");
    DIP_REQUIRE(B > 0 && L > 0 && n_max > 0 && n_max <= L && M_max > 0, "dip_guidance_multi: bad sizes");
    dip_instance none;
    memset(&none, 0, sizeof(none));
    return launch_guidance(dz, viol, obj, nullptr, nullptr, p_full, none, table, inst_of, lam, B, L, n_max, M_max, stream);
}

int dip_round_feasibility(uint8_t* xbits, float* penalty, int64_t* viol_int, int64_t* obj_int, const float* p_full,
                          const int32_t* pos_of_col, const int32_t* csr_rowptr, const int32_t* csr_col, const float* csr_val,
                          const int32_t* csr_val_int, const float* b, const int32_t* b_int, const int32_t* c_int, int B, int L, int n,
                          int M, int64_t nnz, dip_stream_t stream) {
    DIP_REQUIRE(p_full && pos_of_col && csr_rowptr && csr_col && csr_val && b, "dip_round_feasibility: null argument");
    DIP_REQUIRE(B > 0 && L > 0 && n > 0 && n <= L && M > 0, "dip_round_feasibility: bad sizes");
    DIP_REQUIRE((csr_val_int == nullptr) == (b_int == nullptr), "dip_round_feasibility: csr_val_int and b_int go together");
    dip_instance in;
    memset(&in, 0, sizeof(in));
    in.L = L; in.n = n; in.M = M; in.nnz = nnz; in.pos_of_col = pos_of_col;
    in.csr_rowptr = csr_rowptr; in.csr_col = csr_col; in.csr_val = csr_val; in.b = b;
    in.csr_val_int = csr_val_int; in.b_int = b_int; in.c_int = c_int;
    return launch_round_feas(xbits, penalty, viol_int, obj_int, p_full, in, nullptr, nullptr, B, L, n, stream);
}

int dip_round_feasibility_multi(uint8_t* xbits, float* penalty, int64_t* viol_int, int64_t* obj_int, const float* p_full,
                                const dip_instance* table, const int32_t* inst_of, int B, int L, int n_max, dip_stream_t stream) {
    DIP_REQUIRE(p_full && table && inst_of, "dip_round_feasibility_multi: null argument");
    DIP_REQUIRE(B > 0 && L > 0 && n_max > 0 && n_max <= L, "dip_round_feasibility_multi: bad sizes");
    dip_instance none;
    memset(&none, 0, sizeof(none));
    return launch_round_feas(xbits, penalty, viol_int, obj_int, p_full, none, table, inst_of, B, L, n_max, stream);
}

int dip_coo_to_csr_csc_host(int64_t nnz, const int64_t* rows, const int64_t* cols, const float* vals, int M, int n,
                            int32_t* csr_rowptr, int32_t* csr_col, float* csr_val, int32_t* csc_colptr, int32_t* csc_row,
                            float* csc_val, int64_t* nnz_out) {
    DIP_REQUIRE(nnz >= 0 && M > 0 && n > 0 && csr_rowptr && csc_colptr && nnz_out, "dip_coo_to_csr_csc_host: bad argument");
    DIP_REQUIRE(nnz == 0 || (rows && cols && vals && csr_col && csr_val && csc_row && csc_val), "dip_coo_to_csr_csc_host: null array");
    for (int64_t e = 0; e < nnz; ++e)
        DIP_REQUIRE(rows[e] >= 0 && rows[e] < M && cols[e] >= 0 && cols[e] < n, "dip_coo_to_csr_csc_host: entry %lld out of range", (long long)e);
    std::vector<int64_t> order(nnz);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        return rows[a] != rows[b] ? rows[a] < rows[b] : cols[a] < cols[b];
    });
    // coalesce: sum duplicates in their original order (torch CPU coalesce semantics)
    int64_t m = 0;
    std::vector<int32_t> cr, cc;
    std::vector<float> cv;
    cr.reserve(nnz); cc.reserve(nnz); cv.reserve(nnz);
    for (int64_t s = 0; s < nnz;) {
        int64_t e = s;
        float acc = vals[order[s]];
        while (++e < nnz && rows[order[e]] == rows[order[s]] && cols[order[e]] == cols[order[s]]) acc += vals[order[e]];
        cr.push_back((int32_t)rows[order[s]]); cc.push_back((int32_t)cols[order[s]]); cv.push_back(acc);
        s = e; ++m;
    }
    for (int i = 0; i <= M; ++i) csr_rowptr[i] = 0;
    for (int j = 0; j <= n; ++j) csc_colptr[j] = 0;
    for (int64_t e = 0; e < m; ++e) { csr_rowptr[cr[e] + 1]++; csc_colptr[cc[e] + 1]++; }
    for (int i = 0; i < M; ++i) csr_rowptr[i + 1] += csr_rowptr[i];
    for (int j = 0; j < n; ++j) csc_colptr[j + 1] += csc_colptr[j];
    std::vector<int32_t> fill(csc_colptr, csc_colptr + n);
    for (int64_t e = 0; e < m; ++e) {
        csr_col[e] = cc[e]; csr_val[e] = cv[e];
        int32_t pos = fill[cc[e]]++;
        csc_row[pos] = cr[e]; csc_val[pos] = cv[e];   // rows ascending within a column
    }
    *nnz_out = m;
    return 0;
}

}  // extern "C"
