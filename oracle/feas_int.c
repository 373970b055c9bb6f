/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the driver tail's integer arithmetic
 * (reference: sample.py:164-174, sample_partialsol.py:166-170): round half-to-even, per-sample
 * violation sum_i max((A x)_i - b_i, 0) and objective c.x for A in Z^{M x n}, x in {0,1}^n.
 * Built by oracle/Makefile (gcc); used by tests/ as a second, independent checker of
 * dip_round_feasibility.  Never linked into the product library. */
#include <math.h>
#include <stdint.h>

/* rows/cols/vals: coalesced COO, nnz entries.  p: (B, n) probabilities.  Outputs: xbits (B, n),
 * viol (B), obj (B). */
int oracle_round_feasibility(int64_t B, int64_t n, int64_t M, int64_t nnz, const int64_t* rows, const int64_t* cols,
                             const int32_t* vals, const int32_t* b, const int32_t* c, const float* p, uint8_t* xbits,
                             int64_t* viol, int64_t* obj, int64_t* ax_scratch /* M */) {
    for (int64_t s = 0; s < B; ++s) {
        int64_t o = 0;
        for (int64_t j = 0; j < n; ++j) {
            float x = nearbyintf(p[s * n + j]); /* default rounding mode: half to even, like torch.round */
            xbits[s * n + j] = (uint8_t)x;
            o += (int64_t)c[j] * (int64_t)x;
        }
        for (int64_t i = 0; i < M; ++i) ax_scratch[i] = 0;
        for (int64_t e = 0; e < nnz; ++e) ax_scratch[rows[e]] += (int64_t)vals[e] * (int64_t)xbits[s * n + cols[e]];
        int64_t v = 0;
        for (int64_t i = 0; i < M; ++i) {
            int64_t d = ax_scratch[i] - (int64_t)b[i];
            if (d > 0) v += d;
        }
        viol[s] = v;
        obj[s] = o;
    }
    return 0;
}
