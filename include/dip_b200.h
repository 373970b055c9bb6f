/*
 * dip_b200.h -- C-ABI of the B200-native guided DDIM/DDPM sampling path for integer programs.
 *
 * One shared library (libdip_b200.so, built from diffusion-integer-programming_b200/csrc by
 * __graft_entry__.build()) exports everything below with C linkage: plain pointers and sizes,
 * no torch types.  Device pointers are ordinary CUDA device addresses (e.g. tensor.data_ptr());
 * `stream` is a cudaStream_t passed as void*.  Every function returns 0 on success and a non-zero
 * code on failure; dip_last_error() then holds a message (thread-local).  Nothing here falls back
 * to the CPU: a missing GPU / failed launch is an error.
 *
 * The reference (agent-lab/diffusion-integer-programming) is pure Python on torch ops -- it has no
 * FFI of its own.  Each entry point therefore cites the reference Python lines whose torch ops it
 * replaces (paths relative to the reference root); INTEGRATION.md shows the ctypes binding a
 * maintainer adds inside model/diffusion.py, model/decoder.py and model/cmsp.py.
 *
 * Fixed model width: emb_dim = 128, one attention head (sample.py:21-37).  All floating-point
 * tensors are fp32, row-major, innermost dimension DIP_D.
 */
#ifndef DIP_B200_H
#define DIP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIP_D 128
#define DIP_ABI_VERSION 2

typedef void* dip_stream_t;

/* ------------------------------------------------------------------------------------------ */
/* library                                                                                    */
/* ------------------------------------------------------------------------------------------ */
int dip_abi_version(void);
const char* dip_last_error(void);
/* Fails unless device `dev` is compute capability 10.x (the library holds sm_100a code only). */
int dip_check_device(int dev, int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------ */
/* host-side helper (CPU, no GPU needed)                                                      */
/* ------------------------------------------------------------------------------------------ */
/* Replaces the two A.coalesce() sorts the reference runs on EVERY step (model/diffusion.py:624):
 * sorts COO by (row, col), sums duplicates in input order, emits CSR (rowptr[M+1], col[nnz'],
 * val[nnz']) and CSC (colptr[n+1], row[nnz'], val[nnz']).  Output arrays must hold `nnz` entries;
 * *nnz_out receives the coalesced count. */
int dip_coo_to_csr_csc_host(int64_t nnz, const int64_t* rows, const int64_t* cols, const float* vals,
                            int M, int n, int32_t* csr_rowptr, int32_t* csr_col, float* csr_val,
                            int32_t* csc_colptr, int32_t* csc_row, float* csc_val, int64_t* nnz_out);

/* ------------------------------------------------------------------------------------------ */
/* row sources                                                                                */
/* ------------------------------------------------------------------------------------------ */
/* A (B*T, 128) activation whose first `split` rows of every sample come from one batch-shared
 * (split,128) tensor and the remaining T-split rows from a per-sample (B, T-split, 128) tensor.
 * This is how the seq-concat torch.concat([mip, x], dim=1) of model/decoder.py:40 is consumed
 * without being materialised.  split == 0: an ordinary contiguous (B*T,128) tensor in `batch`. */
typedef struct {
    const float* shared;
    const float* batch;
    int T;
    int split;
    /* elements between the shared blocks of consecutive samples: 0 = ONE block for the whole wave (one instance repeated, sample.py:151);
     * split*128 = a (B, split, 128) tensor, i.e. every sample its own conditions -- a wave that mixes instances */
    int64_t shared_stride;
} dip_rows;

/* ------------------------------------------------------------------------------------------ */
/* K1 -- fused sampler updates                                                                */
/* ------------------------------------------------------------------------------------------ */
/* One guided/unguided DDIM update, model/diffusion.py:610,638-643 (guided) and :686-699:
 *   x0-param:  e = (x - out*sqrt_at)/s1m ;  eps-param: e = out, x0 = (x - s1m*e)/sqrt_at
 *   e -= gs*grad (if grad);  x_prev = sqrt_aprev*x0 + c_dir*e + sigma*noise (if noise)
 * fp32, same operation order and rounding as the reference's separate torch ops (no FMA
 * contraction) so the result is bit-identical to the CPU reference for identical inputs.
 * x, x_prev: (B, L*128) contiguous.  model_out / grad: element (b,i) at ptr[b*stride + i].
 * x0_out (nullable): contiguous (B, L*128) copy of pred_x0. */
int dip_ddim_step(float* x_prev, float* x0_out, const float* x, const float* model_out, int64_t out_stride,
                  const float* grad, int64_t grad_stride, const float* noise, int B, int L,
                  float sqrt_at, float s1m, float c_dir, float sqrt_aprev, float sigma, float gs,
                  int eps_param, dip_stream_t stream);

/* DDPM posterior mean, model/diffusion.py:463-485: mean = c1*x0 + c2*x
 * (eps-param: x0 = sqrt_recip*x - sqrt_recipm1*out first, :487-491). */
int dip_ddpm_mean(float* mean, const float* x, const float* model_out, int64_t out_stride, int B, int L,
                  float c1, float c2, int eps_param, float sqrt_recip, float sqrt_recipm1, dip_stream_t stream);
/* DDPM ancestral update, model/diffusion.py:445-454: x_prev = (mean - gs*std*grad) + nonzero*std*noise. */
int dip_ddpm_update(float* x_prev, const float* mean, const float* grad, int64_t grad_stride,
                    const float* noise, int B, int L, float std, float gs, float nonzero, dip_stream_t stream);

/* Counter-based N(0,1) generator (Philox4x32-10 + Box-Muller): element i of stream (seed, subseq)
 * depends only on (seed, subseq, i) -- any sharding of samples over GPUs draws identical noise.
 * Replaces torch.randn_like / noise_like (model/diffusion.py:583, model/utils.py:183-186). */
int dip_randn(float* out, int64_t n, uint64_t seed, uint64_t subseq, dip_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* K4 -- transformer block pieces (fp32 SIMT path)                                            */
/* ------------------------------------------------------------------------------------------ */
#define DIP_ACT_NONE 0
#define DIP_ACT_QUICKGELU 1
#define DIP_ACT_RELU 2
#define DIP_ACT_SIGMOID 3
#define DIP_ACT_MUL_QUICKGELU_GRAD 4   /* C = acc * d/du[u*sigmoid(1.702u)] at u = aux */

/* C = act(A . W^T + bias*row_bias_scale + addmat[t] + residual), the torch.nn.Linear calls of
 * model/modules.py:293-296 and model/diffusion.py:160 with their elementwise neighbours fused.
 * A: (M,K) row-major, lda.  W: (N,K) row-major (nn.Linear layout).  N % 128 == 0, K % 16 == 0.
 * Row remap (rows_in > 0): input row r = b*rows_in + t is written to output row
 * b*rows_out + row_off + t, and addmat row t (rows_in, N) is added -- this is how the feature-concat
 * + time-token concat of model/diffusion.py:156-158 is consumed without being built.
 * residual (nullable via res.batch == NULL): added before the activation.
 * preact (nullable): receives the pre-activation value (saved for the backward pass).
 * aux: operand of DIP_ACT_MUL_QUICKGELU_GRAD, (M,N) with ldc. */
typedef struct {
    const float* A; int lda;
    const float* W;
    const float* bias;
    float* C; int ldc;
    int M, N, K;
    int act;
    int rows_in, rows_out, row_off;
    const float* addmat;
    dip_rows res;
    const float* row_bias_scale;
    float* preact;
    const float* aux;
    /* optional bf16 (hi, lo) planes of the result, x = hi + lo to ~16 mantissa bits, same row/column
     * indexing as C (row stride ldc elements): the operand format of the tensor-core kernels below */
    void* split_hi;
    void* split_lo;
    /* input row window (dip_gemm_tc only; in_T > 0 needs rows_in > 0): GEMM row r = b*rows_in + t reads input row
     * b*in_T + in_off + t (of A, of the row source, and of the LayerNorm statistics).  With rows_out = in_T and
     * row_off = in_off a Linear is applied in place to the tokens in_off .. in_off+rows_in-1 of every sample only --
     * e.g. the last L tokens of the decoder's 2L, the only ones model/decoder.py:45 keeps. */
    int in_T, in_off;
    /* elements between the addmat blocks of consecutive samples (rows_in > 0): 0 = one (rows_in, N) block for every sample */
    int64_t addmat_stride;
} dip_gemm_args;
int dip_gemm_nt(const dip_gemm_args* args, dip_stream_t stream);

/* y = LayerNorm(x) over 128 features, eps 1e-5, affine (model/modules.py:280,288).  mean / rstd
 * (nullable) receive the per-row statistics for the backward pass. */
int dip_layernorm_fwd(float* y, float* mean, float* rstd, dip_rows x, const float* gamma, const float* beta,
                      int64_t rows, dip_stream_t stream);
/* dx = dres + dLayerNorm/dx (dy)   (input gradient only -- no dgamma/dbeta: the sampler never uses
 * weight gradients).  row_begin..row_end restrict the rows written
 * (rows are global indices b*T+t of `x`); rows_per_batch/t_min skip rows with t < t_min. */
int dip_layernorm_bwd(float* dx, const float* dy, dip_rows x, const float* mean, const float* rstd,
                      const float* gamma, const float* dres, int64_t rows, int t_min, dip_stream_t stream);

/* softmax(Q K^T / sqrt(128) + keymask) V  for one head of width 128 -- the
 * F.scaled_dot_product_attention inside nn.MultiheadAttention (model/modules.py:291).
 * q,k,v: row (b*T+t) at ptr + row*ld (ld = 384 for the packed in_proj output).  key_mask: bytes,
 * sample b's T flags start at key_mask + b*mask_stride (stride 0: one mask shared by the wave);
 * non-zero = masked key (key_padding_mask semantics), nullable.  o: (B*T,128) ld 128.
 * lse: (B*T) natural-log row sums (saved for the backward pass). */
int dip_attn_fwd(float* o, float* lse, const float* q, const float* k, const float* v, int ld,
                 const uint8_t* key_mask, int64_t mask_stride, int B, int T, dip_stream_t stream);
/* Input gradients dq, dk, dv (row stride ld_d) from do; deterministic (no atomics).
 * delta_ws: (B*T) floats of scratch. */
int dip_attn_bwd(float* dq, float* dk, float* dv, int ld_d, const float* q, const float* k, const float* v, int ld,
                 const float* o, const float* d_o, const float* lse, const uint8_t* key_mask, int64_t mask_stride,
                 float* delta_ws, int B, int T, dip_stream_t stream);

/* ---- tensor-core (tcgen05 / TMEM / TMA) path, "bf16x2 split" numerics ------------------------- */
/* hi = bf16(x), lo = bf16(x - hi) for n fp32 values (n % 4 == 0). */
int dip_split_bf16(void* hi, void* lo, const float* x, int64_t n, dip_stream_t stream);
/* Same contract as dip_attn_fwd, operands given as the bf16 (hi, lo) planes of the packed (B*T, 384)
 * in_proj output [q | k | v].  Every product is evaluated as hi.hi + hi.lo + lo.hi with fp32
 * accumulation in tensor memory. */
/* q_begin: only the queries t >= q_begin of every sample are computed (o / lse rows below are not written) -- the last
 * decoder layer needs just the L tokens model/decoder.py:45 keeps. */
int dip_attn_fwd_tc(float* o, float* lse, const void* qkv_hi, const void* qkv_lo, const uint8_t* key_mask,
                    int64_t mask_stride, int B, int T, int q_begin, float* prefix_state, int prefix_keys, int prefix_mode,
                    dip_stream_t stream);
/* Key-prefix hoisting: when the first `prefix_keys` tokens of every sample do not change between calls (the instance tokens of the
 * decoder's [mip ; x] sequence, model/decoder.py:40: constants of a wave), a query that is itself one of them sees the same scores
 * against them every time.  prefix_mode 1 ("save", once per wave): run those key tiles for the queries below
 * floor(prefix_keys/128)*128 and keep the softmax state (running maximum, running sums, un-normalised P.V) in `prefix_state`
 * (dip_attn_fwd_tc_state_bytes bytes; o / lse are not written).  prefix_mode 2 ("use"): those queries resume from the state and stream
 * only the remaining keys -- bit-identical to prefix_mode 0, a quarter less work for the layer.  q_begin must be 0. */
size_t dip_attn_fwd_tc_state_bytes(int B, int prefix_keys);

/* Same contract as dip_attn_bwd on split planes; do_hi/do_lo: planes of d_o (B*T,128).  Two tcgen05
 * kernels (key-stationary dK/dV, query-stationary dQ); deterministic. */
int dip_attn_bwd_tc(float* dq, float* dk, float* dv, int ld_d, const void* qkv_hi, const void* qkv_lo,
                    const void* do_hi, const void* do_lo, const float* o, const float* d_o, const float* lse,
                    const uint8_t* key_mask, int64_t mask_stride, float* delta_ws, int B, int T,
                    int do_begin, int q_begin, int k_begin, void* ds_ws, size_t ds_ws_bytes, dip_stream_t stream);
/* ds_ws (nullable, 128-byte aligned, dip_attn_bwd_tc_ds_bytes bytes): scratch for dS^T as bf16 (hi, lo) planes.  When given, the
 * key-stationary kernel -- which forms every dS tile for dK anyway -- also stores it, and dQ = dS K becomes a streaming GEMM over the
 * stored tiles instead of a second kernel that recomputes S and dP (5 instead of 7 T^2 D products per backward); without it the
 * query-stationary recompute kernel runs.  Results agree to fp32 rounding of the accumulation order, both are deterministic. */
size_t dip_attn_bwd_tc_ds_bytes(int B, int T, int do_begin, int k_begin);
/* Row windows of dip_attn_bwd_tc (0, 0, 0 = everything): d_o is taken as zero for queries t < do_begin (those rows of
 * d_o / o / lse are never read); dq is written for t >= q_begin only (q_begin >= do_begin); dk, dv for t >= k_begin
 * only.  Unwritten rows keep their previous content. */

/* LayerNorm(x) (model/modules.py:280,288) written as the bf16 (hi, lo) planes of its result, (rows,128) each: the A operand of
 * dip_gemm_tc_planes.  The q|k|v in-projection uses this pair instead of the fused-LayerNorm dip_gemm_tc: its three 128-column
 * blocks run on different CTAs, which would each repeat the row loads, the LayerNorm and the split. */
/* t_begin > 0: only the tokens t >= t_begin of every sample are normalised (rows = B*T, x.T = T); the planes are then COMPACT,
 * (B*(T-t_begin), 128), while mean / rstd keep the input's row numbering b*T + t. */
int dip_layernorm_split(void* hi, void* lo, float* mean, float* rstd, dip_rows x, const float* gamma, const float* beta,
                        int64_t rows, int t_begin, dip_stream_t stream);
/* C = A . W^T + bias on tcgen05 with A given as bf16 (hi, lo) planes (M,128), fetched by TMA straight into the MMA layout.
 * K == 128, N % 128 == 0, bias-only epilogue; outputs as in dip_gemm_args (C / preact / split planes), including the output row remap
 * rows_in / rows_out / row_off (compact A rows of a token window written in place into full-length buffers). */
int dip_gemm_tc_planes(const dip_gemm_args* args, const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo,
                       dip_stream_t stream);

/* dip_gemm_nt on tcgen05 (bf16x2-split operands): w_hi / w_lo are the bf16 planes of the (N,K) weight
 * (dip_split_bf16 of the nn.Linear weight).  K must be 128 or 384.  Extras over dip_gemm_nt:
 *   a_rows   (nullable) two-source A rows (K == 128), replaces args->A;
 *   ln_gamma/ln_beta (nullable) LayerNorm(128, eps 1e-5) applied to every A row on load (model/modules.py:280,288
 *            fused into the following Linear); ln_mean/ln_rstd (nullable) receive the row statistics;
 *   args->C may be NULL when only the split planes / preact are wanted. */
int dip_gemm_tc(const dip_gemm_args* args, const void* w_hi, const void* w_lo, const dip_rows* a_rows,
                const float* ln_gamma, const float* ln_beta, float* ln_mean, float* ln_rstd, dip_stream_t stream);

/* The row-wise half of one ResidualAttentionBlock in ONE launch (model/modules.py:293-296 after the attention):
 *   x_mid = x_in + attn . W_out^T + b_out ;  u = LN2(x_mid) . W_fc^T + b_fc ;  x_out = x_mid + QuickGELU(u) . W_proj^T + b_proj
 * on the tokens t >= win of every sample (B samples of T tokens; all buffers (B*T,128), rows outside the window untouched).  The three
 * weights (bf16 hi/lo planes of the nn.Linear weights, dip_split_bf16) stay resident in shared memory, the intermediate rows in tensor
 * memory: per row 2 x 512 bytes in, 3 x 512 out instead of 9 row passes of three separate Linears.  x_mid is always written (x_mid may
 * alias x_in's batch part and x_out: in place); mean2 / rstd2 (LayerNorm-2 statistics) and u (pre-activation) are optional saves for
 * dip_block_chain_bwd_tc. */
int dip_block_chain_fwd_tc(float* xout, float* xmid, float* mean2, float* rstd2, float* u, const float* attn, dip_rows xin,
                           const void* out_w_hi, const void* out_w_lo, const float* out_b, const float* ln2_g, const float* ln2_b,
                           const void* fc_w_hi, const void* fc_w_lo, const float* fc_b, const void* proj_w_hi, const void* proj_w_lo,
                           const float* proj_b, int B, int T, int win, dip_stream_t stream);
/* Its input-gradient mirror (what autograd runs under loss.backward(), model/diffusion.py:633-634, without the weight gradients):
 *   dU = (dx_out . W_proj) * gelu'(u) ;  dH = dU . W_fc ;  dx_mid = dx_out + LN2'(dH) ;  d_attn = dx_mid . W_out
 * weights given as the planes of the TRANSPOSED matrices; d_attn also as bf16 planes (nullable) for dip_attn_bwd_tc. */
int dip_block_chain_bwd_tc(float* dxmid, float* dattn, void* dattn_hi, void* dattn_lo, const float* dxout, const float* u,
                           const float* xmid, const float* mean2, const float* rstd2, const float* ln2_g,
                           const void* proj_wT_hi, const void* proj_wT_lo, const void* fc_wT_hi, const void* fc_wT_lo,
                           const void* out_wT_hi, const void* out_wT_lo, int B, int T, int win, dip_stream_t stream);

/* The last piece of a block's input-gradient backward in ONE launch: dx_in = dx_mid + LN1'(dQKV . W_in) -- the gradient of
 * `self.attn(self.ln_1(x), ...)` w.r.t. x (model/modules.py:288-291 under loss.backward(), model/diffusion.py:633-634) plus the residual
 * branch -- on the tokens t >= win of every sample.  dqkv: (B*T, 384) [dQ | dK | dV]; xin: the block's input rows (two-source); mean1 /
 * rstd1: LayerNorm-1 statistics saved by the forward pass (dip_layernorm_split); in_wT: bf16 planes of the TRANSPOSED in_proj_weight,
 * (128, 384).  The three 128-column blocks are three tcgen05 products into one accumulator, the LayerNorm backward is the epilogue. */
int dip_block_in_bwd_tc(float* dxin, const float* dqkv, dip_rows xin, const float* mean1, const float* rstd1, const float* ln1_g,
                        const float* dres, const void* in_wT_hi, const void* in_wT_lo, int B, int T, int win, dip_stream_t stream);

/* Row 0 of every sample of the denoiser's first hidden layer: QuickGELU(W1 . temb(t) + b1) with
 * temb the 256-wide sinusoidal embedding of model/diffusion.py:39-74.  out: (B, T, 128), row 0. */
int dip_time_token(float* out, int B, int T, float t, const float* W1, const float* b1, dip_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* decoder head + K2 guidance                                                                 */
/* ------------------------------------------------------------------------------------------ */
/* p_full[b,t] = sigmoid(y[b, T-L+t,:].w + beta), 0 at padded t (model/decoder.py:47-50).
 * y: (B, T, 128) -- the last L rows of each sample are used.  kpm: bytes, sample b's L flags start at kpm + b*kpm_stride
 * (stride 0: one mask shared by the wave). */
int dip_decoder_head_fwd(float* p_full, const float* y, int B, int T, int L, const float* w, const float* beta,
                         const uint8_t* kpm, int64_t kpm_stride, dip_stream_t stream);
/* dy[b, t, :] = 0 for t < T-L, dz[b, t-(T-L)] * w for the last L rows. */
int dip_decoder_head_bwd(float* dy, const float* dz, int B, int T, int L, const float* w, dip_stream_t stream);

/* The guidance loss and its gradient w.r.t. the decoder logits, model/diffusion.py:623-633:
 *   p = p_full[:, pos_of_col];  r = A p - b;  h = 1[r>0] + .5*1[r==0];
 *   dp = (1-lam) A^T h + lam c;  dz[:, pos_of_col] = dp * p (1-p)   (0 elsewhere)
 *   viol[b] = sum_i max(r_i,0);  obj[b] = p.c;  r_out (nullable): (B,M).
 * CSR/CSC from dip_coo_to_csr_csc_host (device copies).  One CTA per sample; p and h staged in
 * shared memory; per-row / per-column reductions by warp shuffles in a fixed order. */
int dip_guidance(float* dz, float* viol, float* obj, float* r_out, float* dp_out, const float* p_full,
                 const int32_t* pos_of_col, const int32_t* csr_rowptr, const int32_t* csr_col, const float* csr_val,
                 const int32_t* csc_colptr, const int32_t* csc_row, const float* csc_val,
                 const float* b, const float* c, float lam, int B, int L, int n, int M, int64_t nnz,
                 dip_stream_t stream);

/* round + feasibility, sample.py:164-174 / sample_partialsol.py:166-170:
 *   xbits = round-half-even(p) ;  penalty[b] = sum_i max((A x)_i - b_i, 0)  (fp32, reference form)
 * and, when the integer instance (csr_val_int, b_int, c_int) is given, the exact twin
 *   viol_int[b] = sum_i max((A_int x)_i - b_int_i, 0),  obj_int[b] = c_int . x. */
int dip_round_feasibility(uint8_t* xbits, float* penalty, int64_t* viol_int, int64_t* obj_int,
                          const float* p_full, const int32_t* pos_of_col,
                          const int32_t* csr_rowptr, const int32_t* csr_col, const float* csr_val,
                          const int32_t* csr_val_int, const float* b, const int32_t* b_int, const int32_t* c_int,
                          int B, int L, int n, int M, int64_t nnz, dip_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* K3 -- bipartite GNN encoder pieces (model/modules.py:104-268)                              */
/* ------------------------------------------------------------------------------------------ */
/* y = act((x + shift) * scale . W^T + bias) for tiny K (5 / 17 input features): PreNormLayer +
 * first embedding Linear, model/modules.py:219-238.  shift/scale nullable (K entries). */
int dip_linear_smallk(float* y, const float* x, int64_t rows, int K, const float* shift, const float* scale,
                      const float* W, const float* bias, int act, dip_stream_t stream);
/* Half-convolution segment reduce, model/modules.py:203-212 with the per-edge final Linear hoisted
 * after the sum:  acc[i,:] = sum_{e in seg(i)} relu(s * (PL[i,:] + we[:]*a'_e + PR[nbr_e,:])),
 * a'_e = (a_e + a_shift) * a_scale (the edge PreNormLayer, modules.py:228-230), deg[i] = |seg(i)|.
 * seg = CSR over targets (ptr, nbr, raw edge value a_e). */
int dip_conv_reduce(float* acc, float* deg, const float* PL, const float* PR, const float* we, float s,
                    float a_shift, float a_scale, const int32_t* ptr, const int32_t* nbr, const float* a,
                    int n_target, dip_stream_t stream);
/* y[i,:] = sum_e val_e * X[nbr_e,:] + add[i,:]   (A.mm(f) + C, model/modules.py:159-160). */
int dip_spmm_csr(float* y, const float* X, const float* add, const int32_t* ptr, const int32_t* nbr,
                 const float* val, int n_rows, dip_stream_t stream);
/* out[t,:] = t < n ? src[idx[t],:] : 0   (gather int_indices + zero pad, model/cmsp.py:54-60). */
int dip_gather_pad(float* out, const float* src, const int64_t* idx, int n, int L, dip_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* engine -- the S-step guided sampler over one wave of B samples of one instance             */
/* ------------------------------------------------------------------------------------------ */
typedef struct dip_engine dip_engine;

/* One ResidualAttentionBlock's parameters (model/modules.py:274-296), device pointers, reference
 * state_dict layout: in_proj_weight (384,128), out_proj.weight (128,128), c_fc / c_proj (128,128). */
typedef struct {
    const float *ln1_w, *ln1_b, *in_proj_w, *in_proj_b, *out_proj_w, *out_proj_b;
    const float *ln2_w, *ln2_b, *fc_w, *fc_b, *proj_w, *proj_b;
} dip_block_weights;

int dip_engine_create(dip_engine** out);
void dip_engine_destroy(dip_engine* e);

/* NoisePredictV2 (model/diffusion.py:120-151): mlp_xt.linear_1 (128,256), linear_2 (128,128), blocks.
 * Weights are copied (and pre-transposed for the backward GEMMs) into engine-owned memory. */
int dip_engine_set_denoiser(dip_engine* e, const float* w1, const float* b1, const float* w2, const float* b2,
                            const dip_block_weights* blocks, int n_layers, dip_stream_t stream);
/* MipConditionalDecorder (model/decoder.py:11-36): blocks + linear_proj (1,128). */
int dip_engine_set_decoder(dip_engine* e, const dip_block_weights* blocks, int n_layers,
                           const float* proj_w, const float* proj_b, dip_stream_t stream);

/* The instance every sample of the wave shares: mip (L,128) conditions (one row of the repeated
 * `conditions` batch of sample.py:151), kpm (L) bytes, pos_of_col (n), CSR + CSC of A (M,n), b (M),
 * c (n); optional exact-integer twin (csr_val_int, b_int, c_int; nullable).  Pointers are BORROWED:
 * they must stay valid until the next dip_engine_set_instance / destroy. */
typedef struct {
    int L, n, M;
    int64_t nnz;
    const float* mip;
    const uint8_t* kpm;
    const int32_t* pos_of_col;
    const int32_t *csr_rowptr, *csr_col; const float* csr_val;
    const int32_t *csc_colptr, *csc_row; const float* csc_val;
    const float *b, *c;
    const int32_t *csr_val_int, *b_int, *c_int;
} dip_instance;
int dip_engine_set_instance(dip_engine* e, const dip_instance* inst, dip_stream_t stream);

/* Waves that mix instances (sample_partialsol.py:55 draws 15 samples per instance and sample.py:134 walks the instances one after the
 * other; the reference has no such batching -- a wave of 15 leaves most of the GPU idle).  `insts`: n_inst instances of ONE family (same L;
 * n, M, nnz may differ), pointers borrowed like dip_engine_set_instance.  `inst_of` (HOST array, B entries): the instance of every sample
 * of the waves that follow; the engine gathers per-sample conditions / masks once per map, not per step.  Per-sample outputs that are
 * (B, n) for one instance become (B, n_max), rows zero-padded.  set_instance / set_instances with n_inst == 1 clear the map. */
int dip_engine_set_instances(dip_engine* e, const dip_instance* insts, int n_inst, dip_stream_t stream);
int dip_engine_set_wave_map(dip_engine* e, const int32_t* inst_of, int B, dip_stream_t stream);
/* largest n over the resident instances (row width of xbits) */
int dip_engine_n_max(const dip_engine* e);

/* K2 and the rounding / feasibility check for a wave that mixes instances: `table` is a DEVICE array of dip_instance (device pointers
 * inside), `inst_of` a DEVICE array with the table index of every sample; p_full / dz are (B, L), xbits (B, n_max), r_out / dp_out
 * unsupported (pass the single-instance entry points for those). */
int dip_guidance_multi(float* dz, float* viol, float* obj, const float* p_full, const dip_instance* table, const int32_t* inst_of,
                       float lam, int B, int L, int n_max, int M_max, dip_stream_t stream);
int dip_round_feasibility_multi(uint8_t* xbits, float* penalty, int64_t* viol_int, int64_t* obj_int, const float* p_full,
                                const dip_instance* table, const int32_t* inst_of, int B, int L, int n_max, dip_stream_t stream);

/* Arithmetic of the attention kernels (the >95 % of the FLOPs):
 *   DIP_PRECISION_TC_SPLIT  (default) tcgen05 tensor cores, every fp32 operand carried as bf16 hi + lo planes,
 *                           products hi.hi + hi.lo + lo.hi accumulated in fp32 (one-step error ~1e-5, bar 1e-3);
 *   DIP_PRECISION_FP32_SIMT exact fp32 FFMA kernels (one-step error ~1e-7) -- the on-GPU checker of the former. */
#define DIP_PRECISION_FP32_SIMT 0
#define DIP_PRECISION_TC_SPLIT 1
int dip_engine_set_precision(dip_engine* e, int mode);

/* Scratch the caller provides (torch-allocated) for a wave of B samples.  Between the calls of one wave the engine keeps wave
 * constants in it (the q|k|v planes of the instance tokens of decoder layer 0, which depend on neither the latent nor the step): a
 * caller that lets anything else write into the workspace between two calls must call dip_engine_reset_cache before the next one
 * (a different workspace address or wave size, and every set_instance / set_instances / set_wave_map / set_decoder, reset it anyway). */
size_t dip_engine_workspace_bytes(const dip_engine* e, int B);
int dip_engine_reset_cache(dip_engine* e);

/* x-hat0 / eps = NoisePredictV2.apply_model(x, t, mip, kpm)  (model/diffusion.py:153-180).
 * out: (B, L, 128) contiguous. */
int dip_engine_denoise(dip_engine* e, float* out, const float* x, int B, float t,
                       void* workspace, size_t workspace_bytes, dip_stream_t stream);
/* p_full (B,L) = decoder probabilities (0 at padding), model/decoder.py:38-50.  y_last (nullable):
 * (B,L,128) copy of the last-L hidden rows (input of linear_proj / select_net, decoder.py:45-52). */
int dip_engine_decode(dip_engine* e, float* p_full, float* y_last, const float* x, int B,
                      void* workspace, size_t workspace_bytes, dip_stream_t stream);
/* grad (B, L, 128) = d loss / d x with loss of model/diffusion.py:623-633 (decoder forward, K2,
 * decoder input-gradient backward).  viol/obj (B) and p_full (B,L) are side outputs (nullable). */
int dip_engine_guidance_grad(dip_engine* e, float* grad, float* p_full, float* viol, float* obj,
                             const float* x, int B, float lam, void* workspace, size_t workspace_bytes,
                             dip_stream_t stream);

/* Per-step coefficients of the DDIM schedule, computed by the host exactly as the reference does
 * (model/diffusion.py:601-604,640-643; see oracle/restate.py:ddim_step_coeffs). */
typedef struct { float t; float sqrt_at, s1m, c_dir, sqrt_aprev, sigma; } dip_ddim_coeffs;

/* One guided DDIM step in place: x <- x_prev  (model/diffusion.py:594-644).  x0_out nullable.
 * noise nullable (sigma == 0 in every reference script). */
int dip_engine_guided_ddim_step(dip_engine* e, float* x, float* x0_out, int B, const dip_ddim_coeffs* co,
                                float gs, float lam, int eps_param, const float* noise,
                                void* workspace, size_t workspace_bytes, dip_stream_t stream);
/* One unguided DDIM step in place (model/diffusion.py:674-699). */
int dip_engine_ddim_step(dip_engine* e, float* x, float* x0_out, int B, const dip_ddim_coeffs* co, int eps_param,
                         const float* noise, void* workspace, size_t workspace_bytes, dip_stream_t stream);

/* The whole S-step guided loop (model/diffusion.py:582-592), in place on x, steps[0..S-1] in
 * execution order, followed by the final decode + round + feasibility of sample.py:162-174.
 * Outputs (device, nullable): xbits (B,n) bytes, penalty (B) fp32, viol_int (B), obj_int (B). */
int dip_engine_sample_guided_ddim(dip_engine* e, float* x, int B, const dip_ddim_coeffs* steps, int S,
                                  float gs, float lam, int eps_param,
                                  uint8_t* xbits, float* penalty, int64_t* viol_int, int64_t* obj_int,
                                  void* workspace, size_t workspace_bytes, dip_stream_t stream);
/* Final decode + round + feasibility only. */
int dip_engine_final_decode(dip_engine* e, const float* x, int B, float* p_full, uint8_t* xbits, float* penalty,
                            int64_t* viol_int, int64_t* obj_int, void* workspace, size_t workspace_bytes,
                            dip_stream_t stream);

/* DDPM guided step (model/diffusion.py:429-454): x <- x_prev.  noise required unless nonzero == 0. */
typedef struct { float t; float c1, c2, std; float nonzero; float sqrt_recip, sqrt_recipm1; } dip_ddpm_coeffs;
int dip_engine_guided_ddpm_step(dip_engine* e, float* x, int B, const dip_ddpm_coeffs* co, float gs, float lam,
                                int eps_param, const float* noise, void* workspace, size_t workspace_bytes,
                                dip_stream_t stream);
int dip_engine_ddpm_step(dip_engine* e, float* x, int B, const dip_ddpm_coeffs* co, int eps_param,
                         const float* noise, void* workspace, size_t workspace_bytes, dip_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* partial-solution hand-off (sample_partialsol.py:172-199) and a residual add                  */
/* ------------------------------------------------------------------------------------------ */
/* For each of B decoded samples (xbits (B,n) bytes 0/1, as written by dip_round_feasibility) choose k = int(n * coverage) variables
 * uniformly without replacement -- the `random.sample(range(n), num_vars)` of sample_partialsol.py:190, drawn here from a Philox
 * counter keyed by (seed, instance_id, sample_base + b) so the choice does not depend on how samples are sharded -- and write the
 * selection mask and the selected values as packed bits, (B, ceil(n/8)) bytes each, numpy.packbits order.  first_sample_values != 0
 * keeps the reference's quirk (values are read from sample 0 for every sample, sample_partialsol.py:179). */
int dip_partial_select(uint8_t* mask_bits, uint8_t* value_bits, const uint8_t* xbits, int B, int n, int k,
                       uint64_t seed, uint64_t instance_id, int sample_base, int first_sample_values, dip_stream_t stream);
/* out = a*x + b*y (y nullable), n % 4 == 0: the extra residual of SolutionEncoder.forward, `x = x + module(x)` (model/encoder.py:35). */
int dip_axpby(float* out, const float* x, const float* y, float a, float b, int64_t n, dip_stream_t stream);

/* Number of kernel launches the library has issued since load (bench.py reports the delta). */
uint64_t dip_launch_count(void);

/* Optional in-library kernel timer: while enabled every launcher records a CUDA-event pair around
 * its launch on the launching stream; dip_prof_collect waits for them and returns the summed
 * milliseconds and launch counts per kernel class (then resets). */
#define DIP_PROF_GEMM 0
#define DIP_PROF_LN 1
#define DIP_PROF_ATTN_FWD 2
#define DIP_PROF_ATTN_BWD_DKDV 3
#define DIP_PROF_ATTN_BWD_DQ 4
#define DIP_PROF_GUIDANCE 5
#define DIP_PROF_UPDATE 6
#define DIP_PROF_HEAD 7
#define DIP_PROF_GNN 8
#define DIP_PROF_OTHER 9
#define DIP_PROF_NCLASSES 10
void dip_prof_enable(int on);
int dip_prof_collect(double* ms_per_class, uint64_t* launches_per_class, int n_classes);

#ifdef __cplusplus
}
#endif
#endif /* DIP_B200_H */
