"""Kernel-level Python wrappers over the C-ABI (one function per exported kernel launcher).

Used by the parity tests (each kernel against the oracle), by the instance encoder and by the
stand-alone module forwards.  Inputs must be contiguous fp32 CUDA tensors; outputs are new tensors.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr, Rows, GemmArgs
from .engine import _stream, require_cuda

D = 128
ACT = {None: 0, "none": 0, "quickgelu": 1, "relu": 2, "sigmoid": 3, "mul_quickgelu_grad": 4}


def _f32(t, name="tensor"):
    if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise _lib.DipError("%s must be a contiguous fp32 CUDA tensor" % name)
    return t


def gemm_nt(A, W, bias=None, act=None, out=None, rows_in=0, rows_out=0, row_off=0, addmat=None, res=None,
            res_shared=None, res_T=0, res_split=0, row_bias_scale=None, preact=None, aux=None, out_rows=None):
    """act(A @ W.T + bias*row_bias_scale + addmat[t] + residual); see dip_gemm_nt in include/dip_b200.h."""
    lib = _lib.load()
    A, W = _f32(A, "A"), _f32(W, "W")
    M, K = A.shape
    N = W.shape[0]
    if W.shape[1] != K:
        raise _lib.DipError("gemm_nt: K mismatch")
    if out is None:
        out = torch.empty((out_rows if out_rows is not None else M, N), dtype=torch.float32, device=A.device)
    g = GemmArgs()
    g.A, g.lda, g.W, g.bias, g.C, g.ldc = ptr(A), K, ptr(W), ptr(bias), ptr(out), N
    g.M, g.N, g.K, g.act = M, N, K, ACT[act]
    g.rows_in, g.rows_out, g.row_off = rows_in, rows_out, row_off
    g.addmat = ptr(addmat)
    g.res = Rows(ptr(res_shared), ptr(res), int(res_T), int(res_split))
    g.row_bias_scale, g.preact, g.aux = ptr(row_bias_scale), ptr(preact), ptr(aux)
    check(lib.dip_gemm_nt(C.byref(g), _stream()))
    return out


def gemm_tc(A, W, bias=None, act=None, out=None, rows_in=0, rows_out=0, row_off=0, addmat=None, res=None, res_shared=None, res_T=0,
            res_split=0, row_bias_scale=None, preact=None, aux=None, out_rows=None, a_shared=None, a_T=0, a_split=0, ln=None,
            want_stats=False, want_split=False, want_c=True, M=None, in_T=0, in_off=0, n_stats=None):
    """tcgen05 version of gemm_nt.  ln = (gamma, beta) fuses LayerNorm into the A load; a_shared/a_T/a_split
    give a two-source A (then A is the per-sample part and M the total row count)."""
    lib = _lib.load()
    A, W = _f32(A, "A"), _f32(W, "W")
    N, K = W.shape
    if M is None:
        M = A.numel() // K
    w_hi, w_lo = split_bf16(W)
    n_out = out_rows if out_rows is not None else M
    dev = A.device
    if out is None and want_c:
        out = torch.empty((n_out, N), dtype=torch.float32, device=dev)
    hi = torch.empty((n_out, N), dtype=torch.bfloat16, device=dev) if want_split else None
    lo = torch.empty((n_out, N), dtype=torch.bfloat16, device=dev) if want_split else None
    n_stats = M if n_stats is None else n_stats
    mean = torch.empty(n_stats, dtype=torch.float32, device=dev) if want_stats else None
    rstd = torch.empty(n_stats, dtype=torch.float32, device=dev) if want_stats else None
    g = GemmArgs()
    g.A, g.lda, g.W, g.bias, g.C, g.ldc = ptr(A), K, None, ptr(bias), ptr(out), N
    g.M, g.N, g.K, g.act = M, N, K, ACT[act]
    g.rows_in, g.rows_out, g.row_off = rows_in, rows_out, row_off
    g.addmat = ptr(addmat)
    g.res = Rows(ptr(res_shared), ptr(res), int(res_T), int(res_split))
    g.row_bias_scale, g.preact, g.aux = ptr(row_bias_scale), ptr(preact), ptr(aux)
    g.split_hi, g.split_lo = ptr(hi), ptr(lo)
    g.in_T, g.in_off = int(in_T), int(in_off)
    ar = Rows(ptr(a_shared), ptr(A), int(a_T), int(a_split)) if a_shared is not None else None
    check(lib.dip_gemm_tc(C.byref(g), ptr(w_hi), ptr(w_lo), C.byref(ar) if ar is not None else None, ptr(ln[0]) if ln else None,
                          ptr(ln[1]) if ln else None, ptr(mean), ptr(rstd), _stream()))
    res_ = {"C": out, "hi": hi, "lo": lo, "mean": mean, "rstd": rstd}
    return res_


def linear_smallk(x, W, bias=None, shift=None, scale=None, act=None):
    lib = _lib.load()
    x, W = _f32(x), _f32(W)
    rows, K = x.shape
    if tuple(W.shape) != (D, K):
        raise _lib.DipError("linear_smallk: weight must be (128, K)")
    y = torch.empty((rows, D), dtype=torch.float32, device=x.device)
    check(lib.dip_linear_smallk(ptr(y), ptr(x), rows, K, ptr(shift), ptr(scale), ptr(W), ptr(bias), ACT[act], _stream()))
    return y


def layernorm_fwd(x, gamma, beta, shared=None, T=0, split=0, save_stats=False, rows=None):
    lib = _lib.load()
    x = _f32(x)
    n_rows = rows if rows is not None else x.numel() // D
    y = torch.empty((n_rows, D), dtype=torch.float32, device=x.device)
    mean = torch.empty(n_rows, dtype=torch.float32, device=x.device) if save_stats else None
    rstd = torch.empty(n_rows, dtype=torch.float32, device=x.device) if save_stats else None
    check(lib.dip_layernorm_fwd(ptr(y), ptr(mean), ptr(rstd), Rows(ptr(shared), ptr(x), int(T), int(split)), ptr(gamma), ptr(beta),
                                n_rows, _stream()))
    return (y, mean, rstd) if save_stats else y


def layernorm_bwd(dy, x, mean, rstd, gamma, dres=None, shared=None, T=0, split=0, t_min=0):
    lib = _lib.load()
    dy = _f32(dy)
    n_rows = dy.numel() // D
    dx = torch.zeros((n_rows, D), dtype=torch.float32, device=dy.device)
    check(lib.dip_layernorm_bwd(ptr(dx), ptr(dy), Rows(ptr(shared), ptr(x), int(T), int(split)), ptr(mean), ptr(rstd), ptr(gamma),
                                ptr(dres), n_rows, int(t_min), _stream()))
    return dx


def attn_fwd(qkv, B, T, key_mask=None, mask_stride=None):
    """qkv: (B*T, 384) packed [q|k|v].  key_mask: (B,T) or (T,) bool/uint8.  Returns (o (B*T,128), lse (B*T))."""
    lib = _lib.load()
    qkv = _f32(qkv)
    o = torch.empty((B * T, D), dtype=torch.float32, device=qkv.device)
    lse = torch.empty(B * T, dtype=torch.float32, device=qkv.device)
    km, ms = _mask(key_mask, B, T, mask_stride)
    base = qkv.data_ptr()
    check(lib.dip_attn_fwd(ptr(o), ptr(lse), base, base + 4 * D, base + 8 * D, 3 * D, ptr(km), ms, B, T, _stream()))
    return o, lse


def attn_bwd(qkv, o, d_o, lse, B, T, key_mask=None, mask_stride=None):
    lib = _lib.load()
    qkv, o, d_o = _f32(qkv), _f32(o), _f32(d_o)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B * T, dtype=torch.float32, device=qkv.device)
    km, ms = _mask(key_mask, B, T, mask_stride)
    base, dbase = qkv.data_ptr(), dqkv.data_ptr()
    check(lib.dip_attn_bwd(dbase, dbase + 4 * D, dbase + 8 * D, 3 * D, base, base + 4 * D, base + 8 * D, 3 * D, ptr(o), ptr(d_o),
                           ptr(lse), ptr(km), ms, ptr(delta), B, T, _stream()))
    return dqkv


def split_bf16(x):
    lib = _lib.load()
    x = _f32(x)
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(lib.dip_split_bf16(ptr(hi), ptr(lo), ptr(x), x.numel(), _stream()))
    return hi, lo


def attn_fwd_tc(qkv_hi, qkv_lo, B, T, key_mask=None, mask_stride=None, q_begin=0, prefix_keys=0):
    """tcgen05 attention forward on the bf16 (hi, lo) planes of the packed (B*T, 384) qkv.  Rows t < q_begin of
    every sample are not computed (returned as NaN so that a test reading them fails).  prefix_keys > 0: key-prefix hoisting --
    a "save" launch over the first prefix_keys keys, then the resuming launch (see dip_attn_fwd_tc)."""
    lib = _lib.load()
    o = torch.full((B * T, D), float("nan"), dtype=torch.float32, device=qkv_hi.device)
    lse = torch.full((B * T,), float("nan"), dtype=torch.float32, device=qkv_hi.device)
    km, ms = _mask(key_mask, B, T, mask_stride)
    if prefix_keys > 0:
        nb = lib.dip_attn_fwd_tc_state_bytes(B, int(prefix_keys))
        state = torch.full((nb // 4,), float("nan"), dtype=torch.float32, device=qkv_hi.device)
        check(lib.dip_attn_fwd_tc(None, None, ptr(qkv_hi), ptr(qkv_lo), ptr(km), ms, B, T, 0, ptr(state), int(prefix_keys), 1, _stream()))
        check(lib.dip_attn_fwd_tc(ptr(o), ptr(lse), ptr(qkv_hi), ptr(qkv_lo), ptr(km), ms, B, T, 0, ptr(state), int(prefix_keys), 2, _stream()))
        return o, lse
    check(lib.dip_attn_fwd_tc(ptr(o), ptr(lse), ptr(qkv_hi), ptr(qkv_lo), ptr(km), ms, B, T, int(q_begin), None, 0, 0, _stream()))
    return o, lse


def attn_bwd_tc(qkv_hi, qkv_lo, o, d_o, lse, B, T, key_mask=None, mask_stride=None, do_begin=0, q_begin=0, k_begin=0, stream_dq=True):
    """Rows outside the windows (dq: t < q_begin; dk, dv: t < k_begin) come back as NaN.  stream_dq: give the launcher a dS workspace
    (dQ as a streaming GEMM over the stored dS^T) instead of running the recompute dQ kernel."""
    lib = _lib.load()
    o, d_o = _f32(o), _f32(d_o)
    do_hi, do_lo = split_bf16(d_o)
    dqkv = torch.full((B * T, 3 * D), float("nan"), dtype=torch.float32, device=o.device)
    delta = torch.empty(B * T, dtype=torch.float32, device=o.device)
    km, ms = _mask(key_mask, B, T, mask_stride)
    dbase = dqkv.data_ptr()
    ds, ds_bytes = None, 0
    if stream_dq:
        ds_bytes = lib.dip_attn_bwd_tc_ds_bytes(B, T, int(do_begin), int(k_begin))
        ds = torch.empty(ds_bytes + 128, dtype=torch.uint8, device=o.device)
    ds_ptr = None if ds is None else ds.data_ptr() + ((-ds.data_ptr()) % 128)
    check(lib.dip_attn_bwd_tc(dbase, dbase + 4 * D, dbase + 8 * D, 3 * D, ptr(qkv_hi), ptr(qkv_lo), ptr(do_hi), ptr(do_lo), ptr(o), ptr(d_o),
                              ptr(lse), ptr(km), ms, ptr(delta), B, T, int(do_begin), int(q_begin), int(k_begin), ds_ptr, ds_bytes, _stream()))
    return dqkv


def _mask(key_mask, B, T, mask_stride):
    if key_mask is None:
        return None, 0
    km = key_mask.to(torch.uint8).contiguous()
    if mask_stride is None:
        mask_stride = T if km.dim() == 2 and km.shape[0] == B and B > 1 else (0 if km.numel() == T else T)
    return km, mask_stride


def ddim_step(x, model_out, coeffs, grad=None, gs=0.0, noise=None, eps_param=False, want_x0=False):
    """coeffs: dict with sqrt_at, s1m, c_dir, sqrt_aprev, sigma.  Out of place; returns x_prev (and x0)."""
    lib = _lib.load()
    x, model_out = _f32(x), _f32(model_out)
    B, L = x.shape[0], x.shape[1]
    xp = torch.empty_like(x)
    x0 = torch.empty_like(x) if want_x0 else None
    check(lib.dip_ddim_step(ptr(xp), ptr(x0), ptr(x), ptr(model_out), L * D, ptr(grad), L * D if grad is not None else 0, ptr(noise),
                            B, L, coeffs["sqrt_at"], coeffs["s1m"], coeffs["c_dir"], coeffs["sqrt_aprev"], coeffs["sigma"], float(gs),
                            int(eps_param), _stream()))
    return (xp, x0) if want_x0 else xp


def ddpm_mean(x, model_out, c1, c2, eps_param=False, sqrt_recip=0.0, sqrt_recipm1=0.0):
    lib = _lib.load()
    x, model_out = _f32(x), _f32(model_out)
    B, L = x.shape[0], x.shape[1]
    mean = torch.empty_like(x)
    check(lib.dip_ddpm_mean(ptr(mean), ptr(x), ptr(model_out), L * D, B, L, float(c1), float(c2), int(eps_param), float(sqrt_recip),
                            float(sqrt_recipm1), _stream()))
    return mean


def ddpm_update(mean, std, grad=None, gs=0.0, noise=None, nonzero=1.0):
    lib = _lib.load()
    mean = _f32(mean)
    B, L = mean.shape[0], mean.shape[1]
    xp = torch.empty_like(mean)
    check(lib.dip_ddpm_update(ptr(xp), ptr(mean), ptr(grad), L * D if grad is not None else 0, ptr(noise), B, L, float(std), float(gs),
                              float(nonzero), _stream()))
    return xp


def time_token(B, T, t, W1, b1, out=None):
    lib = _lib.load()
    if out is None:
        out = torch.zeros((B, T, D), dtype=torch.float32, device=W1.device)
    check(lib.dip_time_token(ptr(out), B, T, float(t), ptr(W1), ptr(b1), _stream()))
    return out


def decoder_head_fwd(y, B, T, L, w, beta, kpm):
    """kpm: (L,) one mask for the wave, or (B, L) one per sample."""
    lib = _lib.load()
    p = torch.empty((B, L), dtype=torch.float32, device=y.device)
    km = kpm.to(torch.uint8).contiguous()
    stride = L if km.dim() == 2 else 0
    check(lib.dip_decoder_head_fwd(ptr(p), ptr(_f32(y)), B, T, L, ptr(w), ptr(beta), ptr(km), stride, _stream()))
    return p


def decoder_head_bwd(dz, B, T, L, w):
    lib = _lib.load()
    dy = torch.empty((B, T, D), dtype=torch.float32, device=dz.device)
    check(lib.dip_decoder_head_bwd(ptr(dy), ptr(_f32(dz)), B, T, L, ptr(w), _stream()))
    return dy


def guidance(p_full, pos_of_col, csr, b, c, lam, M, want_r=False):
    """csr: dict of device int32/fp32 tensors (csr_rowptr, csr_col, csr_val, csc_colptr, csc_row, csc_val, nnz)."""
    lib = _lib.load()
    p_full = _f32(p_full)
    B, L = p_full.shape
    n = pos_of_col.numel()
    dev = p_full.device
    dz = torch.empty((B, L), dtype=torch.float32, device=dev)
    viol = torch.empty(B, dtype=torch.float32, device=dev)
    obj = torch.empty(B, dtype=torch.float32, device=dev)
    r = torch.empty((B, M), dtype=torch.float32, device=dev) if want_r else None
    dp = torch.empty((B, n), dtype=torch.float32, device=dev) if want_r else None
    check(lib.dip_guidance(ptr(dz), ptr(viol), ptr(obj), ptr(r), ptr(dp), ptr(p_full), ptr(pos_of_col), ptr(csr["csr_rowptr"]),
                           ptr(csr["csr_col"]), ptr(csr["csr_val"]), ptr(csr["csc_colptr"]), ptr(csr["csc_row"]), ptr(csr["csc_val"]),
                           ptr(b), ptr(c), float(lam), B, L, n, M, int(csr["nnz"]), _stream()))
    return {"dz": dz, "viol": viol, "obj": obj, "r": r, "dp": dp}


def round_feasibility(p_full, pos_of_col, csr, b, M, csr_val_int=None, b_int=None, c_int=None):
    lib = _lib.load()
    p_full = _f32(p_full)
    B, L = p_full.shape
    n = pos_of_col.numel()
    dev = p_full.device
    xbits = torch.empty((B, n), dtype=torch.uint8, device=dev)
    pen = torch.empty(B, dtype=torch.float32, device=dev)
    vi = torch.empty(B, dtype=torch.int64, device=dev)
    oi = torch.empty(B, dtype=torch.int64, device=dev)
    check(lib.dip_round_feasibility(ptr(xbits), ptr(pen), ptr(vi), ptr(oi), ptr(p_full), ptr(pos_of_col), ptr(csr["csr_rowptr"]),
                                    ptr(csr["csr_col"]), ptr(csr["csr_val"]), ptr(csr_val_int), ptr(b), ptr(b_int), ptr(c_int), B, L, n, M,
                                    int(csr["nnz"]), _stream()))
    return {"xbits": xbits, "penalty": pen, "viol_int": vi, "obj_int": oi}


def conv_reduce(PL, PR, we, s, a_shift, a_scale, seg_ptr, nbr, a):
    lib = _lib.load()
    n_target = PL.shape[0]
    acc = torch.empty((n_target, D), dtype=torch.float32, device=PL.device)
    deg = torch.empty(n_target, dtype=torch.float32, device=PL.device)
    check(lib.dip_conv_reduce(ptr(acc), ptr(deg), ptr(_f32(PL)), ptr(_f32(PR)), ptr(_f32(we)), float(s), float(a_shift), float(a_scale),
                              ptr(seg_ptr), ptr(nbr), ptr(a), n_target, _stream()))
    return acc, deg


def spmm_csr(X, seg_ptr, nbr, val, n_rows, add=None):
    lib = _lib.load()
    y = torch.empty((n_rows, D), dtype=torch.float32, device=X.device)
    check(lib.dip_spmm_csr(ptr(y), ptr(_f32(X)), ptr(add), ptr(seg_ptr), ptr(nbr), ptr(val), n_rows, _stream()))
    return y


def gather_pad(src, idx, L):
    lib = _lib.load()
    idx = idx.to(torch.int64).contiguous()
    out = torch.empty((L, D), dtype=torch.float32, device=src.device)
    check(lib.dip_gather_pad(ptr(out), ptr(_f32(src)), ptr(idx), idx.numel(), L, _stream()))
    return out


def block_forward(x, w, key_mask=None):
    """One ResidualAttentionBlock forward (reference: model/modules.py:293-296) on (B,T,128).
    ``w``: dict with the block's state_dict keys (ln_1.weight, attn.in_proj_weight, ...)."""
    require_cuda()
    B, T, _ = x.shape
    x2 = _f32(x.contiguous()).view(B * T, D)
    h = layernorm_fwd(x2, w["ln_1.weight"], w["ln_1.bias"])
    qkv = gemm_nt(h, w["attn.in_proj_weight"], w["attn.in_proj_bias"])
    o, _ = attn_fwd(qkv, B, T, key_mask)
    xm = gemm_nt(o, w["attn.out_proj.weight"], w["attn.out_proj.bias"], res=x2)
    h = layernorm_fwd(xm, w["ln_2.weight"], w["ln_2.bias"])
    g = gemm_nt(h, w["mlp.c_fc.weight"], w["mlp.c_fc.bias"], act="quickgelu")
    return gemm_nt(g, w["mlp.c_proj.weight"], w["mlp.c_proj.bias"], res=xm).view(B, T, D)


def layernorm_split(x, gamma, beta, shared=None, T=0, split=0, rows=None, want_stats=False, t_begin=0):
    """LayerNorm(x) as bf16 (hi, lo) planes (rows, 128); x optionally two-source like layernorm_fwd.  t_begin > 0: only the tokens
    t >= t_begin of every sample (T rows each), planes compact over them."""
    lib = _lib.load()
    x = _f32(x)
    rows = x.numel() // D if rows is None else rows
    dev = x.device
    live = rows // T * (T - t_begin) if t_begin > 0 else rows
    hi = torch.empty((live, D), dtype=torch.bfloat16, device=dev)
    lo = torch.empty((live, D), dtype=torch.bfloat16, device=dev)
    mean = torch.empty(rows, dtype=torch.float32, device=dev) if want_stats else None
    rstd = torch.empty(rows, dtype=torch.float32, device=dev) if want_stats else None
    r = Rows(ptr(shared), ptr(x), int(T), int(split))
    check(lib.dip_layernorm_split(ptr(hi), ptr(lo), ptr(mean), ptr(rstd), r, ptr(_f32(gamma)), ptr(_f32(beta)), rows, int(t_begin), _stream()))
    return hi, lo, mean, rstd


def gemm_tc_planes(a_hi, a_lo, W, bias=None, want_c=True, want_split=False):
    """tcgen05 Linear whose A operand is given as bf16 (hi, lo) planes (M, 128)."""
    lib = _lib.load()
    W = _f32(W, "W")
    N, K = W.shape
    M = a_hi.shape[0]
    w_hi, w_lo = split_bf16(W)
    dev = a_hi.device
    out = torch.empty((M, N), dtype=torch.float32, device=dev) if want_c else None
    hi = torch.empty((M, N), dtype=torch.bfloat16, device=dev) if want_split else None
    lo = torch.empty((M, N), dtype=torch.bfloat16, device=dev) if want_split else None
    g = GemmArgs()
    g.A, g.lda, g.W, g.bias, g.C, g.ldc = None, K, None, ptr(bias), ptr(out), N
    g.M, g.N, g.K, g.act = M, N, K, ACT[None]
    g.split_hi, g.split_lo = ptr(hi), ptr(lo)
    check(lib.dip_gemm_tc_planes(C.byref(g), ptr(a_hi), ptr(a_lo), ptr(w_hi), ptr(w_lo), _stream()))
    return {"C": out, "hi": hi, "lo": lo}


def axpby(x, y=None, a=1.0, b=1.0):
    """a*x + b*y (y optional) on fp32 tensors of the same shape."""
    lib = _lib.load()
    x = _f32(x)
    out = torch.empty_like(x)
    check(lib.dip_axpby(ptr(out), ptr(x), ptr(_f32(y)) if y is not None else None, float(a), float(b), x.numel(), _stream()))
    return out


def embedding_rows(weight, tokens):
    """weight (V,128) fp32, tokens (L,) int64 -> (L,128): torch.nn.Embedding lookup (model/encoder.py:33) as a row gather."""
    return gather_pad(weight, tokens, tokens.numel())


def partial_select(xbits, k, seed, instance_id, sample_base=0, first_sample_values=True):
    """Device-side partial-solution selection (sample_partialsol.py:172-199): xbits (B,n) uint8 0/1 ->
    (value_bits, mask_bits), each (B, ceil(n/8)) uint8 in numpy.packbits order."""
    lib = _lib.load()
    if xbits.dtype != torch.uint8 or not xbits.is_cuda or not xbits.is_contiguous():
        raise _lib.DipError("xbits must be a contiguous uint8 CUDA tensor")
    B, n = xbits.shape
    nb = (n + 7) // 8
    mask = torch.empty((B, nb), dtype=torch.uint8, device=xbits.device)
    vals = torch.empty((B, nb), dtype=torch.uint8, device=xbits.device)
    check(lib.dip_partial_select(ptr(mask), ptr(vals), ptr(xbits), B, n, int(k), int(seed), int(instance_id), int(sample_base),
                                 1 if first_sample_values else 0, _stream()))
    return vals, mask


def block_chain_fwd_tc(attn, xin, w, B, T, win=0, shared=None, split=0, save=True):
    """dip_block_chain_fwd_tc: attn (B*T,128), xin (B*T,128) or two-source (shared (split,128) + xin (B*(T-split),128)); w: block
    state_dict keys.  Returns dict(xout, xmid, mean2, rstd2, u); rows below `win` are NaN."""
    lib = _lib.load()
    dev = attn.device
    nan = lambda *sh: torch.full(sh, float("nan"), dtype=torch.float32, device=dev)
    out = {"xout": nan(B * T, D), "xmid": nan(B * T, D), "mean2": nan(B * T) if save else None, "rstd2": nan(B * T) if save else None,
           "u": nan(B * T, D) if save else None}
    planes = {k: split_bf16(_f32(w[k])) for k in ("attn.out_proj.weight", "mlp.c_fc.weight", "mlp.c_proj.weight")}
    r = Rows(ptr(shared), ptr(_f32(xin)), int(T), int(split))
    check(lib.dip_block_chain_fwd_tc(ptr(out["xout"]), ptr(out["xmid"]), ptr(out["mean2"]), ptr(out["rstd2"]), ptr(out["u"]), ptr(_f32(attn)), r,
                                     ptr(planes["attn.out_proj.weight"][0]), ptr(planes["attn.out_proj.weight"][1]), ptr(_f32(w["attn.out_proj.bias"])),
                                     ptr(_f32(w["ln_2.weight"])), ptr(_f32(w["ln_2.bias"])), ptr(planes["mlp.c_fc.weight"][0]), ptr(planes["mlp.c_fc.weight"][1]),
                                     ptr(_f32(w["mlp.c_fc.bias"])), ptr(planes["mlp.c_proj.weight"][0]), ptr(planes["mlp.c_proj.weight"][1]),
                                     ptr(_f32(w["mlp.c_proj.bias"])), B, T, int(win), _stream()))
    return out


def block_chain_bwd_tc(dxout, u, xmid, mean2, rstd2, w, B, T, win=0, planes=True):
    """dip_block_chain_bwd_tc.  Returns dict(dxmid, dattn, hi, lo); rows below `win` are NaN / untouched."""
    lib = _lib.load()
    dev = dxout.device
    out = {"dxmid": torch.full((B * T, D), float("nan"), dtype=torch.float32, device=dev),
           "dattn": torch.full((B * T, D), float("nan"), dtype=torch.float32, device=dev),
           "hi": torch.zeros((B * T, D), dtype=torch.bfloat16, device=dev) if planes else None,
           "lo": torch.zeros((B * T, D), dtype=torch.bfloat16, device=dev) if planes else None}
    wt = {k: split_bf16(_f32(w[k]).t().contiguous()) for k in ("mlp.c_proj.weight", "mlp.c_fc.weight", "attn.out_proj.weight")}
    check(lib.dip_block_chain_bwd_tc(ptr(out["dxmid"]), ptr(out["dattn"]), ptr(out["hi"]), ptr(out["lo"]), ptr(_f32(dxout)), ptr(_f32(u)), ptr(_f32(xmid)),
                                     ptr(_f32(mean2)), ptr(_f32(rstd2)), ptr(_f32(w["ln_2.weight"])), ptr(wt["mlp.c_proj.weight"][0]),
                                     ptr(wt["mlp.c_proj.weight"][1]), ptr(wt["mlp.c_fc.weight"][0]), ptr(wt["mlp.c_fc.weight"][1]),
                                     ptr(wt["attn.out_proj.weight"][0]), ptr(wt["attn.out_proj.weight"][1]), B, T, int(win), _stream()))
    return out


def block_in_bwd_tc(dqkv, xin, mean1, rstd1, w, dres, B, T, win=0, shared=None, split=0):
    """dip_block_in_bwd_tc: dx_in = dres + LN1'(dqkv . W_in) on the tokens t >= win.  Rows below `win` are NaN."""
    lib = _lib.load()
    out = torch.full((B * T, D), float("nan"), dtype=torch.float32, device=dqkv.device)
    hi, lo = split_bf16(_f32(w["attn.in_proj_weight"]).t().contiguous())
    r = Rows(ptr(shared), ptr(_f32(xin)), int(T), int(split))
    check(lib.dip_block_in_bwd_tc(ptr(out), ptr(_f32(dqkv)), r, ptr(_f32(mean1)), ptr(_f32(rstd1)), ptr(_f32(w["ln_1.weight"])), ptr(_f32(dres)),
                                  ptr(hi), ptr(lo), B, T, int(win), _stream()))
    return out
