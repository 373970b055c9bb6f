"""GPU: tensor-core (tcgen05 / TMEM / TMA) kernels against the fp32 SIMT kernels and the fp64 CPU oracle.
Tolerance of the "bf16x2 split" numerics: ~16 mantissa bits per operand -> 1e-4 relative L2 on outputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2

pytestmark = pytest.mark.gpu
D = 128
TC_TOL = 1e-4


def rnd(shape, seed, scale=1.0):
    return torch.from_numpy((scale * np.random.default_rng(seed).standard_normal(shape)).astype(np.float32))


def _attn_ref(qkv, B, T, mask):
    q, k, v = (t.reshape(B, T, D) for t in qkv.double().split(D, dim=-1))
    am = None
    if mask is not None:
        am = torch.zeros((B, 1, T), dtype=torch.float64).masked_fill(mask.view(B, 1, T), float("-inf"))
    return F.scaled_dot_product_attention(q, k, v, attn_mask=am).reshape(B * T, D)


def test_split_bf16(cuda):
    from dip_b200 import ops
    x = rnd((1000, 384), 0, 3.0).to(cuda)
    hi, lo = ops.split_bf16(x)
    rec = hi.float() + lo.float()
    assert float(((rec - x).abs() / x.abs().clamp_min(1e-20)).max()) < 2.0 ** -15
    assert torch.equal(hi, x.to(torch.bfloat16))


@pytest.mark.parametrize("B,T,masked", [(1, 64, False), (2, 100, True), (1, 193, False), (3, 128, True), (2, 257, True), (1, 1000, True)])
def test_attention_fwd_tc(cuda, B, T, masked):
    from dip_b200 import ops
    qkv = rnd((B * T, 3 * D), T, 1.5)
    mask = None
    if masked:
        mask = torch.zeros((B, T), dtype=torch.bool)
        for b in range(B):
            mask[b, T - 5 - 3 * b:] = True
            mask[b, 7 + b] = True
    ref = _attn_ref(qkv, B, T, mask)
    hi, lo = ops.split_bf16(qkv.to(cuda))
    o, lse = ops.attn_fwd_tc(hi, lo, B, T, None if mask is None else mask.to(cuda))
    o_simt, lse_simt = ops.attn_fwd(qkv.to(cuda), B, T, None if mask is None else mask.to(cuda))
    assert rel_l2(o_simt, ref) < 3e-6
    assert rel_l2(o, ref) < TC_TOL, rel_l2(o, ref)
    assert rel_l2(lse, lse_simt) < TC_TOL
    o2, _ = ops.attn_fwd_tc(hi, lo, B, T, None if mask is None else mask.to(cuda))
    assert torch.equal(o, o2)                                  # deterministic


@pytest.mark.parametrize("B", [2, 5])
def test_attention_fwd_tc_full_size(cuda, B):
    """Decoder sequence of the Set-Cover config: T = 4000 keys, against the fp32 SIMT kernel.  B = 5: 160 (sample, query-tile) items on
    148 SMs, so some persistent CTAs walk two items (rings and barrier phases continue across items)."""
    from dip_b200 import ops
    T = 4000
    qkv = rnd((B * T, 3 * D), 5, 1.0).to(cuda)
    hi, lo = ops.split_bf16(qkv)
    o, lse = ops.attn_fwd_tc(hi, lo, B, T)
    o_ref, lse_ref = ops.attn_fwd(qkv, B, T)
    assert rel_l2(o, o_ref) < TC_TOL and rel_l2(lse, lse_ref) < 1e-5


@pytest.mark.parametrize("stream_dq", [True, False])
@pytest.mark.parametrize("B,T,masked", [(1, 128, False), (2, 100, True), (1, 193, False), (2, 257, True), (1, 1000, True)])
def test_attention_bwd_tc(cuda, B, T, masked, stream_dq):
    """stream_dq: dQ as a streaming GEMM over the dS^T tiles the key-stationary kernel stores (default in the engine) / the
    query-stationary kernel that recomputes S and dP."""
    from dip_b200 import ops
    qkv = rnd((B * T, 3 * D), T + 1, 1.5)
    mask = None
    if masked:
        mask = torch.zeros((B, T), dtype=torch.bool)
        for b in range(B):
            mask[b, T - 5 - 3 * b:] = True
            mask[b, 7 + b] = True
    ref_in = qkv.double().requires_grad_(True)
    ref = _attn_ref(ref_in, B, T, mask)
    d_o = rnd((B * T, D), 99)
    ref.backward(d_o.double())
    km = None if mask is None else mask.to(cuda)
    hi, lo = ops.split_bf16(qkv.to(cuda))
    o, lse = ops.attn_fwd_tc(hi, lo, B, T, km)
    dqkv = ops.attn_bwd_tc(hi, lo, o, d_o.to(cuda), lse, B, T, km, stream_dq=stream_dq)
    dq, dk, dv = dqkv.split(D, dim=-1)
    rq, rk, rv = ref_in.grad.split(D, dim=-1)
    assert rel_l2(dv, rv) < TC_TOL, ("dv", rel_l2(dv, rv))
    assert rel_l2(dk, rk) < TC_TOL, ("dk", rel_l2(dk, rk))
    assert rel_l2(dq, rq) < TC_TOL, ("dq", rel_l2(dq, rq))
    assert torch.equal(dqkv, ops.attn_bwd_tc(hi, lo, o, d_o.to(cuda), lse, B, T, km, stream_dq=stream_dq))      # deterministic


def test_attention_bwd_tc_full_size(cuda):
    from dip_b200 import ops
    B, T = 2, 4000
    qkv = rnd((B * T, 3 * D), 5, 1.0).to(cuda)
    d_o = rnd((B * T, D), 6).to(cuda)
    hi, lo = ops.split_bf16(qkv)
    o, lse = ops.attn_fwd_tc(hi, lo, B, T)
    dqkv = ops.attn_bwd_tc(hi, lo, o, d_o, lse, B, T)
    o_ref, lse_ref = ops.attn_fwd(qkv, B, T)
    ref = ops.attn_bwd(qkv, o_ref, d_o, lse_ref, B, T)
    for name, a, r in zip("qkv", dqkv.split(D, dim=-1), ref.split(D, dim=-1)):
        assert rel_l2(a, r) < TC_TOL, (name, rel_l2(a, r))


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (300, 128, 128), (1000, 384, 128), (257, 128, 384), (40000, 128, 128)])
def test_gemm_tc_plain(cuda, M, N, K):
    from dip_b200 import ops
    A, W, bias = rnd((M, K), 0), rnd((N, K), 1, 0.1), rnd((N,), 2)
    ref = A.double() @ W.double().t() + bias.double()
    out = ops.gemm_tc(A.to(cuda), W.to(cuda), bias.to(cuda), want_split=True)
    assert rel_l2(out["C"], ref) < 2e-5, rel_l2(out["C"], ref)
    assert rel_l2(out["hi"].float() + out["lo"].float(), out["C"]) < 1e-5
    out2 = ops.gemm_tc(A.to(cuda), W.to(cuda), bias.to(cuda), act="quickgelu")
    assert rel_l2(out2["C"], ref * torch.sigmoid(1.702 * ref)) < 2e-5
    assert torch.equal(out2["C"], ops.gemm_tc(A.to(cuda), W.to(cuda), bias.to(cuda), act="quickgelu")["C"])


def test_gemm_tc_layernorm_two_source_and_epilogues(cuda):
    from dip_b200 import ops
    B, T, split = 3, 70, 30
    shared, batch = rnd((split, D), 0, 2.0), rnd((B, T - split, D), 1, 2.0)
    gamma, beta = 1 + 0.1 * rnd((D,), 2), 0.1 * rnd((D,), 3)
    W, bias = rnd((3 * D, D), 4, 0.1), rnd((3 * D,), 5)
    x = torch.cat([shared[None].expand(B, -1, -1), batch], dim=1).reshape(B * T, D).double()
    h = F.layer_norm(x, (D,), gamma.double(), beta.double(), 1e-5)
    ref = h @ W.double().t() + bias.double()
    out = ops.gemm_tc(batch.to(cuda), W.to(cuda), bias.to(cuda), a_shared=shared.to(cuda), a_T=T, a_split=split, M=B * T,
                      ln=(gamma.to(cuda), beta.to(cuda)), want_stats=True, want_split=True, want_c=False)
    assert out["C"] is None
    assert rel_l2(out["hi"].float() + out["lo"].float(), ref) < 2e-5
    assert rel_l2(out["mean"], x.mean(-1)) < 1e-5 and rel_l2(out["rstd"], 1.0 / torch.sqrt(x.var(-1, unbiased=False) + 1e-5)) < 1e-5
    # residual (two-source) + preact + mul_quickgelu_grad + row remap/addmat
    W2, b2 = rnd((D, D), 6, 0.1), rnd((D,), 7)
    A2 = rnd((B * T, D), 8)
    pre = torch.empty((B * T, D), device=cuda)
    o = ops.gemm_tc(A2.to(cuda), W2.to(cuda), b2.to(cuda), act="relu", res=batch.to(cuda), res_shared=shared.to(cuda), res_T=T, res_split=split,
                    preact=pre)
    r2 = A2.double() @ W2.double().t() + b2.double() + x
    assert rel_l2(pre, r2) < 2e-5 and rel_l2(o["C"], torch.relu(r2)) < 2e-5
    aux = rnd((B * T, D), 9)
    o3 = ops.gemm_tc(A2.to(cuda), W2.to(cuda), act="mul_quickgelu_grad", aux=aux.to(cuda))
    u = aux.double().requires_grad_(True)
    (u * torch.sigmoid(1.702 * u)).sum().backward()
    assert rel_l2(o3["C"], (A2.double() @ W2.double().t()) * u.grad) < 2e-5
    Lr, T1 = 50, 51
    A4, addmat = rnd((B * Lr, D), 10), rnd((Lr, D), 11)
    out4 = torch.zeros((B * T1, D), device=cuda)
    ops.gemm_tc(A4.to(cuda), W2.to(cuda), act="quickgelu", out=out4, rows_in=Lr, rows_out=T1, row_off=1, addmat=addmat.to(cuda))
    p4 = (A4.double() @ W2.double().t()).view(B, Lr, D) + addmat.double()
    got = out4.view(B, T1, D)
    assert rel_l2(got[:, 1:], p4 * torch.sigmoid(1.702 * p4)) < 2e-5 and float(got[:, 0].abs().max()) == 0.0


@pytest.mark.parametrize("B,T,qb", [(2, 200, 100), (1, 257, 129), (3, 130, 1), (2, 4000, 2000)])
def test_attention_fwd_tc_query_window(cuda, B, T, qb):
    """Only the queries t >= q_begin are computed; they must be bit-identical to the same rows of the full launch when the
    window starts on a 128-row tile boundary of the full launch, and within tolerance otherwise; rows below stay untouched."""
    from dip_b200 import ops
    qkv = rnd((B * T, 3 * D), T + qb, 1.2).to(cuda)
    mask = torch.zeros((B, T), dtype=torch.bool)
    mask[:, T - 3:] = True
    mask[:, 5] = True
    hi, lo = ops.split_bf16(qkv)
    o_full, lse_full = ops.attn_fwd_tc(hi, lo, B, T, mask.to(cuda))
    o, lse = ops.attn_fwd_tc(hi, lo, B, T, mask.to(cuda), q_begin=qb)
    o, lse, o_full, lse_full = o.view(B, T, D), lse.view(B, T), o_full.view(B, T, D), lse_full.view(B, T)
    assert torch.isnan(o[:, :qb]).all() and torch.isnan(lse[:, :qb]).all()             # never written
    assert torch.equal(o[:, qb:], o_full[:, qb:]) and torch.equal(lse[:, qb:], lse_full[:, qb:])   # a query row's result does not depend on its tile


@pytest.mark.parametrize("stream_dq", [True, False])
@pytest.mark.parametrize("B,T,do_b,q_b,k_b", [(2, 200, 100, 100, 0), (2, 200, 0, 100, 100), (1, 257, 129, 129, 129), (3, 130, 1, 7, 3),
                                              (2, 4000, 2000, 2000, 0), (2, 4000, 0, 2000, 2000), (1, 1010, 0, 505, 505)])
def test_attention_bwd_tc_windows(cuda, B, T, do_b, q_b, k_b, stream_dq):
    """Row windows of the backward: dO == 0 below do_begin (and never read: poisoned with NaN here), dq for t >= q_begin,
    dk / dv for t >= k_begin -- against the full launch on the zero-extended dO and against fp64 autograd."""
    from dip_b200 import ops
    qkv = rnd((B * T, 3 * D), T + do_b + 1, 1.2)
    mask = torch.zeros((B, T), dtype=torch.bool)
    mask[:, T - 3:] = True
    mask[:, 5] = True
    d_o = rnd((B, T, D), 17)
    d_o[:, :do_b] = 0
    hi, lo = ops.split_bf16(qkv.to(cuda))
    o, lse = ops.attn_fwd_tc(hi, lo, B, T, mask.to(cuda))
    full = ops.attn_bwd_tc(hi, lo, o, d_o.reshape(B * T, D).to(cuda), lse, B, T, mask.to(cuda), stream_dq=stream_dq).view(B, T, 3 * D)
    # the windowed launch must not read anything below do_begin: poison those rows of dO, o and lse
    d_o_p, o_p, lse_p = d_o.clone().to(cuda), o.clone().view(B, T, D), lse.clone().view(B, T)
    d_o_p[:, :do_b] = float("nan"); o_p[:, :do_b] = float("nan"); lse_p[:, :do_b] = float("nan")
    win = ops.attn_bwd_tc(hi, lo, o_p.reshape(B * T, D), d_o_p.reshape(B * T, D), lse_p.reshape(B * T), B, T, mask.to(cuda),
                          do_begin=do_b, q_begin=q_b, k_begin=k_b, stream_dq=stream_dq).view(B, T, 3 * D)
    assert torch.isnan(win[:, :q_b, :D]).all() and torch.isnan(win[:, :k_b, D:]).all()
    assert not torch.isnan(win[:, q_b:, :D]).any() and not torch.isnan(win[:, k_b:, D:]).any()
    # dq: the same products whatever the window (with a dS workspace and k_begin > 0 the keys below k_begin go through the recompute
    # kernel and are added on top: another summation order); dk / dv sum over fewer (all-zero) query tiles
    assert rel_l2(win[:, q_b:, :D], full[:, q_b:, :D]) < 2e-5
    assert rel_l2(win[:, k_b:, D:], full[:, k_b:, D:]) < 1e-6
    if T <= 300:
        ref_in = qkv.double().requires_grad_(True)
        _attn_ref(ref_in, B, T, mask).backward(d_o.reshape(B * T, D).double())
        g = ref_in.grad.view(B, T, 3 * D)
        assert rel_l2(win[:, q_b:, :D], g[:, q_b:, :D]) < TC_TOL and rel_l2(win[:, k_b:, D:], g[:, k_b:, D:]) < TC_TOL


def test_gemm_tc_row_window(cuda):
    """in_T / in_off: the Linear runs on the tokens t >= w of every sample in place (A rows, LayerNorm statistics, residual,
    outputs all indexed by the full-length row), everything else untouched."""
    from dip_b200 import ops
    B, T, w = 3, 90, 41
    x = rnd((B * T, D), 0, 2.0)
    gamma, beta = 1 + 0.1 * rnd((D,), 2), 0.1 * rnd((D,), 3)
    W, bias = rnd((D, D), 4, 0.1), rnd((D,), 5)
    h = F.layer_norm(x.double(), (D,), gamma.double(), beta.double(), 1e-5)
    ref = (h @ W.double().t() + bias.double() + x.double()).view(B, T, D)
    out = torch.full((B * T, D), 7.0, device=cuda)
    r = ops.gemm_tc(x.to(cuda), W.to(cuda), bias.to(cuda), out=out, M=B * (T - w), rows_in=T - w, rows_out=T, row_off=w, in_T=T, in_off=w,
                    ln=(gamma.to(cuda), beta.to(cuda)), want_stats=True, n_stats=B * T, res=x.to(cuda), res_T=T)
    got = out.view(B, T, D)
    assert rel_l2(got[:, w:], ref[:, w:]) < 2e-5 and float((got[:, :w] - 7.0).abs().max()) == 0.0
    mean = r["mean"].view(B, T)
    assert rel_l2(mean[:, w:], x.view(B, T, D).double().mean(-1)[:, w:]) < 1e-5
    # K = 384 with a window (the in-projection backward of layer 0)
    A3, W3 = rnd((B * T, 3 * D), 6), rnd((D, 3 * D), 7, 0.1)
    out3 = torch.full((B * T, D), -1.0, device=cuda)
    ops.gemm_tc(A3.to(cuda), W3.to(cuda), out=out3, M=B * (T - w), rows_in=T - w, rows_out=T, row_off=w, in_T=T, in_off=w)
    ref3 = (A3.double() @ W3.double().t()).view(B, T, D)
    got3 = out3.view(B, T, D)
    assert rel_l2(got3[:, w:], ref3[:, w:]) < 2e-5 and float((got3[:, :w] + 1.0).abs().max()) == 0.0


@pytest.mark.parametrize("B,T,split,N", [(3, 70, 30, 384), (1, 129, 0, 128), (2, 1000, 500, 384)])
def test_layernorm_split_and_planes_gemm(cuda, B, T, split, N):
    """The in-projection path of the tensor-core mode: LayerNorm written once as bf16 planes, then the TMA-fed Linear."""
    from dip_b200 import ops
    shared = rnd((max(split, 1), D), 0, 2.0)
    batch = rnd((B, T - split, D), 1, 2.0)
    gamma, beta = 1 + 0.1 * rnd((D,), 2), 0.1 * rnd((D,), 3)
    W, bias = rnd((N, D), 4, 0.1), rnd((N,), 5)
    x = (torch.cat([shared[None, :split].expand(B, -1, -1), batch], dim=1) if split else batch).reshape(B * T, D).double()
    h = F.layer_norm(x, (D,), gamma.double(), beta.double(), 1e-5)
    hi, lo, mean, rstd = ops.layernorm_split(batch.to(cuda), gamma.to(cuda), beta.to(cuda), shared=shared.to(cuda) if split else None, T=T, split=split,
                                             rows=B * T, want_stats=True)
    assert rel_l2(hi.float() + lo.float(), h) < 1e-5
    assert rel_l2(mean, x.mean(-1)) < 1e-5 and rel_l2(rstd, 1.0 / torch.sqrt(x.var(-1, unbiased=False) + 1e-5)) < 1e-5
    out = ops.gemm_tc_planes(hi, lo, W.to(cuda), bias.to(cuda), want_split=True)
    ref = h @ W.double().t() + bias.double()
    assert rel_l2(out["C"], ref) < 2e-5, rel_l2(out["C"], ref)
    assert rel_l2(out["hi"].float() + out["lo"].float(), out["C"]) < 1e-5
    assert torch.equal(out["C"], ops.gemm_tc_planes(hi, lo, W.to(cuda), bias.to(cuda))["C"])


def test_layernorm_split_token_window(cuda):
    """t_begin: only the tokens t >= t_begin of every sample are normalised; planes compact, statistics at the input's row numbers."""
    from dip_b200 import ops
    B, T, split, tb = 3, 90, 40, 40
    shared, batch = rnd((split, D), 0, 2.0), rnd((B, T - split, D), 1, 2.0)
    gamma, beta = 1 + 0.1 * rnd((D,), 2), 0.1 * rnd((D,), 3)
    args = dict(shared=shared.to(cuda), T=T, split=split, rows=B * T, want_stats=True)
    hi, lo, mean, rstd = ops.layernorm_split(batch.to(cuda), gamma.to(cuda), beta.to(cuda), **args)
    hw, lw, mw, rw = ops.layernorm_split(batch.to(cuda), gamma.to(cuda), beta.to(cuda), t_begin=tb, **args)
    assert hw.shape == (B * (T - tb), D)
    full = lambda t: t.view(B, T, -1)[:, tb:].reshape(B * (T - tb), -1)
    assert torch.equal(hw, full(hi)) and torch.equal(lw, full(lo))
    assert torch.equal(mw.view(B, T)[:, tb:], mean.view(B, T)[:, tb:]) and torch.equal(rw.view(B, T)[:, tb:], rstd.view(B, T)[:, tb:])


@pytest.mark.parametrize("B,T,prefix,masked", [(2, 600, 300, True), (1, 4000, 2000, False), (3, 257, 129, True), (2, 1500, 1000, True), (6, 4000, 2000, True)])
def test_attention_fwd_tc_key_prefix_hoisting(cuda, B, T, prefix, masked):
    """dip_attn_fwd_tc prefix modes: the queries below floor(prefix/128)*128 resume from the softmax state a "save" launch left after
    the first floor(prefix/64) key tiles; everything must equal the plain launch bit for bit (same operations, same order)."""
    from dip_b200 import ops
    qkv = rnd((B * T, 3 * D), T + prefix, 1.3)
    mask = None
    if masked:
        mask = torch.zeros((B, T), dtype=torch.bool)
        mask[:, prefix - 7:prefix] = True          # padding at the end of the prefix block and of the sequence, like [kpm | kpm]
        mask[:, T - 7:] = True
        mask[:, 3] = True
    hi, lo = ops.split_bf16(qkv.to(cuda))
    km = None if mask is None else mask.to(cuda)
    o0, l0 = ops.attn_fwd_tc(hi, lo, B, T, km)
    o1, l1 = ops.attn_fwd_tc(hi, lo, B, T, km, prefix_keys=prefix)
    assert torch.equal(o0, o1) and torch.equal(l0, l1)


def _block_w(seed):
    g = lambda shape, s, scale: rnd(shape, seed * 100 + s, scale)
    return {"attn.out_proj.weight": g((D, D), 0, 0.1), "attn.out_proj.bias": g((D,), 1, 0.3), "ln_2.weight": 1 + 0.1 * g((D,), 2, 1.0),
            "ln_2.bias": 0.1 * g((D,), 3, 1.0), "mlp.c_fc.weight": g((D, D), 4, 0.15), "mlp.c_fc.bias": g((D,), 5, 0.3),
            "mlp.c_proj.weight": g((D, D), 6, 0.1), "mlp.c_proj.bias": g((D,), 7, 0.3)}


def _chain_ref(attn, xin, w):
    """model/modules.py:293-296 after the attention, fp64."""
    w = {k: v.double() for k, v in w.items()}
    xmid = xin + attn @ w["attn.out_proj.weight"].t() + w["attn.out_proj.bias"]
    u = F.layer_norm(xmid, (D,), w["ln_2.weight"], w["ln_2.bias"], 1e-5) @ w["mlp.c_fc.weight"].t() + w["mlp.c_fc.bias"]
    xout = xmid + (u * torch.sigmoid(1.702 * u)) @ w["mlp.c_proj.weight"].t() + w["mlp.c_proj.bias"]
    return xmid, u, xout


@pytest.mark.parametrize("B,T,win,split", [(1, 128, 0, 0), (2, 100, 0, 0), (3, 257, 100, 0), (2, 1000, 37, 500), (3, 70, 0, 30), (37, 4000, 2000, 2000),
                                           (10, 4000, 0, 0)])
def test_block_chain_tc(cuda, B, T, win, split):
    """chain_tc.cu: the row-wise half of a block in one launch, forward and input-gradient backward, against fp64 torch + autograd;
    windows (rows t < win untouched), two-source residual rows, more tiles than 2 x SMs, bitwise determinism."""
    from dip_b200 import ops
    w = _block_w(T)
    attn = rnd((B * T, D), T + 1, 1.0)
    batch = rnd((B, T - split, D), T + 2, 2.0)
    shared = rnd((max(split, 1), D), T + 3, 2.0)
    xin = (torch.cat([shared[None, :split].expand(B, -1, -1), batch], dim=1) if split else batch).reshape(B * T, D)
    a64 = attn.double().requires_grad_(True)
    x64 = xin.double().requires_grad_(True)
    xmid, u, xout = _chain_ref(a64, x64, w)
    wc = {k: v.to(cuda) for k, v in w.items()}
    fw = ops.block_chain_fwd_tc(attn.to(cuda), batch.to(cuda), wc, B, T, win=win, shared=shared.to(cuda) if split else None, split=split)
    live = lambda t: t.view(B, T, -1)[:, win:]
    dead = lambda t: t.view(B, T, -1)[:, :win]
    for name, ref in (("xmid", xmid), ("u", u), ("xout", xout)):
        assert rel_l2(live(fw[name]), live(ref.detach())) < 2e-5, (name, rel_l2(live(fw[name]), live(ref.detach())))
        assert bool(torch.isnan(dead(fw[name])).all())                   # rows below the window are not written
    xm = xmid.detach()
    assert rel_l2(live(fw["mean2"]), live(xm.mean(-1))) < 1e-5
    assert rel_l2(live(fw["rstd2"]), live(1.0 / torch.sqrt(xm.var(-1, unbiased=False) + 1e-5))) < 1e-5
    fw2 = ops.block_chain_fwd_tc(attn.to(cuda), batch.to(cuda), wc, B, T, win=win, shared=shared.to(cuda) if split else None, split=split)
    assert torch.equal(live(fw["xout"]), live(fw2["xout"])) and torch.equal(live(fw["u"]), live(fw2["u"]))
    # without the optional saves
    fw3 = ops.block_chain_fwd_tc(attn.to(cuda), batch.to(cuda), wc, B, T, win=win, shared=shared.to(cuda) if split else None, split=split, save=False)
    assert torch.equal(live(fw["xout"]), live(fw3["xout"]))
    # backward: gradient of sum(xout * dxout) w.r.t. attn (= d_attn) and the x_mid gradient (= d x_in)
    dxout = rnd((B * T, D), T + 4, 1.0)
    dxout.view(B, T, D)[:, :win] = 0.0
    (xout * dxout.double()).sum().backward()
    nz = lambda t: torch.nan_to_num(t, nan=0.0)
    bw = ops.block_chain_bwd_tc(dxout.to(cuda), nz(fw["u"]), nz(fw["xmid"]), nz(fw["mean2"]), nz(fw["rstd2"]), wc, B, T, win=win)
    assert rel_l2(live(bw["dxmid"]), live(x64.grad)) < 3e-5, rel_l2(live(bw["dxmid"]), live(x64.grad))
    assert rel_l2(live(bw["dattn"]), live(a64.grad)) < 3e-5, rel_l2(live(bw["dattn"]), live(a64.grad))
    assert rel_l2(live(bw["hi"].float() + bw["lo"].float()), live(bw["dattn"])) < 1e-5
    assert bool(torch.isnan(dead(bw["dxmid"])).all()) and bool(torch.isnan(dead(bw["dattn"])).all())
    bw2 = ops.block_chain_bwd_tc(dxout.to(cuda), nz(fw["u"]), nz(fw["xmid"]), nz(fw["mean2"]), nz(fw["rstd2"]), wc, B, T, win=win, planes=False)
    assert torch.equal(live(bw["dattn"]), live(bw2["dattn"])) and torch.equal(live(bw["dxmid"]), live(bw2["dxmid"]))


@pytest.mark.parametrize("B,T,win,split", [(1, 128, 0, 0), (2, 100, 0, 0), (3, 257, 100, 0), (2, 1000, 500, 500), (3, 70, 0, 30), (20, 4000, 2000, 2000)])
def test_block_in_bwd_tc(cuda, B, T, win, split):
    """chain_tc.cu mode 2: dx_in = dres + LN1'(dQKV . W_in) in one launch against fp64 autograd of LN1 + in-projection."""
    from dip_b200 import ops
    w = {"ln_1.weight": 1 + 0.1 * rnd((D,), T + 1), "ln_1.bias": 0.1 * rnd((D,), T + 2), "attn.in_proj_weight": rnd((3 * D, D), T + 3, 0.1)}
    batch = rnd((B, T - split, D), T + 4, 2.0)
    shared = rnd((max(split, 1), D), T + 5, 2.0)
    xin = (torch.cat([shared[None, :split].expand(B, -1, -1), batch], dim=1) if split else batch).reshape(B * T, D)
    dqkv, dres = rnd((B * T, 3 * D), T + 6), rnd((B * T, D), T + 7)
    x64 = xin.double().requires_grad_(True)
    y = F.layer_norm(x64, (D,), w["ln_1.weight"].double(), w["ln_1.bias"].double(), 1e-5) @ w["attn.in_proj_weight"].double().t()
    (y * dqkv.double()).sum().backward()
    ref = x64.grad + dres.double()
    xd = xin.double()
    mean1 = xd.mean(-1).float()
    rstd1 = (1.0 / torch.sqrt(xd.var(-1, unbiased=False) + 1e-5)).float()
    wc = {k: v.to(cuda) for k, v in w.items()}
    got = ops.block_in_bwd_tc(dqkv.to(cuda), batch.to(cuda), mean1.to(cuda), rstd1.to(cuda), wc, dres.to(cuda), B, T, win=win,
                              shared=shared.to(cuda) if split else None, split=split)
    live = lambda t: t.view(B, T, -1)[:, win:]
    assert rel_l2(live(got), live(ref)) < 3e-5, rel_l2(live(got), live(ref))
    assert bool(torch.isnan(got.view(B, T, -1)[:, :win]).all())
    got2 = ops.block_in_bwd_tc(dqkv.to(cuda), batch.to(cuda), mean1.to(cuda), rstd1.to(cuda), wc, dres.to(cuda), B, T, win=win,
                               shared=shared.to(cuda) if split else None, split=split)
    assert torch.equal(live(got), live(got2))
