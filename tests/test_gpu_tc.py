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


def test_attention_fwd_tc_full_size(cuda):
    """Decoder sequence of the Set-Cover config: T = 4000 keys, against the fp32 SIMT kernel."""
    from dip_b200 import ops
    B, T = 2, 4000
    qkv = rnd((B * T, 3 * D), 5, 1.0).to(cuda)
    hi, lo = ops.split_bf16(qkv)
    o, lse = ops.attn_fwd_tc(hi, lo, B, T)
    o_ref, lse_ref = ops.attn_fwd(qkv, B, T)
    assert rel_l2(o, o_ref) < TC_TOL and rel_l2(lse, lse_ref) < 1e-5


@pytest.mark.parametrize("B,T,masked", [(1, 128, False), (2, 100, True), (1, 193, False), (2, 257, True), (1, 1000, True)])
def test_attention_bwd_tc(cuda, B, T, masked):
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
    dqkv = ops.attn_bwd_tc(hi, lo, o, d_o.to(cuda), lse, B, T, km)
    dq, dk, dv = dqkv.split(D, dim=-1)
    rq, rk, rv = ref_in.grad.split(D, dim=-1)
    assert rel_l2(dv, rv) < TC_TOL, ("dv", rel_l2(dv, rv))
    assert rel_l2(dk, rk) < TC_TOL, ("dk", rel_l2(dk, rk))
    assert rel_l2(dq, rq) < TC_TOL, ("dq", rel_l2(dq, rq))
    assert torch.equal(dqkv, ops.attn_bwd_tc(hi, lo, o, d_o.to(cuda), lse, B, T, km))


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
