"""GPU: every exported kernel launcher against the CPU oracle (torch fp64 / oracle/restate.py) on
seeded inputs, through the C-ABI (dip_b200.ops is a 1:1 ctypes wrapper)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2, max_abs, small_case, coalesced

pytestmark = pytest.mark.gpu
D = 128


def rnd(shape, seed, scale=1.0):
    return torch.from_numpy((scale * np.random.default_rng(seed).standard_normal(shape)).astype(np.float32))


def test_ddim_step_bit_exact(cuda):
    from dip_b200 import ops
    B, L = 3, 37
    x, x0, g, nz = rnd((B, L, D), 0), rnd((B, L, D), 1), rnd((B, L, D), 2, 1e-4), rnd((B, L, D), 3)
    co = {"sqrt_at": 0.99989998, "s1m": 0.014142, "c_dir": 0.0100001, "sqrt_aprev": 0.99995, "sigma": 0.25}
    full = lambda v: torch.full((B, 1, 1), v, dtype=torch.float32)
    for eps in (False, True):
        for use_g in (False, True):
            for use_n in (False, True):
                if not eps:
                    e = (x - x0 * full(co["sqrt_at"])) / full(co["s1m"]); p0 = x0
                else:
                    e = x0; p0 = (x - full(co["s1m"]) * e) / full(co["sqrt_at"])
                if use_g:
                    e = e - 1e5 * g
                ref = full(co["sqrt_aprev"]) * p0 + full(co["c_dir"]) * e
                if use_n:
                    ref = ref + full(co["sigma"]) * nz
                out, o0 = ops.ddim_step(x.to(cuda), x0.to(cuda), co, grad=g.to(cuda) if use_g else None, gs=1e5,
                                        noise=nz.to(cuda) if use_n else None, eps_param=eps, want_x0=True)
                assert torch.equal(out.cpu(), ref), (eps, use_g, use_n)
                assert torch.equal(o0.cpu(), p0)


def test_ddpm_mean_update_bit_exact(cuda):
    from dip_b200 import ops
    B, L = 2, 19
    x, out, g, nz = rnd((B, L, D), 0), rnd((B, L, D), 1), rnd((B, L, D), 2, 1e-4), rnd((B, L, D), 3)
    c1, c2, std, gs = 0.3123, 0.6871, 0.0456, 1.5e4
    mean = ops.ddpm_mean(x.to(cuda), out.to(cuda), c1, c2)
    ref_mean = torch.tensor(c1) * out + torch.tensor(c2) * x
    assert torch.equal(mean.cpu(), ref_mean)
    xp = ops.ddpm_update(mean, std, grad=g.to(cuda), gs=gs, noise=nz.to(cuda), nonzero=1.0)
    ref = (ref_mean - gs * torch.tensor(std) * g) + 1.0 * torch.tensor(std) * nz
    assert torch.equal(xp.cpu(), ref)
    m2 = ops.ddpm_mean(x.to(cuda), out.to(cuda), c1, c2, eps_param=True, sqrt_recip=1.7, sqrt_recipm1=1.3)
    assert torch.equal(m2.cpu(), torch.tensor(c1) * (torch.tensor(1.7) * x - torch.tensor(1.3) * out) + torch.tensor(c2) * x)


@pytest.mark.parametrize("M,N,K", [(300, 128, 128), (257, 384, 128), (130, 128, 384), (64, 128, 256)])
def test_gemm_plain_and_activations(cuda, M, N, K):
    from dip_b200 import ops
    A, W, bias = rnd((M, K), 0), rnd((N, K), 1, 0.1), rnd((N,), 2)
    ref = A.double() @ W.double().t() + bias.double()
    for act, fn in (("none", lambda v: v), ("relu", torch.relu), ("sigmoid", torch.sigmoid),
                    ("quickgelu", lambda v: v * torch.sigmoid(1.702 * v))):
        out = ops.gemm_nt(A.to(cuda), W.to(cuda), bias.to(cuda), act=act)
        assert rel_l2(out, fn(ref)) < 2e-6, act


def test_gemm_epilogues(cuda):
    from dip_b200 import ops
    B, Lr, T = 3, 50, 51
    A, W = rnd((B * Lr, D), 0), rnd((D, D), 1, 0.1)
    addmat = rnd((Lr, D), 2)
    # row remap + addmat (denoiser first layer): rows b*Lr+t -> b*T+1+t
    out = torch.zeros((B * T, D), device=cuda)
    ops.gemm_nt(A.to(cuda), W.to(cuda), act="quickgelu", out=out, rows_in=Lr, rows_out=T, row_off=1, addmat=addmat.to(cuda))
    pre = (A.double() @ W.double().t()).view(B, Lr, D) + addmat.double()
    ref = pre * torch.sigmoid(1.702 * pre)
    got = out.view(B, T, D)
    assert rel_l2(got[:, 1:], ref) < 2e-6 and float(got[:, 0].abs().max()) == 0.0
    # two-source residual (decoder layer 0: shared mip rows + per-sample x rows), preact and bias row scale
    T2, split = 40, 25
    A2, bias = rnd((B * T2, D), 3), rnd((D,), 4)
    shared, batch = rnd((split, D), 5), rnd((B, T2 - split, D), 6)
    rbs = torch.arange(B * T2, dtype=torch.float32) % 7
    pre_buf = torch.empty((B * T2, D), device=cuda)
    out2 = ops.gemm_nt(A2.to(cuda), W.to(cuda), bias.to(cuda), act="relu", res=batch.to(cuda), res_shared=shared.to(cuda), res_T=T2,
                       res_split=split, row_bias_scale=rbs.to(cuda), preact=pre_buf)
    resid = torch.cat([shared[None].expand(B, -1, -1), batch], dim=1).reshape(B * T2, D)
    ref2 = A2.double() @ W.double().t() + bias.double() * rbs.double()[:, None] + resid.double()
    assert rel_l2(pre_buf, ref2) < 2e-6 and rel_l2(out2, torch.relu(ref2)) < 2e-6
    # multiply by QuickGELU'(aux)
    aux = rnd((B * T2, D), 7)
    out3 = ops.gemm_nt(A2.to(cuda), W.to(cuda), act="mul_quickgelu_grad", aux=aux.to(cuda))
    u = aux.double().requires_grad_(True)
    (u * torch.sigmoid(1.702 * u)).sum().backward()
    assert rel_l2(out3, (A2.double() @ W.double().t()) * u.grad) < 2e-6


def test_gemm_rejects_bad_shapes(lib):
    from dip_b200 import _lib
    g = _lib.GemmArgs()
    g.A, g.W, g.C = 256, 256, 256
    g.M, g.N, g.K, g.lda, g.ldc = 4, 100, 128, 128, 100
    assert lib.dip_gemm_nt(g, None) != 0 and b"multiple of 128" in lib.dip_last_error()


def test_layernorm_fwd_bwd(cuda):
    from dip_b200 import ops
    B, T, split = 2, 33, 13
    shared, batch = rnd((split, D), 0, 2.0), rnd((B, T - split, D), 1, 2.0)
    gamma, beta = 1 + 0.1 * rnd((D,), 2), 0.1 * rnd((D,), 3)
    x = torch.cat([shared[None].expand(B, -1, -1), batch], dim=1).reshape(B * T, D).double().requires_grad_(True)
    y = F.layer_norm(x, (D,), gamma.double(), beta.double(), 1e-5)
    out, mean, rstd = ops.layernorm_fwd(batch.to(cuda), gamma.to(cuda), beta.to(cuda), shared=shared.to(cuda), T=T, split=split,
                                        save_stats=True, rows=B * T)
    assert rel_l2(out, y) < 2e-6
    dy, dres = rnd((B * T, D), 4), rnd((B * T, D), 5)
    y.backward(dy.double())
    dx = ops.layernorm_bwd(dy.to(cuda), batch.to(cuda), mean, rstd, gamma.to(cuda), dres=dres.to(cuda), shared=shared.to(cuda), T=T, split=split)
    assert rel_l2(dx, x.grad + dres.double()) < 5e-6
    # t_min: rows with t < t_min are left untouched (zeros from the wrapper's allocation)
    dx2 = ops.layernorm_bwd(dy.to(cuda), batch.to(cuda), mean, rstd, gamma.to(cuda), dres=dres.to(cuda), shared=shared.to(cuda), T=T, split=split,
                            t_min=split).view(B, T, D)
    assert float(dx2[:, :split].abs().max()) == 0.0 and torch.equal(dx2[:, split:], dx.view(B, T, D)[:, split:])


def _attn_ref(qkv, B, T, mask):
    q, k, v = (t.reshape(B, T, D) for t in qkv.double().split(D, dim=-1))
    am = None
    if mask is not None:
        am = torch.zeros((B, 1, T), dtype=torch.float64).masked_fill(mask.view(B, 1, T) if mask.dim() == 2 else mask.view(1, 1, T), float("-inf"))
    return F.scaled_dot_product_attention(q, k, v, attn_mask=am).reshape(B * T, D)


@pytest.mark.parametrize("B,T,masked", [(2, 100, True), (1, 193, False), (3, 64, True), (2, 257, True)])
def test_attention_fwd_bwd(cuda, B, T, masked):
    from dip_b200 import ops
    qkv = rnd((B * T, 3 * D), T, 1.5)
    mask = None
    if masked:
        mask = torch.zeros((B, T), dtype=torch.bool)
        for b in range(B):
            mask[b, T - 5 - 3 * b:] = True
            mask[b, 7 + b] = True
    ref_in = qkv.double().requires_grad_(True)
    ref = _attn_ref(ref_in, B, T, mask)
    o, lse = ops.attn_fwd(qkv.to(cuda), B, T, None if mask is None else mask.to(cuda))
    assert rel_l2(o, ref) < 3e-6
    d_o = rnd((B * T, D), 99)
    ref.backward(d_o.double())
    dqkv = ops.attn_bwd(qkv.to(cuda), o, d_o.to(cuda), lse, B, T, None if mask is None else mask.to(cuda))
    assert rel_l2(dqkv, ref_in.grad) < 1e-5
    # shared mask (stride 0) == the same mask repeated
    if masked:
        o2, _ = ops.attn_fwd(qkv.to(cuda), B, T, mask[0].to(cuda))
        o3, _ = ops.attn_fwd(qkv.to(cuda), B, T, mask[0][None].repeat(B, 1).contiguous().to(cuda), mask_stride=T)
        assert torch.equal(o2, o3)


def test_time_token(cuda):
    from dip_b200 import ops
    from oracle import restate
    W1, b1 = rnd((D, 2 * D), 0, 0.06), rnd((D,), 1, 0.06)
    for t in (0, 7, 99, 999):
        emb = restate.timestep_embedding(torch.tensor([t]), 2 * D)
        ref = restate.quick_gelu(emb.double() @ W1.double().t() + b1.double())[0]
        out = ops.time_token(2, 5, t, W1.to(cuda), b1.to(cuda))
        # t * f_k in fp32 amplifies a 1-ulp difference of expf (GPU vs host libm) by t: 999 * 6e-8 ~ 6e-5 rad
        assert rel_l2(out[0, 0], ref) < (5e-6 if t < 100 else 5e-5) and torch.equal(out[0, 0], out[1, 0]) and float(out[:, 1:].abs().max()) == 0.0


def test_decoder_head(cuda):
    from dip_b200 import ops
    B, T, L = 2, 30, 12
    y, w, beta = rnd((B, T, D), 0), rnd((D,), 1, 0.1), rnd((1,), 2)
    kpm = torch.zeros(L, dtype=torch.bool); kpm[-3:] = True
    p = ops.decoder_head_fwd(y.to(cuda), B, T, L, w.to(cuda), beta.to(cuda), kpm.to(cuda))
    ref = torch.sigmoid(y[:, T - L:].double() @ w.double() + beta.double()).masked_fill(kpm[None], 0.0)
    assert rel_l2(p, ref) < 2e-6 and float(p[:, -3:].abs().max()) == 0.0
    dz = rnd((B, L), 3)
    dy = ops.decoder_head_bwd(dz.to(cuda), B, T, L, w.to(cuda))
    assert float(dy[:, :T - L].abs().max()) == 0.0
    assert rel_l2(dy[:, T - L:], dz[:, :, None] * w[None, None, :]) < 1e-7


def _csr_dev(cs, cuda):
    from dip_b200.engine import coo_to_csr_csc
    csr = coo_to_csr_csc(cs.rows, cs.cols, cs.vals, cs.M, cs.n)
    return {k: (torch.from_numpy(v).to(cuda) if isinstance(v, np.ndarray) else v) for k, v in csr.items()}


def test_guidance_vs_oracle(cuda):
    from dip_b200 import ops
    from oracle import restate
    cs = small_case()
    rows, cols, vals, M = coalesced(cs)
    B, L, n = 4, cs.L, cs.n
    p_full = torch.zeros((B, L))
    p_full[:, :n] = torch.from_numpy(np.random.default_rng(0).random((B, n)).astype(np.float32))
    p_full[3, :n] = 0.0            # row 0 has b == 0: exact tie r == 0 -> h = 1/2
    pos = torch.arange(n, dtype=torch.int32)
    b, c = torch.from_numpy(cs.b), torch.from_numpy(cs.c)
    lam = 0.9
    out = ops.guidance(p_full.to(cuda), pos.to(cuda), _csr_dev(cs, cuda), b.to(cuda), c.to(cuda), lam, M, want_r=True)
    p64 = p_full[:, :n].double()
    dp, r, h = restate.guidance_grad_p(p64, rows, cols, vals.double(), M, b.double(), c.double(), float(np.float32(lam)))
    assert rel_l2(out["r"], r) < 2e-6
    assert float(h[3, 0]) == 0.5
    assert rel_l2(out["dp"], dp) < 2e-6
    dz_ref = torch.zeros((B, L), dtype=torch.float64)
    dz_ref[:, :n] = dp * p64 * (1 - p64)
    assert rel_l2(out["dz"], dz_ref) < 2e-6 and float(out["dz"][:, n:].abs().max()) == 0.0
    assert rel_l2(out["viol"], torch.clamp(r, min=0).sum(-1)) < 2e-6
    assert rel_l2(out["obj"], p64 @ c.double()) < 2e-6


def test_round_feasibility_vs_oracle(cuda):
    from dip_b200 import ops, synth
    from dip_b200.engine import coo_to_csr_csc
    from oracle import restate
    inst = synth.set_cover(40, 90, 0.08, seed=3)
    B, L, n = 6, 96, 90
    rng = np.random.default_rng(1)
    p = rng.random((B, n)).astype(np.float32)
    p[0] = 1.0; p[1] = 0.0; p[2, :] = 0.5; p[3] = (p[3] > 0.2)
    p_full = np.zeros((B, L), np.float32); p_full[:, :n] = p
    csr = coo_to_csr_csc(inst.rows, inst.cols, inst.a_val, inst.M, n)
    csri = coo_to_csr_csc(inst.rows, inst.cols, inst.a_int.astype(np.float32), inst.M, n)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    csr_d = {k: (dev(v) if isinstance(v, np.ndarray) else v) for k, v in csr.items()}
    out = ops.round_feasibility(dev(p_full), torch.arange(n, dtype=torch.int32, device=cuda), csr_d, dev(inst.b), inst.M,
                                csr_val_int=dev(csri["csr_val"].astype(np.int32)), b_int=dev(inst.b_int), c_int=dev(inst.c_int))
    xb = torch.round(torch.from_numpy(p)).numpy().astype(np.int64)        # half-to-even like torch.round
    assert np.array_equal(out["xbits"].cpu().numpy(), xb)
    feas, obj, viol = restate.feasibility_int(xb, inst.rows, inst.cols, inst.a_int, inst.M, inst.b_int, inst.c_int)
    assert np.array_equal(out["viol_int"].cpu().numpy(), viol) and np.array_equal(out["obj_int"].cpu().numpy(), obj)
    assert feas[0] and not feas[1]
    r, c_, v, _ = (torch.from_numpy(inst.rows), torch.from_numpy(inst.cols), torch.from_numpy(inst.a_val), None)
    pen = torch.clamp(restate.spmv_rows(r, c_, v, inst.M, torch.from_numpy(xb).float()) - torch.from_numpy(inst.b), min=0).sum(-1)
    # every row term max((Ax)_i - b_i, 0) is exact for integer-valued A; only the order of the row sum differs
    assert torch.allclose(out["penalty"].cpu(), pen, rtol=1e-6, atol=0)
    assert np.array_equal(out["penalty"].cpu().numpy() == 0, pen.numpy() == 0)
    assert np.array_equal((out["penalty"].cpu().numpy() <= 1e-5), feas)


def test_randn(cuda):
    from dip_b200.engine import randn
    a = randn((1 << 20,), seed=7, subseq=3)
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.var()) - 1) < 1e-2 and torch.isfinite(a).all()
    assert torch.equal(a, randn((1 << 20,), seed=7, subseq=3))
    assert not torch.equal(a, randn((1 << 20,), seed=7, subseq=4))
    assert not torch.equal(a, randn((1 << 20,), seed=8, subseq=3))
    odd = randn((1001,), seed=7, subseq=3)                                 # prefix property, ragged tail
    assert torch.equal(odd, a[:1001])
    kurt = float(((a - a.mean()) ** 4).mean() / a.var() ** 2)
    assert abs(kurt - 3.0) < 0.05


def test_gnn_kernels(cuda):
    from dip_b200 import ops
    nt, ns, E = 9, 13, 40
    rng = np.random.default_rng(0)
    ptr = np.zeros(nt + 1, np.int32)
    tgt = np.sort(rng.integers(0, nt, E))
    for t in tgt:
        ptr[t + 1] += 1
    ptr = np.cumsum(ptr).astype(np.int32)
    nbr = rng.integers(0, ns, E).astype(np.int32)
    a = rng.standard_normal(E).astype(np.float32)
    PL, PR, X, add, we = rnd((nt, D), 1), rnd((ns, D), 2), rnd((ns, D), 3), rnd((nt, D), 4), rnd((D,), 5)
    dev = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(cuda) if isinstance(v, np.ndarray) else v.to(cuda)
    acc, deg = ops.conv_reduce(dev(PL), dev(PR), dev(we), 0.7, 0.1, 1.3, dev(ptr), dev(nbr), dev(a))
    y = ops.spmm_csr(dev(X), dev(ptr), dev(nbr), dev(a), nt, add=dev(add))
    ref_acc = torch.zeros((nt, D), dtype=torch.float64); ref_y = add.double().clone()
    for e in range(E):
        t = int(tgt[e])
        ref_acc[t] += torch.relu(0.7 * (PL[t].double() + we.double() * ((float(a[e]) + np.float32(0.1)) * np.float32(1.3)) + PR[nbr[e]].double()))
        ref_y[t] += float(a[e]) * X[nbr[e]].double()
    assert rel_l2(acc, ref_acc) < 2e-6 and rel_l2(y, ref_y) < 2e-6
    assert np.array_equal(deg.cpu().numpy(), np.diff(ptr).astype(np.float32))
    src = rnd((20, D), 6)
    idx = torch.tensor([3, 0, 19, 7])
    g = ops.gather_pad(src.to(cuda), idx.to(cuda), 6)
    assert torch.equal(g[:4].cpu(), src[idx]) and float(g[4:].abs().max()) == 0.0
    xs = rnd((11, 5), 7); W = rnd((D, 5), 8); bb = rnd((D,), 9); sh = rnd((5,), 10); sc = 1 + 0.1 * rnd((5,), 11)
    yk = ops.linear_smallk(xs.to(cuda), W.to(cuda), bb.to(cuda), sh.to(cuda), sc.to(cuda), act="relu")
    assert rel_l2(yk, torch.relu(((xs.double() + sh.double()) * sc.double()) @ W.double().t() + bb.double())) < 2e-6


def test_round_feasibility_toy_lp_known_answers(cuda):
    """Known-answer test from the reference's own fixture toy_example.lp (via tests/golden/toy_is15.npz): all 2^15
    assignments through dip_round_feasibility -> 1118 feasible, optimum 8, 4 optimal."""
    import itertools
    from dip_b200 import ops
    from dip_b200.engine import coo_to_csr_csc
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "toy_is15.npz"))
    n, M = len(d["c"]), len(d["b"])
    X = np.array(list(itertools.product((0, 1), repeat=n)), np.float32)
    L = 16
    # probabilities on either side of 1/2 instead of exact bits, so that the rounding is exercised as well
    p_full = np.zeros((X.shape[0], L), np.float32)
    p_full[:, :n] = np.where(X > 0, 0.5 + 1e-3, 0.5 - 1e-3)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    csr = coo_to_csr_csc(d["rows"], d["cols"], d["vals"].astype(np.float32), M, n)
    csr_d = {k: (dev(v) if isinstance(v, np.ndarray) else v) for k, v in csr.items()}
    out = ops.round_feasibility(dev(p_full), torch.arange(n, dtype=torch.int32, device=cuda), csr_d, dev(d["b"].astype(np.float32)), M,
                                csr_val_int=dev(csr["csr_val"].astype(np.int32)), b_int=dev(d["b"]), c_int=dev(d["c"]))
    assert np.array_equal(out["xbits"].cpu().numpy(), X.astype(np.uint8))
    viol, obj = out["viol_int"].cpu().numpy(), out["obj_int"].cpu().numpy()
    feas = viol == 0
    assert np.array_equal(feas, out["penalty"].cpu().numpy() <= 1e-5)
    assert feas.sum() == 1118 and obj[feas].max() == 8 and (feas & (obj == 8)).sum() == 4
    assert viol.sum() == int(d["viol_checksum"]) and (obj * feas).sum() == int(d["obj_checksum"])
