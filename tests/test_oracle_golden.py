"""CPU: the restatement (oracle/restate.py) against the golden vectors produced by the UNMODIFIED
reference (oracle/gen_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import torch

from dip_b200 import synth
from oracle import restate
from helpers import golden, tw, rel_l2, case_tensors, coalesced, small_case, small_graph

TOL = 2e-5   # fp32 restatement vs fp32 reference: different but equivalent op order (observed ~1e-6)


def _setup(param_seed=1):
    cs = small_case()
    Wd = tw(synth.denoiser_weights(1, seed=param_seed))
    Wdec = tw(synth.decoder_weights(2, seed=param_seed + 1))
    return cs, Wd, Wdec, case_tensors(cs), coalesced(cs)


def test_fixture_inputs_are_reproducible():
    g = golden("ddim_guided_x0")
    cs = small_case()
    assert abs(float(np.abs(cs.x).sum()) - float(g["x_checksum"])) < 1e-3
    w = synth.decoder_weights(2, seed=2)
    assert abs(float(sum(np.abs(v).sum() for v in w.values())) - float(g["w_checksum"])) < 1e-3


def test_ddim_guided_teacher_forced_x0():
    g = golden("ddim_guided_x0")
    cs, Wd, Wdec, t, A = _setup()
    S = int(g["S"])
    tab = restate.ddim_tables(restate.trainer_buffers()["alphas_cumprod"], S)
    for index in g["indices"]:
        index = int(index)
        out = restate.guided_ddim_step(Wd, Wdec, t["x"], S - 1 - index, restate.ddim_step_coeffs(tab, index), t["mipB"], t["kpmB"],
                                       A, t["b"], t["c"], float(g["gs"]), float(g["lam"]))
        assert rel_l2(out["x0"], g["tf%d_x0" % index]) < TOL
        assert rel_l2(out["p"], g["tf%d_p" % index]) < TOL
        assert rel_l2(out["grad"], g["tf%d_grad" % index]) < 1e-4
        assert rel_l2(out["x_prev"], g["tf%d_xprev" % index]) < 1e-4
        # padded positions receive exactly zero gradient
        assert float(out["grad"][:, cs.n:].abs().max()) == 0.0


def test_ddim_guided_free_running_and_final_decode():
    g = golden("ddim_guided_x0")
    cs, Wd, Wdec, t, A = _setup()
    S = int(g["S"])
    tab = restate.ddim_tables(restate.trainer_buffers()["alphas_cumprod"], S)
    x = t["x"]
    for k in range(3):
        out = restate.guided_ddim_step(Wd, Wdec, x, S - 1 - k, restate.ddim_step_coeffs(tab, k), t["mipB"], t["kpmB"], A, t["b"], t["c"],
                                       float(g["gs"]), float(g["lam"]))
        x = out["x_prev"]
        assert rel_l2(x, g["free%d_x" % k]) < 5e-4 * (k + 1)
    # decode the REFERENCE's latent so decode parity is not polluted by trajectory drift
    fd = restate.final_decode(Wdec, t["mipB"], torch.from_numpy(g["free2_x"]), t["kpmB"], A, t["b"])
    assert rel_l2(fd["p"], g["final_p"]) < TOL
    assert np.array_equal(fd["x_int"].numpy(), g["final_x"])
    assert np.allclose(fd["penalty"].numpy(), g["final_penalty"], rtol=1e-6, atol=1e-6)
    un = restate.unguided_ddim_step(Wd, t["x"], S - 1, restate.ddim_step_coeffs(tab, 0), t["mipB"], t["kpmB"])
    assert rel_l2(un["x_prev"], g["unguided_xprev"]) < TOL


def test_ddim_guided_eps_parameterisation():
    g = golden("ddim_guided_eps")
    cs, Wd, Wdec, t, A = _setup()
    S = int(g["S"])
    tab = restate.ddim_tables(restate.trainer_buffers()["alphas_cumprod"], S)
    for index in g["indices"]:
        index = int(index)
        out = restate.guided_ddim_step(Wd, Wdec, t["x"], S - 1 - index, restate.ddim_step_coeffs(tab, index), t["mipB"], t["kpmB"],
                                       A, t["b"], t["c"], float(g["gs"]), float(g["lam"]), parameterization="eps")
        assert rel_l2(out["x0"], g["tf%d_x0" % index]) < 1e-4
        assert rel_l2(out["x_prev"], g["tf%d_xprev" % index]) < 1e-4


def test_ddpm_guided_and_unguided():
    g = golden("ddpm_guided_x0")
    cs = small_case(seed=1)
    Wd, Wdec = tw(synth.denoiser_weights(1, seed=1)), tw(synth.decoder_weights(2, seed=2))
    t, A = case_tensors(cs), coalesced(cs)
    buf = restate.trainer_buffers()
    tabs = restate.ddpm_tables(buf)
    for ti in g["ts"]:
        ti = int(ti)
        out = restate.guided_ddpm_step(Wd, Wdec, t["x"], ti, tabs, buf, t["mipB"], t["kpmB"], A, t["b"], t["c"], float(g["gs"]),
                                       float(g["lam"]), torch.from_numpy(g["t%d_noise" % ti]))
        assert rel_l2(out["mean"], g["t%d_mean" % ti]) < TOL
        assert rel_l2(out["grad"], g["t%d_grad" % ti]) < 1e-4
        assert rel_l2(out["x_prev"], g["t%d_xprev" % ti]) < 1e-4
        un = restate.unguided_ddpm_step(Wd, t["x"], ti, tabs, buf, t["mipB"], t["kpmB"], torch.from_numpy(g["t%d_unoise" % ti]))
        assert rel_l2(un["x_prev"], g["t%d_unguided" % ti]) < TOL


def test_encoder():
    g = golden("encoder_small")
    gr = small_graph()
    W = tw(synth.encoder_weights(seed=3))
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    v = restate.encoder_fwd(W, T(gr.cons), T(gr.edge_index), T(gr.edge_attr), T(gr.var))
    assert rel_l2(v, g["var_features"]) < TOL
    mip, kpm = restate.get_features(W, T(gr.cons), T(gr.edge_index), T(gr.edge_attr), T(gr.var), T(gr.int_indices), 24)
    assert rel_l2(mip, g["mip_features"]) < TOL
    assert np.array_equal(kpm.numpy(), g["kpm"])


def test_guidance_closed_form_matches_autograd():
    cs, Wd, Wdec, t, A = _setup()
    rows, cols, vals, M = A
    p = torch.rand(cs.B, cs.n, dtype=torch.float64, requires_grad=True)
    b, c = t["b"].double(), t["c"].double()
    loss, r, _, _ = restate.guidance_loss(p, rows, cols, vals.double(), M, b, c, 0.7)
    loss.backward()
    dp, _, h = restate.guidance_grad_p(p.detach(), rows, cols, vals.double(), M, b, c, 0.7)
    assert torch.allclose(p.grad, dp, atol=1e-12)
    # exact tie: hinge gradient is split 1/2 (torch.maximum semantics the kernels replicate)
    p0 = torch.zeros(1, cs.n, dtype=torch.float64)
    _, r0, h0 = restate.guidance_grad_p(p0, rows, cols, vals.double(), M, b, c, 0.0)
    assert float(h0[0, 0]) == 0.5 and float(r0[0, 0]) == 0.0       # b[0] == 0 in the fixture


def _toy():
    import itertools
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "toy_is15.npz"))
    X = np.array(list(itertools.product((0, 1), repeat=len(d["c"]))), np.int64)
    return d, X


def test_toy_lp_known_answers_both_oracles():
    """The reference's toy_example.lp (IndependentSet-15) brute-forced: 1118 feasible of 32768 assignments, optimum 8
    reached by 4 of them.  Both integer oracles (numpy and the C restatement) must reproduce that."""
    import ctypes as C
    import subprocess
    from oracle import restate
    d, X = _toy()
    assert (int(d["n_feasible"]), int(d["best_obj"]), int(d["n_optimal"])) == (1118, 8, 4)
    M = len(d["b"])
    feas, obj, viol = restate.feasibility_int(X, d["rows"], d["cols"], d["vals"], M, d["b"], d["c"])
    assert feas.sum() == 1118 and obj[feas].max() == 8 and (feas & (obj == 8)).sum() == 4
    assert viol.sum() == int(d["viol_checksum"]) and (obj * feas).sum() == int(d["obj_checksum"])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "oracle", "_build", "libfeas_int.so")
    if not os.path.isfile(so):
        subprocess.run(["make", "-C", os.path.join(root, "oracle")], check=True)
    lib = C.CDLL(so)
    B, n = X.shape
    p = X.astype(np.float32)
    xb = np.zeros((B, n), np.uint8); v2 = np.zeros(B, np.int64); o2 = np.zeros(B, np.int64); scratch = np.zeros(M, np.int64)
    ptr = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
    rows, cols, vals, b, c = (np.ascontiguousarray(d[k]) for k in ("rows", "cols", "vals", "b", "c"))
    lib.oracle_round_feasibility(C.c_int64(B), C.c_int64(n), C.c_int64(M), C.c_int64(len(vals)), ptr(rows), ptr(cols), ptr(vals), ptr(b),
                                 ptr(c), ptr(p), ptr(xb), ptr(v2), ptr(o2), ptr(scratch))
    assert np.array_equal(v2, viol) and np.array_equal(o2, obj) and np.array_equal(xb, X)
