"""GPU parity tests proper: the engine (C-ABI, libdip_b200.so) against the golden vectors produced
by the unmodified reference and against the CPU oracle on the same seeded inputs.

Protocol: P1 teacher-forced single steps at several schedule positions,
P2 free-running first steps, P3 decode / round / feasibility bit-exact for identical latents.
Tolerances: fp32 one-step relative L2 <= 1e-4 on latents/gradients (north_star bar: 1e-3; the
reference's own fp32-vs-fp64 noise floor is ~2e-7, the kernels' differing summation order ~1e-6);
0/1 assignments, feasibility flags and integer objectives exact.
"""
import numpy as np
import pytest
import torch

from dip_b200 import synth
from helpers import golden, tw, rel_l2, case_tensors, coalesced, small_case, small_graph, sc_case

pytestmark = pytest.mark.gpu
# one-step relative-L2 tolerance per precision mode: exact fp32 SIMT kernels, and the tcgen05 "bf16x2 split"
# mode whose bar is north_star's 1e-3 (observed ~1e-5)
TOLS = {"fp32": 1e-4, "tc": 1e-3}
MODES = ["fp32", "tc"]


def make_engine(cuda, cs, n_den=1, n_dec=2, wseed=1, ints=False, precision="fp32"):
    from dip_b200.engine import Engine
    eng = Engine()
    eng.set_precision(precision)
    eng.set_denoiser(tw(synth.denoiser_weights(n_den, seed=wseed), cuda))
    eng.set_decoder(tw(synth.decoder_weights(n_dec, seed=wseed + 1), cuda))
    eng.set_instance(torch.from_numpy(cs.mip).to(cuda), torch.from_numpy(cs.kpm).to(cuda), cs.rows, cs.cols, cs.vals, cs.M, cs.b, cs.c)
    return eng


def coeffs(S, index):
    from dip_b200.engine import Engine
    from oracle import restate
    tab = restate.ddim_tables(restate.trainer_buffers()["alphas_cumprod"], S)
    c = restate.ddim_step_coeffs(tab, index)
    return Engine.ddim_coeffs(S - 1 - index, c["sqrt_at"], c["s1m"], c["c_dir"], c["sqrt_aprev"], c["sigma"])


@pytest.mark.parametrize("mode", MODES)
def test_p1_teacher_forced_ddim_x0(cuda, mode):
    TOL = TOLS[mode]
    g = golden("ddim_guided_x0")
    cs = small_case()
    eng = make_engine(cuda, cs, precision=mode)
    S, gs, lam = int(g["S"]), float(g["gs"]), float(g["lam"])
    x_in = torch.from_numpy(cs.x).to(cuda)
    for index in g["indices"]:
        index = int(index)
        x0 = eng.denoise(x_in, S - 1 - index)
        assert rel_l2(x0, g["tf%d_x0" % index]) < TOL
        grad, p_full, viol, obj = eng.guidance_grad(x_in, lam)
        assert rel_l2(p_full[:, :cs.n], g["tf%d_p" % index]) < TOL
        assert rel_l2(grad, g["tf%d_grad" % index]) < TOL
        assert float(grad[:, cs.n:].abs().max()) == 0.0          # padded latents get exactly zero gradient
        x = x_in.clone()
        x0_out = torch.empty_like(x)
        eng.guided_ddim_step(x, coeffs(S, index), gs, lam, x0_out=x0_out)
        assert rel_l2(x, g["tf%d_xprev" % index]) < TOL
        assert torch.equal(x0_out, x0)


@pytest.mark.parametrize("mode", MODES)
def test_p2_free_running_and_p3_final_decode(cuda, mode):
    TOL = TOLS[mode]
    g = golden("ddim_guided_x0")
    cs = small_case()
    eng = make_engine(cuda, cs, precision=mode)
    S, gs, lam = int(g["S"]), float(g["gs"]), float(g["lam"])
    x = torch.from_numpy(cs.x).to(cuda).clone()
    for k in range(3):
        eng.guided_ddim_step(x, coeffs(S, k), gs, lam)
        assert rel_l2(x, g["free%d_x" % k]) < 10 * TOL * (k + 1)
    # P3: identical latents in -> bit-identical assignment / feasibility out
    fd = eng.final_decode(torch.from_numpy(g["free2_x"]).to(cuda))
    assert rel_l2(fd["p_full"][:, :cs.n], g["final_p"]) < TOL
    assert np.array_equal(fd["xbits"].cpu().numpy().astype(np.float32), g["final_x"])
    assert np.allclose(fd["penalty"].cpu().numpy(), g["final_penalty"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(fd["penalty"].cpu().numpy() <= 1e-5, g["final_penalty"] <= 1e-5)
    margin = float((torch.from_numpy(g["final_p"]) - 0.5).abs().min())
    assert margin > 1e-4, "fixture decode margin too small for a meaningful bit-exact check"
    # unguided step
    xu = torch.from_numpy(cs.x).to(cuda).clone()
    eng.ddim_step(xu, coeffs(S, 0))
    assert rel_l2(xu, g["unguided_xprev"]) < TOL


@pytest.mark.parametrize("mode", MODES)
def test_eps_parameterisation(cuda, mode):
    TOL = TOLS[mode]
    g = golden("ddim_guided_eps")
    cs = small_case()
    eng = make_engine(cuda, cs, precision=mode)
    S, gs, lam = int(g["S"]), float(g["gs"]), float(g["lam"])
    for index in g["indices"]:
        index = int(index)
        x = torch.from_numpy(cs.x).to(cuda).clone()
        x0 = torch.empty_like(x)
        eng.guided_ddim_step(x, coeffs(S, index), gs, lam, eps_param=True, x0_out=x0)
        assert rel_l2(x0, g["tf%d_x0" % index]) < TOL
        assert rel_l2(x, g["tf%d_xprev" % index]) < TOL


@pytest.mark.parametrize("mode", MODES)
def test_ddpm_steps(cuda, mode):
    TOL = TOLS[mode]
    from dip_b200 import _lib
    from oracle import restate
    g = golden("ddpm_guided_x0")
    cs = small_case(seed=1)
    eng = make_engine(cuda, cs, precision=mode)
    buf = restate.trainer_buffers()
    tabs = restate.ddpm_tables(buf)
    for ti in g["ts"]:
        ti = int(ti)
        std = float((0.5 * tabs["posterior_log_variance_clipped"][ti]).exp())
        co = _lib.DdpmCoeffs(float(ti), float(tabs["posterior_mean_coef1"][ti]), float(tabs["posterior_mean_coef2"][ti]), std,
                             0.0 if ti == 0 else 1.0, float(buf["sqrt_recip_alphas_cumprod"][ti]), float(buf["sqrt_recipm1_alphas_cumprod"][ti]))
        x = torch.from_numpy(cs.x).to(cuda).clone()
        eng.guided_ddpm_step(x, co, float(g["gs"]), float(g["lam"]), noise=torch.from_numpy(g["t%d_noise" % ti]).to(cuda))
        assert rel_l2(x, g["t%d_xprev" % ti]) < TOL
        xu = torch.from_numpy(cs.x).to(cuda).clone()
        eng.ddpm_step(xu, co, noise=torch.from_numpy(g["t%d_unoise" % ti]).to(cuda))
        assert rel_l2(xu, g["t%d_unguided" % ti]) < TOL


def test_encoder_vs_golden(cuda):
    TOL = 1e-4
    from dip_b200 import encoder, ops
    g = golden("encoder_small")
    gr = small_graph()
    W = tw(synth.encoder_weights(seed=3), cuda)
    w = {k[len("mip_encoder.gnn_model."):]: v for k, v in W.items()}
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    v = encoder.gnn_forward(w, T(gr.cons), T(gr.edge_index), T(gr.edge_attr), T(gr.var))
    assert rel_l2(v, g["var_features"]) < TOL
    mip = ops.gather_pad(v, T(gr.int_indices), 24)
    assert rel_l2(mip, g["mip_features"][0]) < TOL
    assert float(mip[len(gr.int_indices):].abs().max()) == 0.0


@pytest.mark.parametrize("mode", MODES)
def test_one_call_sampler_equals_stepwise_and_is_deterministic(cuda, mode):
    TOL = TOLS[mode]
    cs = small_case()
    eng = make_engine(cuda, cs, precision=mode)
    S = 5
    steps = [coeffs(S, k) for k in range(S)]
    x1 = torch.from_numpy(cs.x).to(cuda).clone()
    out = eng.sample_guided_ddim(x1, steps, 1e5, 0.9)
    x2 = torch.from_numpy(cs.x).to(cuda).clone()
    for co in steps:
        eng.guided_ddim_step(x2, co, 1e5, 0.9)
    fd = eng.final_decode(x2)
    assert torch.equal(x1, x2) and torch.equal(out["xbits"], fd["xbits"]) and torch.equal(out["penalty"], fd["penalty"])
    x3 = torch.from_numpy(cs.x).to(cuda).clone()
    out3 = eng.sample_guided_ddim(x3, steps, 1e5, 0.9)
    assert torch.equal(x1, x3) and torch.equal(out["xbits"], out3["xbits"])        # run-to-run bitwise


@pytest.mark.parametrize("mode", MODES)
def test_batch_invariance(cuda, mode):
    TOL = TOLS[mode]
    """A sample's trajectory must not depend on which wave it is in (sharding invariance)."""
    cs = small_case(B=4)
    eng = make_engine(cuda, cs, precision=mode)
    S = 4            # S = 3 would index alpha-bar[1000]
    steps = [coeffs(S, k) for k in range(S)]
    xa = torch.from_numpy(cs.x).to(cuda).clone()
    eng.sample_guided_ddim(xa, steps, 1e5, 0.9)
    for b in (0, 3):
        xb = torch.from_numpy(cs.x[b:b + 1]).to(cuda).clone()
        eng.sample_guided_ddim(xb, steps, 1e5, 0.9)
        assert torch.equal(xa[b:b + 1], xb)


def test_engine_errors(cuda):
    from dip_b200 import _lib
    from dip_b200.engine import Engine
    eng = Engine()
    with pytest.raises(_lib.DipError):
        eng.workspace(2)                                     # no instance yet
    cs = small_case()
    eng.set_conditions(torch.from_numpy(cs.mip).to(cuda), torch.from_numpy(cs.kpm).to(cuda))
    with pytest.raises(_lib.DipError, match="set_denoiser"):
        eng.denoise(torch.from_numpy(cs.x).to(cuda), 3)
    with pytest.raises(_lib.DipError):
        eng._x(torch.from_numpy(cs.x))                       # CPU latent: no CPU path
    with pytest.raises(_lib.DipError, match="objective vector"):
        eng.set_instance(torch.from_numpy(cs.mip).to(cuda), torch.from_numpy(cs.kpm).to(cuda), cs.rows, cs.cols, cs.vals, cs.M, cs.b,
                         np.zeros(cs.n + 1, np.float32))


@pytest.mark.parametrize("mode", MODES)
def test_full_size_set_cover_step_vs_oracle(cuda, mode):
    TOL = TOLS[mode]
    """BASELINE config shapes (L = n = 2000, M = 1000, nnz = 1e5, D = 128, n_d = 1, n_e = 2): one guided
    step of 2 samples against the CPU oracle (a few seconds on the host)."""
    from dip_b200.engine import Engine
    from oracle import restate
    inst, mip, kpm, x = sc_case(B=2)
    Wd, Wdec = synth.denoiser_weights(1, seed=1), synth.decoder_weights(2, seed=2)
    eng = Engine()
    eng.set_precision(mode)
    eng.set_denoiser(tw(Wd, cuda)); eng.set_decoder(tw(Wdec, cuda))
    eng.set_instance(torch.from_numpy(mip).to(cuda), torch.from_numpy(kpm).to(cuda), inst.rows, inst.cols, inst.a_val, inst.M, inst.b, inst.c,
                     a_int=inst.a_int, b_int=inst.b_int, c_int=inst.c_int)
    S, gs, lam = 100, 1e5, 0.9
    tab = restate.ddim_tables(restate.trainer_buffers()["alphas_cumprod"], S)
    A = (torch.from_numpy(inst.rows), torch.from_numpy(inst.cols), torch.from_numpy(inst.a_val), inst.M)
    mipB, kpmB = torch.from_numpy(mip)[None].repeat(2, 1, 1), torch.from_numpy(kpm)[None].repeat(2, 1)
    ref = restate.guided_ddim_step(tw(Wd), tw(Wdec), torch.from_numpy(x), S - 1, restate.ddim_step_coeffs(tab, 0), mipB, kpmB, A,
                                   torch.from_numpy(inst.b), torch.from_numpy(inst.c), gs, lam)
    xg = torch.from_numpy(x).to(cuda).clone()
    x0 = torch.empty_like(xg)
    grad, p_full, viol, obj = eng.guidance_grad(xg, lam)
    assert rel_l2(p_full, ref["p_full"]) < TOL
    assert rel_l2(grad, ref["grad"]) < TOL
    eng.guided_ddim_step(xg, coeffs(S, 0), gs, lam, x0_out=x0)
    assert rel_l2(x0, ref["x0"]) < TOL
    assert rel_l2(xg, ref["x_prev"]) < TOL
    # P3 at full size: decode the oracle's latent on both sides
    fd = eng.final_decode(ref["x_prev"].to(cuda))
    ofd = restate.final_decode(tw(Wdec), mipB, ref["x_prev"], kpmB, A, torch.from_numpy(inst.b))
    margin = (ofd["p"] - 0.5).abs()
    safe = margin > 1e-4
    assert np.array_equal(fd["xbits"].cpu().numpy()[safe.numpy()], ofd["x_int"].numpy().astype(np.uint8)[safe.numpy()])
    feas, objv, violv = restate.feasibility_int(fd["xbits"].cpu().numpy(), inst.rows, inst.cols, inst.a_int, inst.M, inst.b_int, inst.c_int)
    assert np.array_equal(fd["viol_int"].cpu().numpy(), violv) and np.array_equal(fd["obj_int"].cpu().numpy(), objv)
    assert np.array_equal(fd["penalty"].cpu().numpy() <= 1e-5, feas)


@pytest.mark.parametrize("mode", MODES)
def test_independent_set_config_step_vs_oracle(cuda, mode):
    """configs[2] shapes scaled to a seconds-long oracle run: Independent Set (2 non-zeros per edge row + singleton rows, M > n),
    n < L so a third of the tokens are padding, TWO denoiser layers like the shipped IS checkpoint, gs = 2e4, lambda = 0.5."""
    from dip_b200.engine import Engine
    from oracle import restate
    TOL = TOLS[mode]
    n, L, B = 600, 900, 3
    inst = synth.independent_set(n_nodes=n, n_edges=1900, n_single=n, seed=5)
    mip = synth.mip_features(L, n, seed=15)
    kpm = np.zeros(L, dtype=bool); kpm[n:] = True
    x = synth.initial_noise(5, range(B), L)
    Wd, Wdec = synth.denoiser_weights(2, seed=3), synth.decoder_weights(2, seed=4)
    eng = Engine()
    eng.set_precision(mode)
    eng.set_denoiser(tw(Wd, cuda)); eng.set_decoder(tw(Wdec, cuda))
    eng.set_instance(torch.from_numpy(mip).to(cuda), torch.from_numpy(kpm).to(cuda), inst.rows, inst.cols, inst.a_val, inst.M, inst.b, inst.c,
                     a_int=inst.a_int, b_int=inst.b_int, c_int=inst.c_int)
    S, gs, lam = 100, 2e4, 0.5
    tab = restate.ddim_tables(restate.trainer_buffers()["alphas_cumprod"], S)
    A = (torch.from_numpy(inst.rows), torch.from_numpy(inst.cols), torch.from_numpy(inst.a_val), inst.M)
    mipB, kpmB = torch.from_numpy(mip)[None].repeat(B, 1, 1), torch.from_numpy(kpm)[None].repeat(B, 1)
    for index in (0, 57):
        ref = restate.guided_ddim_step(tw(Wd), tw(Wdec), torch.from_numpy(x), S - 1 - index, restate.ddim_step_coeffs(tab, index), mipB, kpmB, A,
                                       torch.from_numpy(inst.b), torch.from_numpy(inst.c), gs, lam)
        xg = torch.from_numpy(x).to(cuda).clone()
        x0 = torch.empty_like(xg)
        grad, p_full, viol, obj = eng.guidance_grad(xg, lam)
        assert rel_l2(p_full, ref["p_full"]) < TOL and float(p_full[:, n:].abs().max()) == 0.0
        assert rel_l2(grad[:, :n], ref["grad"][:, :n]) < TOL
        assert float(grad[:, n:].abs().max()) == 0.0 and float(ref["grad"][:, n:].abs().max()) == 0.0      # padding tokens get no gradient
        eng.guided_ddim_step(xg, coeffs(S, index), gs, lam, x0_out=x0)
        assert rel_l2(x0[:, :n], ref["x0"][:, :n]) < TOL and rel_l2(xg[:, :n], ref["x_prev"][:, :n]) < TOL
    fd = eng.final_decode(ref["x_prev"].to(cuda))
    ofd = restate.final_decode(tw(Wdec), mipB, ref["x_prev"], kpmB, A, torch.from_numpy(inst.b))
    safe = ((ofd["p"] - 0.5).abs() > 1e-4).numpy()
    assert np.array_equal(fd["xbits"].cpu().numpy()[safe], ofd["x_int"].numpy().astype(np.uint8)[safe])
    feas, objv, violv = restate.feasibility_int(fd["xbits"].cpu().numpy(), inst.rows, inst.cols, inst.a_int, inst.M, inst.b_int, inst.c_int)
    assert np.array_equal(fd["viol_int"].cpu().numpy(), violv) and np.array_equal(fd["obj_int"].cpu().numpy(), objv)


def test_facility_location_length_tc_vs_fp32(cuda):
    """configs[3]'s longest sequence (CF: L = 5050, decoder sequence 10100; too slow for the CPU oracle inside the test budget): the
    tensor-core path against the fp32 SIMT path of the same engine -- which the other tests tie to the oracle -- on one guided step."""
    from dip_b200.engine import Engine
    inst, mip, kpm, x = sc_case(L=5050, n=5050, M=600, density=0.004, B=1, seed=9)
    Wd, Wdec = synth.denoiser_weights(1, seed=1), synth.decoder_weights(2, seed=2)
    out = {}
    for mode in MODES:
        eng = Engine()
        eng.set_precision(mode)
        eng.set_denoiser(tw(Wd, cuda)); eng.set_decoder(tw(Wdec, cuda))
        eng.set_instance(torch.from_numpy(mip).to(cuda), torch.from_numpy(kpm).to(cuda), inst.rows, inst.cols, inst.a_val, inst.M, inst.b, inst.c)
        xg = torch.from_numpy(x).to(cuda).clone()
        grad, p_full, _, _ = eng.guidance_grad(xg, 0.7)
        eng.guided_ddim_step(xg, coeffs(100, 0), 1e3, 0.7)
        out[mode] = (grad.clone(), p_full.clone(), xg)
    for a, r in zip(out["tc"], out["fp32"]):
        assert rel_l2(a, r) < 1e-3, rel_l2(a, r)


def test_toy_instance_end_to_end_known_answers(cuda):
    """The reference's toy instance (15 binaries, 25 rows) through the whole path -- lpio.make_instance -> engine -> S guided DDIM
    steps -> decode -> round -> integer check -- with every returned flag / objective looked up in the brute-forced table of all
    2^15 assignments (1118 feasible, optimum 8)."""
    import itertools
    import os
    from dip_b200 import lpio, schedule
    from dip_b200.engine import Engine
    from oracle import restate
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "toy_is15.npz"))
    inst = lpio.make_instance([str(s) for s in d["names"]], d["rows"], d["cols"], d["vals"], d["b"], (-d["c"]).astype(np.int32), sense=-1)
    n, L, B, S = inst.n, 16, 64, 10
    X = np.array(list(itertools.product((0, 1), repeat=n)), np.int64)
    feas_tab, obj_tab, viol_tab = restate.feasibility_int(X, inst.rows, inst.cols, inst.a_int, inst.M, inst.b_int, inst.c_int)
    assert feas_tab.sum() == 1118 and obj_tab[feas_tab].min() == -8
    mip = synth.mip_features(L, n, seed=21)
    kpm = np.zeros(L, dtype=bool); kpm[n:] = True
    eng = Engine()
    eng.set_denoiser(tw(synth.denoiser_weights(1, seed=1), cuda)); eng.set_decoder(tw(synth.decoder_weights(2, seed=2), cuda))
    eng.set_instance(torch.from_numpy(mip).to(cuda), torch.from_numpy(kpm).to(cuda), inst.rows, inst.cols, inst.a_val, inst.M, inst.b, inst.c,
                     a_int=inst.a_int, b_int=inst.b_int, c_int=inst.c_int)
    x = torch.from_numpy(synth.initial_noise(3, range(B), L)).to(cuda)
    out = eng.sample_guided_ddim(x, schedule.ddim_coeffs(S), 2e4, 0.5)
    bits = out["xbits"].cpu().numpy().astype(np.int64)
    idx = (bits * (1 << np.arange(n - 1, -1, -1))).sum(1)               # row of X holding this assignment
    assert np.array_equal(X[idx], bits)
    assert np.array_equal(out["viol_int"].cpu().numpy(), viol_tab[idx]) and np.array_equal(out["obj_int"].cpu().numpy(), obj_tab[idx])
    assert np.array_equal(out["penalty"].cpu().numpy() <= 1e-5, feas_tab[idx])


@pytest.mark.parametrize("L,n,M,B", [(130, 130, 9, 1), (130, 97, 40, 5), (65, 64, 3, 41), (200, 1, 2, 2)])
def test_ragged_shapes_and_degenerate_rows(cuda, L, n, M, B):
    """Edge cases around the kernels' tile sizes and the sparse passes: L not a multiple of 32/64/128, no padding and heavy padding,
    a single integer variable, an empty constraint row, an all-zero objective, duplicate COO entries (summed like coalesce()), a wave
    larger than the SM count / cluster size -- one guided step against the oracle, tensor-core mode."""
    from dip_b200.engine import Engine
    from oracle import restate
    cs = small_case(seed=L + n, L=L, n=n, M=M, nnz=max(4, 3 * M), B=B, dup=2)
    cs.rows[cs.rows == M - 1] = 0                      # row M-1 is empty
    cs.c[:] = 0.0                                      # objective without any weight
    eng = Engine()
    Wd, Wdec = synth.denoiser_weights(1, seed=1), synth.decoder_weights(2, seed=2)
    eng.set_denoiser(tw(Wd, cuda)); eng.set_decoder(tw(Wdec, cuda))
    eng.set_instance(torch.from_numpy(cs.mip).to(cuda), torch.from_numpy(cs.kpm).to(cuda), cs.rows, cs.cols, cs.vals, cs.M, cs.b, cs.c)
    S, gs, lam = 100, 1e3, 0.3
    tab = restate.ddim_tables(restate.trainer_buffers()["alphas_cumprod"], S)
    t = case_tensors(cs)
    ref = restate.guided_ddim_step(tw(Wd), tw(Wdec), t["x"], S - 1, restate.ddim_step_coeffs(tab, 0), t["mipB"], t["kpmB"], coalesced(cs), t["b"], t["c"], gs, lam)
    xg = t["x"].to(cuda).clone()
    grad, p_full, viol, obj = eng.guidance_grad(xg, lam)
    assert rel_l2(p_full, ref["p_full"]) < 1e-3 and rel_l2(grad[:, :n], ref["grad"][:, :n]) < 1e-3
    assert float(obj.abs().max()) == 0.0
    eng.guided_ddim_step(xg, coeffs(S, 0), gs, lam)
    assert rel_l2(xg[:, :n], ref["x_prev"][:, :n]) < 1e-3


@pytest.mark.parametrize("mode", MODES)
def test_waves_mixing_instances_equal_single_instance_waves(cuda, mode):
    """dip_engine_set_instances + dip_engine_set_wave_map: a wave whose samples belong to three different instances (different n, M,
    nnz, masks, conditions; same padding_len) must give every sample bit-for-bit what it gets in a wave of its own instance -- guided
    steps, the one-call sampler, rounding, feasibility and integer objective (sample_partialsol.py:55: 15 samples per instance)."""
    from dip_b200 import schedule
    from dip_b200.engine import Engine
    L, S = 40, 6
    cases = [small_case(seed=50 + i, L=L, n=n, M=M, nnz=nnz, B=1, dup=0) for i, (n, M, nnz) in enumerate([(40, 9, 30), (33, 14, 50), (25, 5, 12)])]
    for cs in cases:                                    # the exact-integer twin needs a duplicate-free COO
        _, first = np.unique(cs.rows * 1000 + cs.cols, return_index=True)
        first.sort()
        cs.rows, cs.cols, cs.vals = cs.rows[first], cs.cols[first], cs.vals[first]
    eng = Engine()
    eng.set_precision(mode)
    eng.set_denoiser(tw(synth.denoiser_weights(1, seed=1), cuda)); eng.set_decoder(tw(synth.decoder_weights(2, seed=2), cuda))

    def prep(cs):
        ai = np.rint(cs.vals).astype(np.int32)
        return eng.prepare_instance(torch.from_numpy(cs.mip).to(cuda), torch.from_numpy(cs.kpm).to(cuda), cs.rows, cs.cols, cs.vals, cs.M, cs.b, cs.c,
                                    a_int=ai, b_int=np.floor(cs.b).astype(np.int32), c_int=np.rint(100 * cs.c).astype(np.int32))

    preps = [prep(cs) for cs in cases]
    inst_of = [0, 1, 2, 1, 0, 2, 2, 1]
    B = len(inst_of)
    x0 = torch.from_numpy(synth.initial_noise(77, range(B), L)).to(cuda)
    steps = schedule.ddim_coeffs(100)[:S]
    gs, lam = 2e4, 0.5
    # mixed wave
    info = eng.use_instances(preps, inst_of)
    assert info["n"] == 40 and info["n_inst"] == 3
    xm = x0.clone()
    for k in range(2):
        eng.guided_ddim_step(xm, steps[k], gs, lam)
    mixed_step = xm.clone()
    mixed = eng.sample_guided_ddim(x0.clone(), steps, gs, lam)
    assert mixed["xbits"].shape == (B, 40)
    with pytest.raises(_lib_err()):
        eng.guided_ddim_step(x0[:3].clone(), steps[0], gs, lam)            # the map was set for 8 samples
    # every instance alone
    for i, (cs, pr) in enumerate(zip(cases, preps)):
        idx = [s for s, k in enumerate(inst_of) if k == i]
        eng.use_instance(pr)
        xs = x0[idx].clone()
        for k in range(2):
            eng.guided_ddim_step(xs, steps[k], gs, lam)
        assert torch.equal(xs, mixed_step[idx]), i
        alone = eng.sample_guided_ddim(x0[idx].clone(), steps, gs, lam)
        assert torch.equal(alone["x"], mixed["x"][idx]), i
        assert torch.equal(alone["xbits"], mixed["xbits"][idx][:, :cs.n]) and int(mixed["xbits"][idx][:, cs.n:].sum()) == 0
        for k in ("penalty", "viol_int", "obj_int"):
            assert torch.equal(alone[k], mixed[k][idx]), (i, k)


def _lib_err():
    from dip_b200 import _lib
    return _lib.DipError


def test_layer0_instance_token_planes_are_hoisted_without_changing_results(cuda):
    """Tensor-core mode keeps the q|k|v planes of the L instance tokens of decoder layer 0 in the workspace after the first decode of a
    wave (they depend on the instance and the weights only) and projects just the x tokens afterwards: the second, hoisted, evaluation
    must equal the first bit for bit, and so must one after an explicit cache reset."""
    from dip_b200.engine import Engine
    cs = small_case(seed=3, L=150, n=130, M=30, nnz=200, B=5, dup=0)
    eng = make_engine(cuda, cs, precision="tc")
    x = torch.from_numpy(cs.x).to(cuda)
    g1 = eng.guidance_grad(x, 0.5)          # first evaluation: full layer-0 in-projection
    g2 = eng.guidance_grad(x, 0.5)          # hoisted
    eng.reset_cache()
    g3 = eng.guidance_grad(x, 0.5)
    for a, b, c in zip(g1, g2, g3):
        assert torch.equal(a, b) and torch.equal(a, c)
    x2 = torch.from_numpy(small_case(seed=4, L=150, n=130, M=30, nnz=200, B=5, dup=0).x).to(cuda)
    h1 = eng.guidance_grad(x2, 0.5)         # other latents through the hoisted path ...
    eng.reset_cache()
    h2 = eng.guidance_grad(x2, 0.5)         # ... and through the full one
    for a, b in zip(h1, h2):
        assert torch.equal(a, b)
