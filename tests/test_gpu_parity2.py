"""GPU: round-2 parity cases against outputs of the UNMODIFIED reference (tests/golden/*, written by oracle/gen_golden2.py):

* the two checkpoints the reference ships (trained Independent-Set weights: two denoiser layers, select_net head) on configs[2] shapes
  (n = L = 1500, M = 6360) -- P1 teacher-forced steps, the select_net branch of the decoder, P3 decode / feasibility;
* configs[3]'s longest sequence (CF, L = 5050, decoder sequence 10 100) -- one guided step against the reference itself;
* P4: end-to-end statistical parity of whole S = 100 guided runs (feasible-solution rate and integer-objective distribution over
  4 instances x 256 samples per family) + P2 latents of the first free-running steps from identical noise;
* the drop-in DDPMSampler against the oracle with the sampler's own noise; two model objects sharing one engine.
"""
import json
import os

import numpy as np
import pytest
import torch

from dip_b200 import synth
from helpers import GOLDEN, golden, tw, rel_l2, small_case, case_tensors, coalesced

pytestmark = pytest.mark.gpu

MODES = ["fp32", "tc"]
TOLS = {"fp32": 1e-4, "tc": 1e-3}            # rel-L2, one step (north_star: <= 1e-3 in fp32 mode; measured ~1e-6 / ~1e-5)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ckpt(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=True)


def ddim_co(S, index):
    from dip_b200 import schedule
    return schedule.ddim_coeffs(S)[index]


def strided_check(got, ref_strided, ref_norm, stride, tol, what):
    """Elementwise rel-L2 on the stored rows (every `stride`-th token) + the whole-tensor norm."""
    assert rel_l2(got[:, ::stride], ref_strided) < tol, (what, rel_l2(got[:, ::stride], ref_strided))
    assert abs(float(got.double().norm()) - ref_norm) < tol * ref_norm, what


# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", MODES)
def test_shipped_is_checkpoints_config2_vs_reference(cuda, mode):
    """agents/model_hub/{ddpm,decoder}_independent_set.pth through the engine at IS config shapes (1500 / 1500 / 6360), gs = 2e4,
    lambda = 0.5 (the reference's tuned IS setting): x0-hat, p, gradient and x_prev of teacher-forced steps at schedule index 0 and 50,
    then final decode -> identical bits, penalties and integer checks on the reference's own latent."""
    from dip_b200.engine import Engine
    from oracle import restate
    g = golden("is_ckpt")
    TOL = TOLS[mode]
    n = L = 1500
    inst = synth.independent_set(n_nodes=n, n_edges=4860, n_single=n, seed=7)
    assert inst.M == int(g["M"]) == 6360
    mip = synth.mip_features(L, n, seed=17)
    kpm = np.zeros(L, dtype=bool)
    B, S, gs, lam, stride = 2, int(g["S"]), float(g["gs"]), float(g["lam"]), int(g["stride"])
    x = synth.initial_noise(7, range(B), L)
    assert abs(float(np.abs(x).sum()) - float(g["x_checksum"])) < 1e-3 * float(g["x_checksum"])
    sd_t, sd_d = ckpt("ddpm_independent_set.pth"), ckpt("decoder_independent_set.pth")
    eng = Engine()
    eng.set_precision(mode)
    eng.set_denoiser({k: v.to(cuda) for k, v in sd_t.items()})
    eng.set_decoder({k: v.to(cuda) for k, v in sd_d.items()})
    assert eng.n_den == 2 and eng.n_dec == 2
    eng.set_instance(torch.from_numpy(mip).to(cuda), torch.from_numpy(kpm).to(cuda), inst.rows, inst.cols, inst.a_val, inst.M, inst.b, inst.c,
                     a_int=inst.a_int, b_int=inst.b_int, c_int=inst.c_int)
    for index in g["indices"].tolist():
        pre = "tf%d_" % index
        norms = g[pre + "norms"]
        xg = torch.from_numpy(x).to(cuda).clone()
        x0 = torch.empty_like(xg)
        grad, p_full, _, _ = eng.guidance_grad(xg, lam)
        assert rel_l2(p_full[:, :n], g[pre + "p"]) < TOL
        strided_check(grad, g[pre + "grad"], norms[1], stride, TOL, "grad %d" % index)
        eng.guided_ddim_step(xg, ddim_co(S, index), gs, lam, x0_out=x0)
        strided_check(x0, g[pre + "x0"], norms[0], stride, TOL, "x0 %d" % index)
        strided_check(xg, g[pre + "xprev"], norms[2], stride, TOL, "xprev %d" % index)
    # P3: the reference's latent in -> bit-identical assignment, feasibility flag and integer objective out
    lat = torch.from_numpy(g["xprev_full"]).to(cuda)
    fd = eng.final_decode(lat)
    assert rel_l2(fd["p_full"][:, :n], g["dec_xprev_p"]) < TOL
    margin = np.abs(g["dec_xprev_p"] - 0.5)
    safe = margin > 1e-4
    bits = fd["xbits"].cpu().numpy()
    assert np.array_equal(bits[safe], g["final_x"].astype(np.uint8)[safe])
    feas, objv, violv = restate.feasibility_int(bits, inst.rows, inst.cols, inst.a_int, inst.M, inst.b_int, inst.c_int)
    assert np.array_equal(fd["viol_int"].cpu().numpy(), violv) and np.array_equal(fd["obj_int"].cpu().numpy(), objv)
    if safe.all():
        assert np.allclose(fd["penalty"].cpu().numpy(), g["final_penalty"], rtol=1e-5, atol=1e-5)
        assert np.array_equal(fd["penalty"].cpu().numpy() <= 1e-5, g["final_penalty"] <= 1e-5)


def test_select_net_branch_of_dropin_decoder_vs_reference(cuda):
    """MipConditionalDecorder(use_select_net=True).forward (reference: decoder.py:52-56) with the shipped decoder checkpoint."""
    from model.decoder import MipConditionalDecorder
    g = golden("is_ckpt")
    n = L = 1500
    mip = synth.mip_features(L, n, seed=17)
    kpm = np.zeros(L, dtype=bool)
    B = 2
    dec = MipConditionalDecorder(128, 1, 2, use_select_net=True).to(cuda)
    dec.load_state_dict(ckpt("decoder_independent_set.pth"))
    dec.eval()
    mipB = torch.from_numpy(mip).to(cuda)[None].repeat(B, 1, 1)
    kpmB = torch.from_numpy(kpm).to(cuda)[None].repeat(B, 1)
    for tag, lat in (("dec_x", synth.initial_noise(7, range(B), L)), ("dec_xprev", g["xprev_full"])):
        sols, sel = dec(mipB, torch.from_numpy(lat).to(cuda), kpmB)
        assert sols.shape == (B * n,) and sel.shape == (B * n,)
        assert rel_l2(sols.view(B, n), g[tag + "_p"]) < 1e-3, tag
        assert rel_l2(sel.view(B, n), g[tag + "_sel"]) < 1e-3, tag


@pytest.mark.parametrize("mode", MODES)
def test_facility_location_length_vs_reference(cuda, mode):
    """CF (5151 x 5050, 20 100 non-zeros with non-unit coefficients; decoder sequence 10 100), gs = 1e3, lambda = 0.7: one guided step
    against the unmodified reference (fixture cf_step.npz)."""
    from dip_b200.engine import Engine
    g = golden("cf_step")
    TOL = TOLS[mode]
    inst = synth.facility_location(100, 50, seed=9)
    assert (inst.M, inst.n, inst.nnz) == (5151, 5050, 20100) == (int(g["M"]), 5050, int(g["nnz"]))
    n = L = inst.n
    mip = synth.mip_features(L, n, seed=19)
    kpm = np.zeros(L, dtype=bool)
    S, gs, lam, stride = int(g["S"]), float(g["gs"]), float(g["lam"]), int(g["stride"])
    x = synth.initial_noise(9, range(1), L)
    assert abs(float(np.abs(x).sum()) - float(g["x_checksum"])) < 1e-3 * float(g["x_checksum"])
    eng = Engine()
    eng.set_precision(mode)
    eng.set_denoiser(tw(synth.denoiser_weights(1, seed=1), cuda)); eng.set_decoder(tw(synth.decoder_weights(2, seed=2), cuda))
    eng.set_instance(torch.from_numpy(mip).to(cuda), torch.from_numpy(kpm).to(cuda), inst.rows, inst.cols, inst.a_val, inst.M, inst.b, inst.c,
                     a_int=inst.a_int, b_int=inst.b_int, c_int=inst.c_int)
    norms = g["tf0_norms"]
    xg = torch.from_numpy(x).to(cuda).clone()
    x0 = torch.empty_like(xg)
    grad, p_full, _, _ = eng.guidance_grad(xg, lam)
    assert rel_l2(p_full, g["tf0_p"]) < TOL
    strided_check(grad, g["tf0_grad"], norms[1], stride, TOL, "grad")
    eng.guided_ddim_step(xg, ddim_co(S, 0), gs, lam, x0_out=x0)
    strided_check(x0, g["tf0_x0"], norms[0], stride, TOL, "x0")
    strided_check(xg, g["tf0_xprev"], norms[2], stride, TOL, "xprev")


# --------------------------------------------------------------------------------------------------
def test_ddpm_sampler_level_vs_oracle(cuda):
    """DDPMSampler.constraint_guided_sample / .sample (drop-in, reference: diffusion.py:415-461) on a short chain against
    restate.guided_ddpm_step / unguided_ddpm_step fed with the very noise the sampler drew (torch's CUDA generator replayed)."""
    from model.diffusion import DDPMSampler, DDPMTrainer
    from model.decoder import MipConditionalDecorder
    from oracle import restate
    T = 6
    cs = small_case(B=2)
    Wd, Wdec = synth.denoiser_weights(1, seed=1), synth.decoder_weights(2, seed=2)
    tr = DDPMTrainer(128, 1, 1, cuda, timesteps=T)
    tr.load_state_dict(tw(Wd), strict=False)
    dec = MipConditionalDecorder(128, 1, 2).to(cuda)
    dec.load_state_dict(tw(Wdec))
    t = case_tensors(cs, cuda)
    tc_ = case_tensors(cs)
    A = torch.sparse_coo_tensor(torch.from_numpy(np.stack([cs.rows, cs.cols])), torch.from_numpy(cs.vals), size=(cs.M, cs.n)).to(cuda)
    gs, lam = 1.5e4, 0.1
    s = DDPMSampler(tr, dec, gradient_scale=gs, obj_guided_coef=lam, device=cuda)
    buf = restate.trainer_buffers(T)
    tabs = restate.ddpm_tables(buf)
    for guided in (True, False):
        torch.manual_seed(11)
        x = s.constraint_guided_sample(t["mipB"], t["kpmB"], A, t["b"], t["c"]) if guided else s.sample(t["mipB"], t["kpmB"])
        # replay the generator: initial latent, then one draw per step (diffusion.py:424,447 / 407,458)
        torch.manual_seed(11)
        xr = torch.randn_like(t["mipB"]).cpu()
        for i in reversed(range(T)):
            noise = torch.randn(t["mipB"].shape, device=cuda).cpu()
            if guided:
                xr = restate.guided_ddpm_step(tw(Wd), tw(Wdec), xr, i, tabs, buf, tc_["mipB"], tc_["kpmB"], coalesced(cs), tc_["b"], tc_["c"],
                                              gs, lam, noise)["x_prev"]
            else:
                xr = restate.unguided_ddpm_step(tw(Wd), xr, i, tabs, buf, tc_["mipB"], tc_["kpmB"], noise)["x_prev"]
        assert rel_l2(x[:, :cs.n], xr[:, :cs.n]) < 5e-3, (guided, rel_l2(x[:, :cs.n], xr[:, :cs.n]))


def test_two_models_sharing_the_engine_alternate(cuda):
    """Two trainers / decoders with different weights used in turn (A, B, A): each call must run with its own weights
    (the engine, not the module, records whose weights it holds)."""
    from model.diffusion import DDPMTrainer, DDIMSampler
    from model.decoder import MipConditionalDecorder
    cs = small_case(B=2)
    t = case_tensors(cs, cuda)
    A = torch.sparse_coo_tensor(torch.from_numpy(np.stack([cs.rows, cs.cols])), torch.from_numpy(cs.vals), size=(cs.M, cs.n)).to(cuda)

    def build(seed):
        tr = DDPMTrainer(128, 1, 1, cuda)
        tr.load_state_dict(tw(synth.denoiser_weights(1, seed=seed)), strict=False)
        dec = MipConditionalDecorder(128, 1, 2).to(cuda)
        dec.load_state_dict(tw(synth.decoder_weights(2, seed=seed + 1)))
        s = DDIMSampler(tr, dec, 1e4, 0.5, cuda)
        s.initial_noise = t["x"].clone()
        return s, dec

    (sa, da), (sb, db) = build(1), build(31)
    run = lambda s: s.constraint_guided_sample(t["mipB"], t["kpmB"], A, t["b"], t["c"], S=5)[0].clone()
    xa1, xb1, xa2, xb2 = run(sa), run(sb), run(sa), run(sb)
    assert torch.equal(xa1, xa2) and torch.equal(xb1, xb2) and not torch.equal(xa1, xb1)
    pa1 = da(t["mipB"], t["x"], t["kpmB"])[0].clone()
    pb1 = db(t["mipB"], t["x"], t["kpmB"])[0].clone()
    pa2 = da(t["mipB"], t["x"], t["kpmB"])[0].clone()
    assert torch.equal(pa1, pa2) and not torch.equal(pa1, pb1)
    # in-place weight change is picked up as well
    with torch.no_grad():
        da.linear_proj.linear_projection.bias.add_(1.0)
    assert not torch.equal(da(t["mipB"], t["x"], t["kpmB"])[0], pa1)


# --------------------------------------------------------------------------------------------------
def _p4_run(cuda, kind, mode):
    """The P4 workload on the GPU: same instances, weights and initial noise as oracle/gen_golden2.gen_p4."""
    from oracle.gen_golden2 import p4_cases, P4
    from dip_b200 import schedule
    from dip_b200.engine import Engine
    cfg = P4[kind]
    B, S = cfg["B"], cfg["S"]
    eng = Engine()
    eng.set_precision(mode)
    if kind == "sc":
        eng.set_denoiser(tw(synth.denoiser_weights(1, seed=1), cuda)); eng.set_decoder(tw(synth.decoder_weights(2, seed=2), cuda))
    else:
        eng.set_denoiser({k: v.to(cuda) for k, v in ckpt("ddpm_independent_set.pth").items()})
        eng.set_decoder({k: v.to(cuda) for k, v in ckpt("decoder_independent_set.pth").items()})
    steps = schedule.ddim_coeffs(S)
    res = []
    for i, (inst, mip, kpm, L) in enumerate(p4_cases(kind)):
        eng.set_instance(torch.from_numpy(mip).to(cuda), torch.from_numpy(kpm).to(cuda), inst.rows, inst.cols, inst.a_val, inst.M, inst.b, inst.c,
                         a_int=inst.a_int, b_int=inst.b_int, c_int=inst.c_int)
        x0 = synth.initial_noise(cfg["inst_base"] + i, range(B), L)
        # P2: first free-running steps of 8 samples, step by step
        xs = torch.from_numpy(x0[:8]).to(cuda).clone()
        early = {}
        for k in range(5):
            eng.guided_ddim_step(xs, steps[k], cfg["gs"], cfg["lam"])
            if k + 1 in (1, 2, 5):
                early[k + 1] = xs.clone()
        out = eng.sample_guided_ddim(torch.from_numpy(x0).to(cuda).clone(), steps, cfg["gs"], cfg["lam"])
        res.append({"feasible": (out["viol_int"] == 0).cpu().numpy(), "feasible_fp": (out["penalty"] <= 1e-5).cpu().numpy(),
                    "obj_int": out["obj_int"].cpu().numpy(), "bits": out["xbits"].cpu().numpy(), "early": early, "inst": inst})
    return res


def _p4_compare(kind, mode, res, g):
    from scipy import stats
    rep = {"kind": kind, "mode": mode, "instances": []}
    n_all = f_gpu_all = f_ref_all = 0
    for i, r in enumerate(res):
        ref_feas = g["i%d_feasible" % i]
        ref_obj = g["i%d_obj_int" % i]
        B = len(ref_feas)
        # integer flag == the reference's fp penalty test on the same assignment
        assert np.array_equal(r["feasible"], r["feasible_fp"])
        early = {k: rel_l2(v[:, :r["inst"].n], g["i%d_x_after%d" % (i, k)][:, :r["inst"].n]) for k, v in r["early"].items()}
        pg, pr = float(r["feasible"].mean()), float(ref_feas.mean())
        pooled = (r["feasible"].sum() + ref_feas.sum()) / (2.0 * B)
        se = float(np.sqrt(max(2.0 * pooled * (1.0 - pooled) / B, 1e-12)))
        og, orf = r["obj_int"][r["feasible"]], ref_obj[ref_feas]
        ks = stats.ks_2samp(og, orf) if len(og) >= 8 and len(orf) >= 8 else None
        same_bits = float((np.packbits(r["bits"], axis=1) == g["i%d_bits" % i]).all(axis=1).mean())
        rep["instances"].append({
            "instance": i, "samples": B, "feasible_ratio_gpu": pg, "feasible_ratio_reference": pr, "diff_in_sigma": abs(pg - pr) / se if se > 0 else 0.0,
            "obj_mean_gpu": float(og.mean()) if len(og) else None, "obj_mean_reference": float(orf.mean()) if len(orf) else None,
            "obj_std_gpu": float(og.std()) if len(og) else None, "obj_std_reference": float(orf.std()) if len(orf) else None,
            "ks_stat": float(ks.statistic) if ks else None, "ks_p": float(ks.pvalue) if ks else None,
            "identical_assignments_frac": same_bits, "p2_rel_l2_after_steps": {str(k): v for k, v in early.items()}})
        n_all += B; f_gpu_all += int(r["feasible"].sum()); f_ref_all += int(ref_feas.sum())
    rep["total"] = {"samples": n_all, "feasible_ratio_gpu": f_gpu_all / n_all, "feasible_ratio_reference": f_ref_all / n_all}
    return rep


@pytest.mark.parametrize("kind,mode", [("sc", "tc"), ("sc", "fp32"), ("is", "tc")])
def test_p4_statistical_parity(cuda, kind, mode):
    """north_star: 'identical feasible-solution rate and objective distribution'.  1024 samples per family (4 instances x 256), whole
    S = 100 guided DDIM runs from the same initial noise as the CPU reference: per instance the feasible ratio must agree within 4 sigma
    of the binomial difference, the integer objectives of the feasible samples must pass a two-sample KS test (p > 1e-3), and the
    latents after 1 / 2 / 5 free-running steps must agree (P2; tolerances below)."""
    g = golden("p4_%s_small" % kind)
    res = _p4_run(cuda, kind, mode)
    rep = _p4_compare(kind, mode, res, g)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "p4_report_%s_%s.json" % (kind, mode)), "w") as f:
            json.dump(rep, f, indent=1)
    for it in rep["instances"]:
        assert it["diff_in_sigma"] < 4.0, it
        if it["ks_p"] is not None:
            assert it["ks_p"] > 1e-3, it
        for k, v in it["p2_rel_l2_after_steps"].items():
            # P2.  One step from identical noise: 1e-3 (north_star).  Later free-running steps: random-init weights stay below 1e-3 per
            # step; the trained IS checkpoints amplify rounding noise ~400x per step -- the CPU reference against an fp64 run of itself
            # reads 3.4e-4 after 2 steps and up to 1e-2 after 5 (tests/test_oracle_golden2.py) -- so only a loose bound is meaningful there.
            # (profiles/r02_p2_noise_floor_is.txt: the fp32 reference against an fp64 run of itself reads 2e-4..7e-4 after 2 steps and
            # 4e-4..0.14 after 5 -- instance 1 is O(0.1) away from ITSELF -- so step 5 is reported, not asserted, for the trained weights.)
            tol = 1e-3 * max(1, int(k)) if kind == "sc" else {"1": 1e-3, "2": 5e-2, "5": float("inf")}[k]
            assert v < tol, it
