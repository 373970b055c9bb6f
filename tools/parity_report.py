"""Prints the measured one-step parity errors (relative L2) of both precision modes against the golden
vectors of the unmodified reference (small case) and against the CPU oracle at full Set-Cover size."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-integer-programming_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from dip_b200 import synth, schedule
from dip_b200.engine import Engine
from oracle import restate
from helpers import golden, tw, rel_l2, small_case, sc_case

dev = torch.device("cuda")
out = {}
g = golden("ddim_guided_x0"); cs = small_case()
S, gs, lam = int(g["S"]), float(g["gs"]), float(g["lam"])
for mode in ("fp32", "tc"):
    eng = Engine(); eng.set_precision(mode)
    eng.set_denoiser(tw(synth.denoiser_weights(1, seed=1), dev)); eng.set_decoder(tw(synth.decoder_weights(2, seed=2), dev))
    eng.set_instance(torch.from_numpy(cs.mip).to(dev), torch.from_numpy(cs.kpm).to(dev), cs.rows, cs.cols, cs.vals, cs.M, cs.b, cs.c)
    r = {}
    for index in (0, 50, 99):
        x = torch.from_numpy(cs.x).to(dev)
        x0 = eng.denoise(x, S - 1 - index); grad, p, _, _ = eng.guidance_grad(x, lam)
        xs = x.clone(); eng.guided_ddim_step(xs, schedule.ddim_coeffs(S)[index], gs, lam)
        r["golden_idx%d" % index] = {"x0": rel_l2(x0, g["tf%d_x0" % index]), "p": rel_l2(p[:, :cs.n], g["tf%d_p" % index]),
                                     "grad": rel_l2(grad, g["tf%d_grad" % index]), "x_prev": rel_l2(xs, g["tf%d_xprev" % index])}
    out[mode] = r
inst, mip, kpm, x = sc_case(B=2)
Wd, Wdec = synth.denoiser_weights(1, seed=1), synth.decoder_weights(2, seed=2)
tab = restate.ddim_tables(restate.trainer_buffers()["alphas_cumprod"], 100)
A = (torch.from_numpy(inst.rows), torch.from_numpy(inst.cols), torch.from_numpy(inst.a_val), inst.M)
mipB, kpmB = torch.from_numpy(mip)[None].repeat(2, 1, 1), torch.from_numpy(kpm)[None].repeat(2, 1)
ref = restate.guided_ddim_step(tw(Wd), tw(Wdec), torch.from_numpy(x), 99, restate.ddim_step_coeffs(tab, 0), mipB, kpmB, A,
                               torch.from_numpy(inst.b), torch.from_numpy(inst.c), 1e5, 0.9)
for mode in ("fp32", "tc"):
    eng = Engine(); eng.set_precision(mode)
    eng.set_denoiser(tw(Wd, dev)); eng.set_decoder(tw(Wdec, dev))
    eng.set_instance(torch.from_numpy(mip).to(dev), torch.from_numpy(kpm).to(dev), inst.rows, inst.cols, inst.a_val, inst.M, inst.b, inst.c)
    xg = torch.from_numpy(x).to(dev).clone(); x0 = torch.empty_like(xg)
    grad, p, _, _ = eng.guidance_grad(xg, 0.9)
    eng.guided_ddim_step(xg, schedule.ddim_coeffs(100)[0], 1e5, 0.9, x0_out=x0)
    out[mode]["full_size_SC_vs_oracle"] = {"x0": rel_l2(x0, ref["x0"]), "p": rel_l2(p, ref["p_full"]), "grad": rel_l2(grad, ref["grad"]),
                                           "x_prev": rel_l2(xg, ref["x_prev"])}
# the reference's shipped (trained) IS checkpoints at configs[2] shapes and the CF-length step, against the unmodified reference's outputs
def strided(got, ref, stride):
    return rel_l2(got[:, ::stride], ref)

GOLD = os.path.join(ROOT, "tests", "golden")
ck = lambda n: torch.load(os.path.join(GOLD, n), map_location="cpu", weights_only=True)
g = golden("is_ckpt")
n = L = 1500
inst = synth.independent_set(n_nodes=n, n_edges=4860, n_single=n, seed=7)
mip = synth.mip_features(L, n, seed=17); kpm = np.zeros(L, dtype=bool)
x = synth.initial_noise(7, range(2), L)
for mode in ("fp32", "tc"):
    eng = Engine(); eng.set_precision(mode)
    eng.set_denoiser({k: v.to(dev) for k, v in ck("ddpm_independent_set.pth").items()}); eng.set_decoder({k: v.to(dev) for k, v in ck("decoder_independent_set.pth").items()})
    eng.set_instance(torch.from_numpy(mip).to(dev), torch.from_numpy(kpm).to(dev), inst.rows, inst.cols, inst.a_val, inst.M, inst.b, inst.c)
    r = {}
    for index in (0, 50):
        xg = torch.from_numpy(x).to(dev).clone(); x0 = torch.empty_like(xg)
        grad, p, _, _ = eng.guidance_grad(xg, float(g["lam"]))
        eng.guided_ddim_step(xg, schedule.ddim_coeffs(100)[index], float(g["gs"]), float(g["lam"]), x0_out=x0)
        st = int(g["stride"])
        r["idx%d" % index] = {"x0": strided(x0, g["tf%d_x0" % index], st), "p": rel_l2(p[:, :n], g["tf%d_p" % index]),
                              "grad": strided(grad, g["tf%d_grad" % index], st), "x_prev": strided(xg, g["tf%d_xprev" % index], st)}
    fd = eng.final_decode(torch.from_numpy(g["xprev_full"]).to(dev))
    r["decode"] = {"p": rel_l2(fd["p_full"][:, :n], g["dec_xprev_p"]), "min_margin_|p-0.5|": float(np.abs(g["dec_xprev_p"] - 0.5).min()),
                   "bits_equal": bool(np.array_equal(fd["xbits"].cpu().numpy(), g["final_x"].astype(np.uint8)))}
    out[mode]["shipped_IS_checkpoints_1500_6360_vs_reference"] = r
g = golden("cf_step")
inst = synth.facility_location(100, 50, seed=9)
L = inst.n
mip = synth.mip_features(L, L, seed=19); kpm = np.zeros(L, dtype=bool)
x = synth.initial_noise(9, range(1), L)
for mode in ("fp32", "tc"):
    eng = Engine(); eng.set_precision(mode)
    eng.set_denoiser(tw(synth.denoiser_weights(1, seed=1), dev)); eng.set_decoder(tw(synth.decoder_weights(2, seed=2), dev))
    eng.set_instance(torch.from_numpy(mip).to(dev), torch.from_numpy(kpm).to(dev), inst.rows, inst.cols, inst.a_val, inst.M, inst.b, inst.c)
    xg = torch.from_numpy(x).to(dev).clone(); x0 = torch.empty_like(xg)
    grad, p, _, _ = eng.guidance_grad(xg, float(g["lam"]))
    eng.guided_ddim_step(xg, schedule.ddim_coeffs(100)[0], float(g["gs"]), float(g["lam"]), x0_out=x0)
    st = int(g["stride"])
    out[mode]["CF_5050_5151_vs_reference"] = {"x0": strided(x0, g["tf0_x0"], st), "p": rel_l2(p, g["tf0_p"]), "grad": strided(grad, g["tf0_grad"], st),
                                              "x_prev": strided(xg, g["tf0_xprev"], st)}
print(json.dumps(out, indent=1))
