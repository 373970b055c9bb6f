"""TEST INFRASTRUCTURE ONLY.  Round-2 fixtures, again outputs of the UNMODIFIED reference
(/root/reference through oracle/shims.py), run in the build container with one MKL thread:

    python -m oracle.gen_golden2 [is_ckpt] [cf_step] [p4_sc_small] [p4_is_small] [ndiving] [solenc]

* is_ckpt      the two checkpoints the reference ships (agents/model_hub/{ddpm,decoder}_independent_set.pth: two denoiser layers,
               use_select_net=True; copied byte-for-byte into tests/golden/) on configs[2] shapes (Independent Set, n = L = 1500,
               M = 6360): teacher-forced guided DDIM steps, the decoder's select_net head, final decode.
* cf_step      configs[3]'s longest sequence (CF: L = 5050, decoder sequence 10 100): one guided DDIM step.
* p4_*_small   P4, end-to-end statistical parity: whole S = 100 guided runs of 4 instances x 256 samples through the reference's
               DDIMSampler.constraint_guided_sample + decode + round + penalty (sample.py:151-174): per-sample feasibility flags and
               integer objectives, plus the latents of the first free-running steps (P2).  The shapes are small because the CPU
               reference needs ~25 s per full-size SC sample; the distributions are what the GPU run is compared against.
* ndiving      NeuralDiving head on the shared GNN (gnn_model/gnn_model.py:61-86).
* solenc       SolutionEncoder.forward / CMSP.encode_solution (model/encoder.py:24-36, model/cmsp.py:65-69).

Large tensors are stored row-strided (every ROW_STRIDE-th token) to keep the fixtures small; probabilities, residuals and integer
outputs are stored whole.  Inputs are regenerated from seeds by dip_b200.synth (numpy, portable) at test time.
"""
import os
import shutil
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-integer-programming_b200"))

from oracle import ref_loader  # noqa: E402
from oracle.gen_golden import Recorder, load_weights, tens, small_graph  # noqa: E402
from dip_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
D = 128
CKPT = {"ddpm": "ddpm_independent_set.pth", "decoder": "decoder_independent_set.pth"}


def stage_checkpoints(R):
    """Byte copies of the shipped checkpoints (data artefacts of the reference, SURVEY section 2.1 'Artefacts')."""
    for f in CKPT.values():
        dst = os.path.join(GOLDEN, f)
        if not os.path.isfile(dst):
            shutil.copyfile(os.path.join(R.root, "agents", "model_hub", f), dst)
            os.chmod(dst, 0o644)


def load_ckpt(name):
    return torch.load(os.path.join(GOLDEN, CKPT[name]), map_location="cpu", weights_only=True)


def trained_models(R):
    stage_checkpoints(R)
    sd_t, sd_d = load_ckpt("ddpm"), load_ckpt("decoder")
    n_den = 1 + max(int(k.split(".")[2]) for k in sd_t if k.startswith("predict_model.resblocks."))
    n_dec = 1 + max(int(k.split(".")[1]) for k in sd_d if k.startswith("resblocks."))
    tr = R.DDPMTrainer(D, 1, n_den, "cpu")
    tr.load_state_dict(sd_t)
    dec = R.MipConditionalDecorder(D, 1, n_dec, use_select_net=True)
    dec.load_state_dict(sd_d)
    tr.eval(); dec.eval()
    return tr, dec


def inst_tensors(inst, mip, kpm, B):
    A = torch.sparse_coo_tensor(tens(np.stack([inst.rows, inst.cols])), tens(inst.a_val), size=(inst.M, inst.n))
    return (tens(mip)[None].repeat(B, 1, 1), tens(kpm)[None].repeat(B, 1), A, tens(inst.b), tens(inst.c))


def penalty_of(R, A, xr, b):
    import torch_sparse
    Ac = A.coalesce()
    return torch.max(torch_sparse.spmm(Ac.indices(), Ac.values(), A.size()[0], A.size()[1], xr).squeeze(-1) - b, torch.tensor(0)).sum(axis=-1)


def guided_steps(R, tr, dec, inst, mip, kpm, x, gs, lam, S, indices, stride, out, tag=""):
    B = x.shape[0]
    rec = Recorder(dec)
    s = R.DDIMSampler(tr, rec, gradient_scale=gs, obj_guided_coef=lam, device="cpu")
    mipB, kpmB, A, b, c = inst_tensors(inst, mip, kpm, B)
    s.make_schedule(S)
    s.predict_model, s.batch_size, s.key_padding_mask, s.conditions = tr.predict_model, B, kpmB, mipB
    xp = None
    for index in indices:
        t0 = time.time()
        xp, x0 = s.consraint_guided_p_sample_ddim(tens(x).clone(), S - 1 - index, index, A, b, c)
        pre = "%stf%d_" % (tag, index)
        out[pre + "x0"] = x0.detach().numpy()[:, ::stride].copy()
        out[pre + "grad"] = rec.x_t.grad.detach().numpy()[:, ::stride].copy()
        out[pre + "xprev"] = xp.detach().numpy()[:, ::stride].copy()
        out[pre + "p"] = rec.out[0].detach().view(B, inst.n).numpy()
        # whole-tensor norms: the strided comparison is backed by a global one
        out[pre + "norms"] = np.array([float(x0.double().norm()), float(rec.x_t.grad.double().norm()), float(xp.double().norm())])
        print("  step index %d: %.1f s" % (index, time.time() - t0), flush=True)
    return s, rec, (mipB, kpmB, A, b, c), xp.detach()


def gen_is_ckpt(R, name="is_ckpt", stride=4):
    tr, dec = trained_models(R)
    n = L = 1500
    inst = synth.independent_set(n_nodes=n, n_edges=4860, n_single=n, seed=7)
    mip = synth.mip_features(L, n, seed=17)
    kpm = np.zeros(L, dtype=bool)
    B, S, gs, lam = 2, 100, 2e4, 0.5
    x = synth.initial_noise(7, range(B), L)
    out = {"S": S, "gs": gs, "lam": lam, "stride": stride, "indices": np.array([0, 50]), "M": inst.M, "nnz": inst.nnz,
           "x_checksum": float(np.abs(x).sum()), "mip_checksum": float(np.abs(mip).sum())}
    s, rec, (mipB, kpmB, A, b, c), xp = guided_steps(R, tr, dec, inst, mip, kpm, x, gs, lam, S, (0, 50), stride, out)
    with torch.no_grad():
        # decoder with the select_net head on the noisy latent and on the step's output (decoder.py:52-56)
        for tag, lat in (("dec_x", tens(x)), ("dec_xprev", xp)):
            sols, sel = dec(mipB, lat, kpmB)
            out[tag + "_p"] = sols.view(B, n).numpy()
            out[tag + "_sel"] = sel.view(B, n).numpy()
        xr = torch.round(sols).view(B, -1, 1)
        out["final_x"] = xr.view(B, n).numpy()
        out["final_penalty"] = penalty_of(R, A, xr, b).numpy()
        out["xprev_full"] = xp.numpy()          # the latent the final decode ran on (P3 needs identical latents in)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("wrote", name)


def gen_cf_step(R, name="cf_step", stride=16):
    from oracle.gen_golden import build_reference
    inst = synth.facility_location(100, 50, seed=9)
    n = L = inst.n
    mip = synth.mip_features(L, n, seed=19)
    kpm = np.zeros(L, dtype=bool)
    B, S, gs, lam = 1, 100, 1e3, 0.7
    tr, dec = build_reference(R, 1, 2, "x0", gs, lam)
    x = synth.initial_noise(9, range(B), L)
    out = {"S": S, "gs": gs, "lam": lam, "stride": stride, "indices": np.array([0]), "M": inst.M, "nnz": inst.nnz,
           "x_checksum": float(np.abs(x).sum()), "mip_checksum": float(np.abs(mip).sum())}
    guided_steps(R, tr, dec, inst, mip, kpm, x, gs, lam, S, (0,), stride, out)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("wrote", name)


def p4_cases(kind):
    """(instance, mip, kpm, L) for the four P4 instances of a family; shared with the tests."""
    cases = []
    for i in range(4):
        if kind == "sc":
            inst = synth.set_cover(40, 80, 0.1, seed=100 + i)
            L = 96
        else:
            inst = synth.independent_set(n_nodes=120, n_edges=400, n_single=120, seed=200 + i)
            L = 128
        mip = synth.mip_features(L, inst.n, seed=300 + i)
        kpm = np.zeros(L, dtype=bool)
        kpm[inst.n:] = True
        cases.append((inst, mip, kpm, L))
    return cases


P4 = {"sc": {"gs": 1e5, "lam": 0.9, "B": 256, "S": 100, "inst_base": 100},
      "is": {"gs": 2e4, "lam": 0.5, "B": 256, "S": 100, "inst_base": 200}}


def gen_p4(R, kind):
    from oracle.gen_golden import build_reference
    cfg = P4[kind]
    B, S, gs, lam = cfg["B"], cfg["S"], cfg["gs"], cfg["lam"]
    if kind == "sc":
        tr, dec = build_reference(R, 1, 2, "x0", gs, lam)
    else:
        tr, dec = trained_models(R)
    out = {"S": S, "gs": gs, "lam": lam, "B": B}
    for i, (inst, mip, kpm, L) in enumerate(p4_cases(kind)):
        t0 = time.time()
        x0 = synth.initial_noise(cfg["inst_base"] + i, range(B), L)
        mipB, kpmB, A, b, c = inst_tensors(inst, mip, kpm, B)
        s = R.DDIMSampler(tr, dec, gradient_scale=gs, obj_guided_coef=lam, device="cpu")
        s.initial_noise = tens(x0)
        x, inter = s.constraint_guided_sample(mipB, kpmB, A, b, c, S=S)            # sample.py:157
        with torch.no_grad():
            pred, _ = dec(mipB, x, kpmB)                                           # sample.py:162-164
            xr = torch.round(pred).view(B, -1, 1)
            pen = penalty_of(R, A, xr, b)                                          # sample_partialsol.py:167-168
        bits = xr.view(B, inst.n).numpy().astype(np.int64)
        out["i%d_feasible" % i] = (pen <= 1e-5).numpy()                            # sample.py:174
        out["i%d_penalty" % i] = pen.numpy()
        out["i%d_obj_int" % i] = bits @ inst.c_int.astype(np.int64)
        out["i%d_bits" % i] = np.packbits(bits.astype(np.uint8), axis=1)
        out["i%d_p_margin" % i] = float((pred - 0.5).abs().min())
        # P2: the first free-running steps of 8 samples (intermediates[k] = latent after k steps, diffusion.py:587-591)
        for k in (1, 2, 5):
            out["i%d_x_after%d" % (i, k)] = inter[k][:8].detach().numpy()
        for p in dec.parameters():
            p.grad = None                                                          # the reference accumulates unused weight grads
        print("  %s instance %d: %.0f s, feasible %.3f, mean obj (feasible) %.2f" % (
            kind, i, time.time() - t0, float(out["i%d_feasible" % i].mean()),
            float(out["i%d_obj_int" % i][out["i%d_feasible" % i]].mean()) if out["i%d_feasible" % i].any() else float("nan")), flush=True)
    np.savez_compressed(os.path.join(GOLDEN, "p4_%s_small.npz" % kind), **out)
    print("wrote p4_%s_small" % kind)


def gen_ndiving(R, name="ndiving_small"):
    """gnn_model/gnn_model.py:61-86 run unmodified: same GNN classes as model/modules.py + output head + SelectiveNet."""
    import importlib
    saved_path = list(sys.path)
    sys.path[:] = [R.root] + [p for p in saved_path if not os.path.isdir(os.path.join(p or ".", "gnn_model")) and p != R.root]
    try:
        gm = importlib.import_module("gnn_model.gnn_model")
    finally:
        sys.path[:] = saved_path
    g = small_graph(seed=6, M=18, N=30, E=90)
    nd = gm.NeuralDiving(emb_size=D, gcn_mlp_layer_num=5)
    load_weights(nd, synth.neural_diving_weights(seed=8), strict=True)
    nd.eval()
    with torch.no_grad():
        outp, sel = nd(tens(g.cons), tens(g.edge_index), tens(g.edge_attr), tens(g.var))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), output=outp.numpy(), select=sel.numpy())
    print("wrote", name, outp.shape)


def gen_solenc(R, name="solenc_small", L=24):
    """CMSP.get_features(mip, x) (model/cmsp.py:35-42,65-69) -> SolutionEncoder.forward (model/encoder.py:32-36)."""
    g = small_graph()
    cm = R.CMSP(emb_num=3, emb_dim=D, n_heads=1, n_layers=2, padding_len=L)
    w = dict(synth.encoder_weights(seed=3))
    w.update(synth.solution_encoder_weights(n_layers=2, seed=12))
    load_weights(cm, w, strict=False)
    cm.eval()
    n = len(g.int_indices)
    sol = np.random.default_rng(13).integers(0, 2, size=n).astype(np.int64)
    mipobj = types.SimpleNamespace(constraint_features=tens(g.cons), edge_index=tens(g.edge_index), edge_attr=tens(g.edge_attr),
                                   variable_features=tens(g.var), int_indices=tens(g.int_indices), n_int_vars=n)
    feats, xf, kpm = cm.get_features(mipobj, tens(sol))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), x_features=xf.numpy(), kpm=kpm.numpy(), sol=sol)
    print("wrote", name, xf.shape)


def main(argv):
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(1)
    R = ref_loader.load()
    todo = argv or ["is_ckpt", "cf_step", "p4_sc_small", "p4_is_small", "ndiving", "solenc"]
    for nm in todo:
        t0 = time.time()
        if nm == "is_ckpt":
            gen_is_ckpt(R)
        elif nm == "cf_step":
            gen_cf_step(R)
        elif nm == "p4_sc_small":
            gen_p4(R, "sc")
        elif nm == "p4_is_small":
            gen_p4(R, "is")
        elif nm == "ndiving":
            gen_ndiving(R)
        elif nm == "solenc":
            gen_solenc(R)
        else:
            raise SystemExit("unknown fixture " + nm)
        print("%s: %.0f s" % (nm, time.time() - t0), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
