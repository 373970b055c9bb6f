"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, through oracle/shims.py) on small seeded cases.  Run in the build container:

    python -m oracle.gen_golden

The reference has no golden vectors of its own; these files are the pins.  Inputs are
regenerated from seeds by ``dip_b200.synth`` (numpy) at test time; the fixtures store the seeds'
checksums and the reference's outputs.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-integer-programming_b200"))

from oracle import ref_loader  # noqa: E402
from dip_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
D = 128


def small_case(seed=0, L=24, n=20, M=12, nnz=44, B=2, dup=4):
    """A tiny instance in the reference's tensor convention; a few duplicate COO entries exercise coalesce()."""
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, M, size=nnz)
    cols = rng.integers(0, n, size=nnz)
    if dup > 0:
        rows[-dup:] = rows[:dup]
        cols[-dup:] = cols[:dup]
    vals = -np.ones(nnz, dtype=np.float32)
    vals[::5] = 2.0
    b = (-rng.random(M)).astype(np.float32)
    b[0] = 0.0
    c = rng.random(n).astype(np.float32)
    mip = synth.mip_features(L, n, seed=seed + 10)
    kpm = np.zeros(L, dtype=bool)
    kpm[n:] = True
    x = np.random.default_rng(seed + 20).standard_normal((B, L, D)).astype(np.float32)
    return types.SimpleNamespace(L=L, n=n, M=M, B=B, rows=rows, cols=cols, vals=vals, b=b, c=c, mip=mip, kpm=kpm, x=x)


def tens(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def load_weights(module, w, strict=True):
    sd = {k: tens(v) for k, v in w.items()}
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    if strict:
        assert not missing, missing
    return module


class Recorder:
    """Wraps the decoder module: remembers the x tensor autograd differentiates (diffusion.py:619-621)."""

    def __init__(self, decoder):
        self.decoder = decoder
        self.x_t = None
        self.out = None

    def __call__(self, conditions, x_t, kpm):
        self.x_t = x_t
        self.out = self.decoder(conditions, x_t, kpm)
        return self.out

    def parameters(self):
        return self.decoder.parameters()


def build_reference(R, n_den, n_dec, param, gs, lam, wseed=1):
    tr = R.DDPMTrainer(D, 1, n_den, "cpu", parameterization=param)
    load_weights(tr, synth.denoiser_weights(n_den, seed=wseed), strict=False)
    dec = R.MipConditionalDecorder(D, 1, n_dec)
    load_weights(dec, synth.decoder_weights(n_dec, seed=wseed + 1))
    tr.eval(); dec.eval()
    return tr, dec


def gen_ddim(R, name, param="x0", gs=1e5, lam=0.9, S=100, indices=(0, 1, 50, 99), free_steps=3):
    cs = small_case()
    tr, dec = build_reference(R, 1, 2, param, gs, lam)
    rec = Recorder(dec)
    s = R.DDIMSampler(tr, rec, gradient_scale=gs, obj_guided_coef=lam, device="cpu")
    mip = tens(cs.mip)[None].repeat(cs.B, 1, 1)
    kpm = tens(cs.kpm)[None].repeat(cs.B, 1)
    A = torch.sparse_coo_tensor(tens(np.stack([cs.rows, cs.cols])), tens(cs.vals), size=(cs.M, cs.n))
    b, c = tens(cs.b), tens(cs.c)
    s.make_schedule(S)
    s.predict_model, s.batch_size, s.key_padding_mask, s.conditions = tr.predict_model, cs.B, kpm, mip
    out = {"S": S, "gs": gs, "lam": lam, "indices": np.array(indices), "x_checksum": float(np.abs(cs.x).sum()),
           "w_checksum": float(sum(np.abs(v).sum() for v in synth.decoder_weights(2, seed=2).values()))}
    # teacher-forced single steps from the same input latent at several schedule positions
    for index in indices:
        step = S - 1 - index
        x = tens(cs.x).clone()
        x_prev, x0 = s.consraint_guided_p_sample_ddim(x, step, index, A, b, c)
        out["tf%d_x0" % index] = x0.detach().numpy()
        out["tf%d_grad" % index] = rec.x_t.grad.detach().numpy()
        out["tf%d_xprev" % index] = x_prev.detach().numpy()
        out["tf%d_p" % index] = rec.out[0].detach().view(cs.B, cs.n).numpy()
    # free-running first steps
    x = tens(cs.x).clone()
    for k in range(free_steps):
        step = S - 1 - k
        x, _ = s.consraint_guided_p_sample_ddim(x, step, k, A, b, c)
        out["free%d_x" % k] = x.detach().numpy()
    # unguided step
    with torch.no_grad():
        xu, x0u = s.p_sample_ddim(tens(cs.x).clone(), S - 1, mip, 0)
    out["unguided_xprev"] = xu.numpy()
    # final decode + round + penalty on the free-running latent (sample_partialsol.py:164-168)
    import torch_sparse
    with torch.no_grad():
        pred, _ = dec(mip, x, kpm)
        xr = torch.round(pred).view(cs.B, -1, 1)
        Ac = A.coalesce()
        pen = torch.max(torch_sparse.spmm(Ac.indices(), Ac.values(), A.size()[0], A.size()[1], xr).squeeze() - b,
                        torch.tensor(0)).sum(axis=-1)
    out["final_p"] = pred.view(cs.B, cs.n).numpy()
    out["final_x"] = xr.view(cs.B, cs.n).numpy()
    out["final_penalty"] = pen.numpy()
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("wrote", name, {k: getattr(v, "shape", v) for k, v in out.items() if k.startswith("tf0") or k.startswith("final")})


def gen_ddpm(R, name, gs=1.5e4, lam=0.1, ts=(999, 500, 0)):
    cs = small_case(seed=1)
    tr, dec = build_reference(R, 1, 2, "x0", gs, lam)
    rec = Recorder(dec)
    s = R.DDPMSampler(tr, rec, gradient_scale=gs, obj_guided_coef=lam, device="cpu")
    mip = tens(cs.mip)[None].repeat(cs.B, 1, 1)
    kpm = tens(cs.kpm)[None].repeat(cs.B, 1)
    A = torch.sparse_coo_tensor(tens(np.stack([cs.rows, cs.cols])), tens(cs.vals), size=(cs.M, cs.n))
    b, c = tens(cs.b), tens(cs.c)
    s.conditions, s.batch_size, s.key_padding_mask = mip, cs.B, kpm
    out = {"gs": gs, "lam": lam, "ts": np.array(ts)}
    captured = {}
    orig = R.diffusion.noise_like

    def cap(shape, device, repeat=False):
        captured["noise"] = orig(shape, device, repeat)
        return captured["noise"]

    R.diffusion.noise_like = cap
    try:
        for t in ts:
            torch.manual_seed(100 + t)
            x = tens(cs.x).clone()
            xp = s.constraint_guided_p_sample(x, torch.full((cs.B,), t, dtype=torch.long), A, b, c)
            out["t%d_xprev" % t] = xp.detach().numpy()
            out["t%d_noise" % t] = captured["noise"].numpy()
            out["t%d_grad" % t] = rec.x_t.grad.detach().numpy()
            out["t%d_mean" % t] = rec.x_t.detach().numpy()
            with torch.no_grad():
                torch.manual_seed(200 + t)
                xu = s.p_sample(tens(cs.x).clone(), torch.full((cs.B,), t, dtype=torch.long))
            out["t%d_unguided" % t] = xu.numpy()
            out["t%d_unoise" % t] = captured["noise"].numpy()
    finally:
        R.diffusion.noise_like = orig
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("wrote", name)


def small_graph(seed=5, M=14, N=22, E=60):
    rng = np.random.default_rng(seed)
    pairs = set()
    while len(pairs) < E:
        pairs.add((int(rng.integers(M)), int(rng.integers(N))))
    pr = np.array(sorted(pairs), dtype=np.int64)
    perm = rng.permutation(E)                       # edges NOT sorted: exercises the CSR/CSC builders
    ei = np.stack([pr[perm, 0], pr[perm, 1]])
    ea = rng.standard_normal((E, 1)).astype(np.float32)
    cons = rng.standard_normal((M, 5)).astype(np.float32)
    var = rng.standard_normal((N, 17)).astype(np.float32)
    int_idx = np.sort(rng.choice(N, size=N - 3, replace=False)).astype(np.int64)
    return types.SimpleNamespace(M=M, N=N, E=E, edge_index=ei, edge_attr=ea, cons=cons, var=var, int_indices=int_idx)


def gen_encoder(R, name, L=24):
    g = small_graph()
    cm = R.CMSP(emb_num=3, emb_dim=D, n_heads=1, n_layers=1, padding_len=L)
    load_weights(cm, synth.encoder_weights(seed=3), strict=False)
    cm.eval()
    mipobj = types.SimpleNamespace(constraint_features=tens(g.cons), edge_index=tens(g.edge_index), edge_attr=tens(g.edge_attr),
                                   variable_features=tens(g.var), int_indices=tens(g.int_indices), n_int_vars=len(g.int_indices))
    feats, _, kpm = cm.get_features(mipobj)
    with torch.no_grad():
        raw = cm.mip_encoder(mipobj.constraint_features, mipobj.edge_index, mipobj.edge_attr, mipobj.variable_features)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), mip_features=feats.numpy(), kpm=kpm.numpy(), var_features=raw.numpy())
    print("wrote", name, feats.shape)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(1)
    R = ref_loader.load()
    gen_ddim(R, "ddim_guided_x0")
    gen_ddim(R, "ddim_guided_eps", param="eps", gs=2e4, lam=0.5, indices=(0, 50), free_steps=1)
    gen_ddpm(R, "ddpm_guided_x0")
    gen_encoder(R, "encoder_small")


if __name__ == "__main__":
    main()
