"""GPU: the drop-in ``model.*`` classes (same API as the reference) against the CPU oracle."""
import types

import numpy as np
import pytest
import torch

from dip_b200 import synth
from helpers import tw, rel_l2, small_case, small_graph, coalesced, case_tensors

pytestmark = pytest.mark.gpu


def build_models(cuda, n_den=1, n_dec=2, param="x0"):
    from model.diffusion import DDPMTrainer
    from model.decoder import MipConditionalDecorder
    tr = DDPMTrainer(attn_dim=128, n_heads=1, n_layers=n_den, device=cuda, parameterization=param)
    missing, unexpected = tr.load_state_dict(tw(synth.denoiser_weights(n_den, seed=1)), strict=False)
    assert not unexpected and all(not k.startswith("predict_model") for k in missing)
    dec = MipConditionalDecorder(attn_dim=128, n_heads=1, n_layers=n_dec).to(cuda)
    dec.load_state_dict(tw(synth.decoder_weights(n_dec, seed=2)))
    return tr.eval(), dec.eval()


def test_ddim_sampler_dropin(cuda):
    from model.diffusion import DDIMSampler
    from oracle import restate
    cs = small_case(B=3)
    tr, dec = build_models(cuda)
    t = case_tensors(cs, cuda)
    A = torch.sparse_coo_tensor(torch.from_numpy(np.stack([cs.rows, cs.cols])), torch.from_numpy(cs.vals), size=(cs.M, cs.n)).to(cuda)
    s = DDIMSampler(trainer_model=tr, decoder=dec, gradient_scale=1e5, obj_guided_coef=0.9, device=cuda)
    s.initial_noise = t["x"].clone()                       # the reference's noise hook (diffusion.py:584-586)
    S = 5
    x, inter = s.constraint_guided_sample(t["mipB"], t["kpmB"], A, t["b"], t["c"], S=S)
    assert len(inter) == S + 1 and torch.equal(inter[0], t["x"]) and torch.equal(inter[-1], x)
    tc = case_tensors(cs)
    xo, recs = restate.guided_ddim_sample(tw(synth.denoiser_weights(1, seed=1)), tw(synth.decoder_weights(2, seed=2)), tc["x"], S,
                                          restate.trainer_buffers()["alphas_cumprod"], tc["mipB"], tc["kpmB"], coalesced(cs), tc["b"], tc["c"],
                                          1e5, 0.9, record=True)
    for k in range(S):
        assert rel_l2(inter[k + 1], recs[k]["x_prev"]) < 2e-3 * (k + 1), k
    # initial_noise is cached and reused by later calls, like the reference
    x2, _ = s.constraint_guided_sample(t["mipB"], t["kpmB"], A, t["b"], t["c"], S=S)
    assert torch.equal(x, x2)
    # unguided sampling and the decoder's own forward
    xu, inter_u = s.sample(t["mipB"], t["kpmB"], S=S)
    assert len(inter_u) == S + 1 and torch.isfinite(xu).all()
    sols, selected = dec(t["mipB"], x, t["kpmB"])
    assert selected is None and sols.shape == (cs.B * cs.n,)
    fd = restate.final_decode(tw(synth.decoder_weights(2, seed=2)), tc["mipB"], x.cpu(), tc["kpmB"], coalesced(cs), tc["b"])
    assert rel_l2(sols.view(cs.B, cs.n), fd["p"]) < 1e-4


def test_torch_seed_gives_reference_initial_noise(cuda):
    """Seeded like the reference (torch.manual_seed) the sampler starts from torch.randn_like(conditions)."""
    from model.diffusion import DDIMSampler
    cs = small_case(B=2)
    tr, dec = build_models(cuda)
    t = case_tensors(cs, cuda)
    s = DDIMSampler(tr, dec, 1e5, 0.9, cuda)
    torch.manual_seed(123)
    _, inter = s.sample(t["mipB"], t["kpmB"], S=2)
    torch.manual_seed(123)
    assert torch.equal(inter[0], torch.randn_like(t["mipB"]))


def test_ddpm_sampler_dropin_short_chain(cuda):
    from model.diffusion import DDPMSampler, DDPMTrainer
    from model.decoder import MipConditionalDecorder
    cs = small_case(B=2)
    tr = DDPMTrainer(128, 1, 1, cuda, timesteps=4)           # 4-step chain keeps the oracle comparison cheap
    tr.load_state_dict(tw(synth.denoiser_weights(1, seed=1)), strict=False)
    dec = MipConditionalDecorder(128, 1, 2).to(cuda)
    dec.load_state_dict(tw(synth.decoder_weights(2, seed=2)))
    t = case_tensors(cs, cuda)
    A = torch.sparse_coo_tensor(torch.from_numpy(np.stack([cs.rows, cs.cols])), torch.from_numpy(cs.vals), size=(cs.M, cs.n)).to(cuda)
    s = DDPMSampler(tr, dec, gradient_scale=1.5e4, obj_guided_coef=0.1, device=cuda)
    torch.manual_seed(5)
    x = s.constraint_guided_sample(t["mipB"], t["kpmB"], A, t["b"], t["c"])
    torch.manual_seed(5)
    x2 = s.constraint_guided_sample(t["mipB"], t["kpmB"], A, t["b"], t["c"])
    assert x.shape == t["x"].shape and torch.isfinite(x).all() and torch.equal(x, x2)
    torch.manual_seed(5)
    xu = s.sample(t["mipB"], t["kpmB"])
    assert torch.isfinite(xu).all()


def test_cmsp_get_features(cuda):
    from model.cmsp import CMSP
    from helpers import golden
    g = golden("encoder_small")
    gr = small_graph()
    cm = CMSP(emb_num=3, emb_dim=128, n_heads=1, n_layers=1, padding_len=24).to(cuda)
    missing, unexpected = cm.load_state_dict(tw(synth.encoder_weights(seed=3)), strict=False)
    assert not unexpected and all(k.startswith("sol_encoder") or k == "logit_scale" for k in missing)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    mip = types.SimpleNamespace(constraint_features=T(gr.cons), edge_index=T(gr.edge_index), edge_attr=T(gr.edge_attr),
                                variable_features=T(gr.var), int_indices=T(gr.int_indices), n_int_vars=len(gr.int_indices))
    feats, xf, kpm = cm.get_features(mip)
    assert xf is None and feats.shape == (1, 24, 128) and kpm.dtype == torch.bool
    assert rel_l2(feats, g["mip_features"]) < 1e-4 and np.array_equal(kpm.cpu().numpy(), g["kpm"])


def test_no_cpu_path():
    """Product classes refuse CPU tensors instead of falling back."""
    import model.diffusion as md
    from dip_b200 import _lib
    tr = md.DDPMTrainer(128, 1, 1, "cpu")
    s = md.DDIMSampler(tr, None, 1e3, 0.0, "cpu")
    with pytest.raises(_lib.DipError):
        s.sample(torch.zeros(1, 4, 128), torch.zeros(1, 4, dtype=torch.bool), S=5)


@pytest.mark.parametrize("param", ["x0", "eps"])
def test_trainer_loss_evaluation_matches_oracle(cuda, param):
    """DDPMTrainer.p_losses / forward (reference: diffusion.py:275-332) as a forward-only loss evaluation with per-sample
    timesteps: q_sample, the denoiser at each sample's own t (one engine call per distinct t), the masked l2 loss."""
    from model.diffusion import DDPMTrainer
    from oracle import restate
    cs = small_case(B=4)
    Wd = synth.denoiser_weights(1, seed=1)
    tr = DDPMTrainer(128, 1, 1, cuda, parameterization=param)
    tr.load_state_dict(tw(Wd), strict=False)
    t = case_tensors(cs, cuda)
    x0 = t["x"]
    ts = torch.tensor([3, 977, 3, 500], device=cuda)
    noise = torch.from_numpy(np.random.default_rng(4).standard_normal(tuple(x0.shape)).astype(np.float32)).to(cuda)
    pred, loss = tr.p_losses(x0, ts, t["mipB"], t["kpmB"], noise=noise)
    # oracle: same arithmetic on the CPU, sample by sample
    buf = restate.trainer_buffers()
    tc_ = case_tensors(cs)
    x_noisy = torch.stack([float(np.sqrt(buf["alphas_cumprod"][k])) * tc_["x"][i] + float(np.sqrt(1.0 - buf["alphas_cumprod"][k])) * noise[i].cpu()
                           for i, k in enumerate(ts.tolist())])
    out = torch.cat([restate.denoiser_fwd(tw(Wd), x_noisy[i:i + 1], torch.tensor([k]), tc_["mipB"][i:i + 1], tc_["kpmB"][i:i + 1]) for i, k in enumerate(ts.tolist())])
    target = tc_["x"] if param == "x0" else noise.cpu()
    per_tok = ((target - out) ** 2).mean(-1) * ~tc_["kpmB"]
    ref_loss = (per_tok.sum(-1) / (~tc_["kpmB"]).sum(-1)).mean()
    assert abs(float(loss) - float(ref_loss)) < 1e-3 * abs(float(ref_loss))
    if param == "x0":
        assert rel_l2(pred, out) < 1e-3
    torch.manual_seed(0)
    p2, l2 = tr(x0, t["mipB"], t["kpmB"])                       # random timesteps: finite, right shapes, no autograd graph
    assert p2.shape == x0.shape and torch.isfinite(l2) and not l2.requires_grad
