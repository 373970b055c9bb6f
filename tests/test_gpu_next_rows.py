"""GPU: the SURVEY section-8f rows -- device-side partial-solution hand-off, the Neural-Diving head on the shared K3 encoder, the
solution encoder forward -- against the reference (golden fixtures of the unmodified modules) and the reference's hand-off loop."""
import random
import types

import numpy as np
import pytest
import torch

from dip_b200 import synth, handoff
from helpers import golden, tw, rel_l2, small_graph

pytestmark = pytest.mark.gpu


def test_partial_solution_selection_on_device_matches_reference_loop(cuda):
    """sample_partialsol.py:171-199: per sample `num_vars = int(n * coverage)` variables get the decoded value (sample 0's values for
    every sample -- the reference's quirk -- unless switched off).  The device selection must (a) hand the host loop exactly what the
    reference's loop builds when it is given the same index sets, (b) choose exactly num_vars variables, uniformly, as a pure function
    of (seed, instance, sample) independent of how samples are batched."""
    B, n, cov = 96, 203, 0.2
    rng = np.random.default_rng(0)
    xbits = torch.from_numpy(rng.integers(0, 2, size=(B, n)).astype(np.uint8)).to(cuda)
    names = ["x%d" % i for i in range(n)]
    for quirk in (True, False):
        vals, mask = handoff.select_gpu(xbits, cov, seed=5, instance_id=17, reference_quirk=quirk)
        dev_sols = handoff.unpack(vals.cpu().numpy(), mask.cpu().numpy(), n)
        sel = [sorted(d.keys()) for d in dev_sols]
        assert all(len(s) == int(n * cov) for s in sel)

        class Replay(random.Random):            # feeds the reference loop the device's index sets
            def __init__(self):
                super().__init__(0)
                self.j = 0

            def sample(self, population, k):
                out = sel[self.j]
                assert k == len(out)
                self.j += 1
                return out

        ref = handoff.partial_solutions(xbits, names, cov, rng=Replay(), reference_quirk=quirk)
        assert ref == [{names[k]: v for k, v in d.items()} for d in dev_sols]
        # .sol export of one sample round-trips
        txt = handoff.to_sol_text(names, vals[3].cpu().numpy(), mask[3].cpu().numpy())
        parsed = {ln.split()[0]: int(ln.split()[1]) for ln in txt.splitlines()[1:]}
        assert parsed == ref[3]
    # sharding invariance: samples 40..59 selected alone equal rows 40..59 of the full batch
    v2, m2 = handoff.select_gpu(xbits[40:60].contiguous(), cov, seed=5, instance_id=17, sample_base=40, reference_quirk=False)
    vals, mask = handoff.select_gpu(xbits, cov, seed=5, instance_id=17, reference_quirk=False)
    assert torch.equal(m2, mask[40:60]) and torch.equal(v2, vals[40:60])
    # uniformity: every variable is chosen with probability k/n (chi-square over 96 x 40 draws, 203 cells)
    counts = np.zeros(n)
    for inst in range(40):
        _, m = handoff.select_gpu(xbits, cov, seed=9, instance_id=inst)
        counts += np.unpackbits(m.cpu().numpy(), axis=1)[:, :n].sum(0)
    expect = 40 * B * int(n * cov) / n
    chi2 = float(((counts - expect) ** 2 / expect).sum())
    assert chi2 < 203 + 5 * np.sqrt(2 * 203), chi2
    # degenerate coverages
    _, m0 = handoff.select_gpu(xbits, 0.0)
    _, m1 = handoff.select_gpu(xbits, 1.0)
    assert int(m0.sum()) == 0 and int(np.unpackbits(m1.cpu().numpy(), axis=1)[:, :n].sum()) == B * n


def test_neural_diving_head_vs_reference(cuda):
    """gnn_model/gnn_model.py:61-86 (5 GCN layers + output head + SelectiveNet) on the K3 kernels vs the unmodified reference."""
    from gnn_model.gnn_model import NeuralDiving
    g = golden("ndiving_small")
    gr = small_graph(seed=6, M=18, N=30, E=90)
    nd = NeuralDiving(emb_size=128, gcn_mlp_layer_num=5).to(cuda)
    missing, unexpected = nd.load_state_dict(tw(synth.neural_diving_weights(seed=8)), strict=True)
    assert not missing and not unexpected
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    out, sel = nd(T(gr.cons), T(gr.edge_index), T(gr.edge_attr), T(gr.var))
    assert out.shape == (30,) and sel.shape == (30,)
    assert rel_l2(out, g["output"]) < 1e-4 and rel_l2(sel, g["select"]) < 1e-4


def test_solution_encoder_forward_vs_reference(cuda):
    """CMSP.get_features(mip, x) (model/cmsp.py:35-42,65-69; model/encoder.py:32-36): token-2 padding, embedding, doubly-residual blocks."""
    from model.cmsp import CMSP
    g = golden("solenc_small")
    gr = small_graph()
    cm = CMSP(emb_num=3, emb_dim=128, n_heads=1, n_layers=2, padding_len=24).to(cuda)
    w = dict(synth.encoder_weights(seed=3))
    w.update(synth.solution_encoder_weights(n_layers=2, seed=12))
    missing, unexpected = cm.load_state_dict(tw(w), strict=False)
    assert not unexpected and missing == ["logit_scale"]
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    mip = types.SimpleNamespace(constraint_features=T(gr.cons), edge_index=T(gr.edge_index), edge_attr=T(gr.edge_attr),
                                variable_features=T(gr.var), int_indices=T(gr.int_indices), n_int_vars=len(gr.int_indices))
    feats, xf, kpm = cm.get_features(mip, T(g["sol"]))
    n = len(gr.int_indices)
    assert xf.shape == (1, 24, 128) and np.array_equal(kpm.cpu().numpy(), g["kpm"])
    assert rel_l2(xf[:, :n], g["x_features"][:, :n]) < 1e-4
    assert rel_l2(feats, golden("encoder_small")["mip_features"]) < 1e-4
