"""Driver for ncu: the instance encoder (K3), a few guided DDIM steps (K1, K2, K4) and the final decode of one Set-Cover wave at
the bench shapes (1000 x 2000, nnz 1e5, wave 37) -- the same launches bench.py times, without its repetition.

    python tools/prof_step.py [--steps 2] [--wave 37] [--config sc|is|cf]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-integer-programming_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from dip_b200 import synth, schedule, encoder as enc, ops  # noqa: E402
from dip_b200.engine import Engine, randn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--wave", type=int, default=37)
    ap.add_argument("--config", default="sc")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    if a.config == "sc":
        inst, gs, lam, n_den = synth.set_cover(1000, 2000, 0.05, seed=0), 1e5, 0.9, 1
    elif a.config == "is":
        inst, gs, lam, n_den = synth.independent_set(1500, 4860, 1500, seed=0), 2e4, 0.5, 2
    else:
        inst, gs, lam, n_den = synth.facility_location(100, 50, seed=0), 1e3, 0.7, 1
    L = inst.n
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    to_dev = lambda w: {k: T(v) for k, v in w.items()}
    eng = Engine()
    eng.set_denoiser(to_dev(synth.denoiser_weights(n_den, seed=1)))
    eng.set_decoder(to_dev(synth.decoder_weights(2, seed=2)))
    Wenc = {k[len("mip_encoder.gnn_model."):]: v for k, v in to_dev(synth.encoder_weights(seed=3)).items()}
    v = enc.gnn_forward(Wenc, T(inst.constraint_features), T(inst.edge_index), T(inst.edge_attr), T(inst.variable_features))
    mip = ops.gather_pad(v, T(inst.int_indices), L)
    eng.set_instance(mip, T(np.zeros(L, dtype=bool)), inst.rows, inst.cols, inst.a_val, inst.M, inst.b, inst.c,
                     a_int=inst.a_int, b_int=inst.b_int, c_int=inst.c_int)
    x = randn((a.wave, L, 128), seed=1, subseq=0, device=dev)
    out = eng.sample_guided_ddim(x, schedule.ddim_coeffs(100)[:a.steps], gs, lam)
    torch.cuda.synchronize()
    print("ok feasible", int((out["viol_int"] == 0).sum()), "of", a.wave)


if __name__ == "__main__":
    main()
