"""Shared by the CPU and GPU tests: seeded small cases (same generators as oracle/gen_golden.py),
torch views of the synthetic weights, error metrics."""
import os

import numpy as np
import torch

from dip_b200 import synth
from oracle.gen_golden import small_case, small_graph  # noqa: F401  (re-exported)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def tw(w, device="cpu", dtype=torch.float32):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(device=device, dtype=dtype) for k, v in w.items()}


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs(a, b):
    return float((torch.as_tensor(a).double().cpu() - torch.as_tensor(b).double().cpu()).abs().max())


def case_tensors(cs, device="cpu"):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    B = cs.B
    return {
        "mip": t(cs.mip), "kpm": t(cs.kpm), "x": t(cs.x), "b": t(cs.b), "c": t(cs.c),
        "mipB": t(cs.mip)[None].repeat(B, 1, 1), "kpmB": t(cs.kpm)[None].repeat(B, 1),
    }


def coalesced(cs):
    """(rows, cols, vals, M) torch tensors of the coalesced matrix (oracle path)."""
    from oracle import restate
    r, c, v = restate.coalesce_coo(cs.rows, cs.cols, cs.vals, cs.M, cs.n)
    return torch.from_numpy(r), torch.from_numpy(c), torch.from_numpy(v), cs.M


def sc_case(L=2000, n=2000, M=1000, density=0.05, B=1, seed=0):
    """Set-Cover shaped case at (optionally) full size with n <= L (padding at the tail)."""
    inst = synth.set_cover(M, n, density, seed=seed)
    mip = synth.mip_features(L, n, seed=seed + 10)
    kpm = np.zeros(L, dtype=bool)
    kpm[n:] = True
    x = synth.initial_noise(seed, range(B), L)
    return inst, mip, kpm, x
