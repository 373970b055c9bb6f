"""Host-side schedule scalars for engine users that do not go through model.diffusion (bench.py, the
multi-GPU driver).  Same numbers as DDIMSampler.make_schedule / DDPMSampler.register_schedule produce:
float64 numpy betas -> float32 cumulative products -> per-step fp32 coefficients."""
import numpy as np
import torch

from . import _lib
from model.utils import make_beta_schedule, make_ddim_sampling_parameters, make_ddim_timesteps, to_torch


def alphas_cumprod(timesteps=1000, beta_schedule="linear"):
    betas = make_beta_schedule(beta_schedule, timesteps)
    return to_torch(np.cumprod(1. - betas, axis=0))


def ddim_coeffs(S, timesteps=1000, eta=0.0):
    """DdimCoeffs for the S steps in execution order (index 0..S-1, denoiser t = S-1-index)."""
    ac = alphas_cumprod(timesteps)
    ts = make_ddim_timesteps("uniform", S, timesteps, verbose=False)
    if int(ts.max()) >= timesteps:
        raise IndexError("ddim timestep out of range for S=%d" % S)
    sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(ac, ts, eta, verbose=False)
    sigmas, a_prev32 = to_torch(sigmas), torch.from_numpy(alphas_prev).float()
    s1m = to_torch(np.sqrt(1. - alphas.numpy()))
    out = []
    for i in range(S):
        a_t, a_p, sg = alphas[i].float(), a_prev32[i], sigmas[i]
        out.append(_lib.DdimCoeffs(float(S - 1 - i), float(a_t.sqrt()), float(s1m[i]), float((1. - a_p - sg ** 2).sqrt()), float(a_p.sqrt()), float(sg)))
    return out
