"""Schedule / padding helpers with the reference's names (/root/reference/model/utils.py:97-217).

Host-side numpy: these tables are a few hundred scalars computed once per sampler call; their
values (including the float32/float64 mix the reference produces) feed the fused update kernel as
plain fp32 coefficients.
"""
import numpy as np
import torch


def exists(x):
    return x is not None


def default(val, d):
    if val is not None:
        return val
    return d() if callable(d) else d


def to_torch(x):
    """float32 tensor from a tensor or numpy array (reference: model/utils.py:146-151)."""
    if torch.is_tensor(x):
        return x.clone().detach().float()
    return torch.from_numpy(np.asarray(x)).float()


def extract_into_tensor(a, t, x_shape):
    """a[t] reshaped to broadcast over x (reference: model/utils.py:135-144)."""
    out = a.gather(-1, t)
    return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """float64 numpy betas (reference: model/utils.py:154-180)."""
    if schedule == "linear":
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    elif schedule == "cosine":
        steps = torch.arange(n_timestep + 1, dtype=torch.float64) / n_timestep + cosine_s
        ac = torch.cos(steps / (1 + cosine_s) * np.pi / 2).pow(2)
        ac = ac / ac[0]
        betas = torch.clamp(1 - ac[1:] / ac[:-1], min=0, max=0.999)
    elif schedule == "sqrt_linear":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64)
    elif schedule == "sqrt":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas.numpy()


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    """Selected DDPM indices, each +1 (reference: model/utils.py:189-203)."""
    if ddim_discr_method == "uniform":
        stride = num_ddpm_timesteps // num_ddim_timesteps
        picked = np.arange(0, num_ddpm_timesteps, stride)
    elif ddim_discr_method == "quad":
        picked = (np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    steps_out = picked + 1
    if verbose:
        print(f"Selected timesteps for ddim sampler: {steps_out}")
    return steps_out


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """(sigmas, alphas, alphas_prev) (reference: model/utils.py:206-217).  ``alphacums`` is the float32
    CPU tensor of cumulative alphas; alphas stays float32, alphas_prev is float64 holding f32 values."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    if verbose:
        print(f"Selected alphas for ddim sampler: a_t: {alphas}; a_(t-1): {alphas_prev}")
        print(f"For the chosen value of eta, which is {eta}, this results in the following sigma_t schedule "
              f"for ddim sampler {sigmas}")
    return sigmas, alphas, alphas_prev


def noise_like(shape, device, repeat=False):
    """N(0,1) from torch's generator on `device` (reference: model/utils.py:183-186)."""
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)


def get_padding(features, n_int_vars, padding_len, pad_object):
    """Split a concatenated batch and right-pad every piece to ``padding_len``; returns the padded
    batch and the key-padding mask, True on padding (reference: model/utils.py:97-114)."""
    sizes = [n_int_vars] if isinstance(n_int_vars, int) else n_int_vars.tolist()
    pieces = torch.split(features, sizes)
    if pad_object not in ("mip", "solution"):
        raise Exception("Padding object not found!")
    padded, masks = [], []
    for v in pieces:
        extra = padding_len - v.shape[0]
        if pad_object == "mip":
            padded.append(torch.nn.functional.pad(v, (0, 0, 0, extra), "constant", 0).unsqueeze(0))
        else:
            padded.append(torch.nn.functional.pad(v, (0, extra), "constant", 2).unsqueeze(0))
        m = torch.zeros(padding_len, dtype=torch.bool, device=v.device)
        m[v.shape[0]:] = True
        masks.append(m.unsqueeze(0))
    return torch.concat(padded, dim=0), torch.concat(masks, dim=0)
