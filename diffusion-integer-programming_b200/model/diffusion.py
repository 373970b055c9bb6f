"""DDPMTrainer / DDPMSampler / DDIMSampler with the reference's API
(/root/reference/model/diffusion.py:232-714), executed by the sm_100a engine.

What stays in Python: the S-step loop, the schedule tables (a few hundred host scalars, computed
with the same float32/float64 mix as the reference) and torch's random generator for the noise
(so a run seeded with torch.manual_seed draws the same latents as the reference on the same GPU).
Everything else -- denoiser, decoder forward + input-gradient backward, CSR guidance, fused update --
is one library call per step.  There is no autograd and no CPU path.
"""
from collections import OrderedDict

import numpy as np
import torch

from dip_b200 import _lib
from dip_b200.engine import Engine, default_engine
from model.decoder import _module_uid, same_rows, tensor_key, use_conditions, weights_signature
from model.modules import QuickGELU, ResidualAttentionBlock
from model.utils import (default, exists, extract_into_tensor, make_beta_schedule, make_ddim_sampling_parameters,
                         make_ddim_timesteps, noise_like, to_torch)


def get_clip_loss(mip_features, x_features):
    """Contrastive training loss (reference: diffusion.py:18-36): training only, out of scope."""
    raise NotImplementedError("get_clip_loss belongs to training; out of scope of the sampling path")


def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1, scale=1, max_period=10000):
    """Host version of the sinusoidal embedding (reference: diffusion.py:39-74).  The sampling path
    computes it inside dip_time_token; this function exists for API compatibility."""
    assert len(timesteps.shape) == 1, "Timesteps should be a 1d-array"
    half = embedding_dim // 2
    freq = torch.exp(-np.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / (half - downscale_freq_shift))
    ang = scale * timesteps[:, None].float() * freq[None, :]
    emb = torch.cat([torch.cos(ang), torch.sin(ang)] if flip_sin_to_cos else [torch.sin(ang), torch.cos(ang)], dim=-1)
    if embedding_dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1, 0, 0))
    return emb


class NoisePredictV2(torch.nn.Module):
    """x0 / eps predictor (reference: diffusion.py:120-180): time token + [mip | x] -> 2-layer MLP ->
    n_layers attention blocks over L+1 tokens -> last L tokens."""

    def __init__(self, attn_dim: int, n_heads: int, n_layers: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.attn_dim = attn_dim
        self.resblocks = torch.nn.ModuleList([ResidualAttentionBlock(attn_dim, n_heads, attn_mask) for _ in range(n_layers)])
        self.mlp_xt = torch.nn.Sequential(OrderedDict([
            ("linear_1", torch.nn.Linear(2 * attn_dim, attn_dim)), ("geru", QuickGELU()),
            ("linear_2", torch.nn.Linear(attn_dim, attn_dim)), ("geru", QuickGELU())]))
        self._uid = next(_module_uid)

    def sync_engine(self, engine):
        """Upload the weights unless the ENGINE already holds exactly these (the engine, not the module, remembers whose weights it
        has: two trainers sharing one engine re-upload when they alternate)."""
        for blk in self.resblocks:
            blk.check_supported()
        sig = (self._uid, weights_signature(self))
        if engine.den_sig != sig:
            engine.set_denoiser({"predict_model." + k: v.detach() for k, v in self.state_dict().items()})
            engine.den_sig = sig

    def forward(self, x_features, t, mip_features, key_padding_mask):
        if x_features.requires_grad:
            raise NotImplementedError("training / autograd through the denoiser is out of scope")
        if not bool((t == t[0]).all()):
            # per-sample timesteps (the loss evaluation of DDPMTrainer.forward): one engine call per distinct t
            out = torch.empty_like(x_features, dtype=torch.float32)
            for tv in torch.unique(t).tolist():
                idx = torch.nonzero(t == tv).flatten()
                out[idx] = self.forward(x_features[idx], t[idx], mip_features[idx], key_padding_mask[idx])
            return out
        eng = default_engine(x_features.device)
        self.sync_engine(eng)
        x = x_features.detach().to(torch.float32).contiguous()
        tt = float(t[0])
        if same_rows(mip_features) and same_rows(key_padding_mask):
            use_conditions(eng, mip_features[0], key_padding_mask[0])
            return eng.denoise(x, tt)
        outs = []
        for i in range(x.shape[0]):
            use_conditions(eng, mip_features[i], key_padding_mask[i])
            outs.append(eng.denoise(x[i:i + 1], tt))
        return torch.cat(outs)

    def apply_model(self, x_features, t, mip_features, key_padding_mask):
        with torch.no_grad():
            self.eval()
            return self(x_features, t, mip_features, key_padding_mask)


class DDPMTrainer(torch.nn.Module):
    """Owner of the denoiser weights and of the beta schedule buffers (reference: diffusion.py:232-273,
    334-335); same state_dict layout.  forward / q_sample / p_losses evaluate the training loss (diffusion.py:275-332)
    without gradients -- validation and monitoring; optimisation (dW kernels) is out of scope and nothing here is differentiable."""

    def __init__(self, attn_dim, n_heads, n_layers, device, timesteps=1000, loss_type="l2", beta_schedule="linear",
                 parameterization="x0"):
        super().__init__()
        self.device = device
        self.loss_type = loss_type
        self.parameterization = parameterization
        self.predict_model = NoisePredictV2(attn_dim, n_heads, n_layers).to(device)
        self.register_schedule(beta_schedule=beta_schedule, timesteps=timesteps)

    def register_schedule(self, beta_schedule="linear", timesteps=1000, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
        betas = make_beta_schedule(beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end, cosine_s=cosine_s)
        alphas = 1. - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1., ac[:-1])
        self.num_timesteps = int(timesteps)
        self.linear_start = linear_start
        self.linear_end = linear_end
        assert ac.shape[0] == self.num_timesteps, 'alphas have to be defined for each timestep'
        for name, val in (("alphas", alphas), ("betas", betas), ("alphas_cumprod", ac), ("alphas_cumprod_prev", ac_prev),
                          ("sqrt_alphas_cumprod", np.sqrt(ac)), ("sqrt_one_minus_alphas_cumprod", np.sqrt(1. - ac)),
                          ("log_one_minus_alphas_cumprod", np.log(1. - ac)), ("sqrt_recip_alphas_cumprod", np.sqrt(1. / ac)),
                          ("sqrt_recipm1_alphas_cumprod", np.sqrt(1. / ac - 1))):
            self.register_buffer(name, to_torch(val).to(self.device))

    @torch.no_grad()
    def forward(self, x, condition, key_padding_mask):
        """Loss evaluation at random timesteps (reference: diffusion.py:275-279); returns (pred_x_start, loss), no autograd graph."""
        t = torch.randint(0, self.num_timesteps, (x.shape[0],), device=x.device).long()
        return self.p_losses(x, t, condition, key_padding_mask)

    def q_sample(self, x_start, t, noise=None):
        """x_t = sqrt(abar_t) x_0 + sqrt(1 - abar_t) eps (reference: diffusion.py:281-284)."""
        noise = default(noise, lambda: torch.randn_like(x_start))
        return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start +
                extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    def predict_start_from_noise(self, x_t, t, noise):
        return (extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t -
                extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise)

    @torch.no_grad()
    def p_losses(self, x_start, t, condition, key_padding_mask, noise=None):
        """reference: diffusion.py:292-308 (note its 'eps' branch reconstructs x_0 from the TRUE noise, kept)."""
        noise = default(noise, lambda: torch.randn_like(x_start))
        x_noisy = self.q_sample(x_start=x_start, t=t, noise=noise)
        model_out = self.predict_model(x_noisy, t, condition, key_padding_mask)
        if self.parameterization == "eps":
            target, pred_x_start = noise, self.predict_start_from_noise(x_noisy, t, noise)
        elif self.parameterization == "x0":
            target, pred_x_start = x_start, model_out
        else:
            raise NotImplementedError("unknown parameterization '%s'" % self.parameterization)
        return pred_x_start, self.get_loss(model_out.squeeze(), target.squeeze(), key_padding_mask, mean=True)

    def get_loss(self, pred, target, key_padding_mask, mean=True):
        """reference: diffusion.py:310-332 (masked mean over the integer-variable tokens for 'l2')."""
        if self.loss_type == "l1":
            loss = (target - pred).abs()
            return loss.mean() if mean else loss
        if self.loss_type == "l2":
            loss = torch.nn.functional.mse_loss(target, pred, reduction="none")
            if not mean:
                return loss
            loss = loss.mean(dim=-1) * ~key_padding_mask
            return (loss.sum(dim=-1) / (~key_padding_mask).sum(dim=-1)).mean()
        raise NotImplementedError("unknown loss type '%s'" % self.loss_type)

    def get_predict_model(self):
        return self.predict_model


# --------------------------------------------------------------------------------------------------
class _SamplerBase(torch.nn.Module):
    """Wave management shared by both samplers: one engine instance per distinct (conditions, mask) row
    group; in the reference's scripts every row is the same instance (sample.py:151-152)."""

    def _engine(self, device=None):
        eng = default_engine(device)
        self.predict_model = self.model.predict_model
        self.predict_model.sync_engine(eng)
        return eng

    def _check(self, conditions, key_padding_mask):
        if not (torch.is_tensor(conditions) and conditions.is_cuda):
            raise _lib.DipError("conditions must live on a CUDA device: this sampler has no CPU path")
        if conditions.dim() != 3 or conditions.shape[-1] != 128:
            raise _lib.DipError("conditions must be (B, L, 128)")
        if key_padding_mask.shape != conditions.shape[:2]:
            raise _lib.DipError("key_padding_mask must be (B, L)")

    def _waves(self, conditions, key_padding_mask):
        if same_rows(conditions) and same_rows(key_padding_mask):
            return [slice(0, conditions.shape[0])]
        return [slice(i, i + 1) for i in range(conditions.shape[0])]

    def _prepare(self, eng, conditions, key_padding_mask, waves, A, b, c):
        """One prepared engine instance per wave, built BEFORE the step loop (host CSR/CSC sort + uploads happen here, never inside
        it) and cached by the identity of A / b / c / the wave's conditions and mask rows ('the CSR cache keyed by the A
        tensor identity'): a second call with the same tensors does no host work at all."""
        cache = self.__dict__.setdefault("_inst_cache", OrderedDict())
        coo, akey = None, None
        if A is not None:
            if A.layout != torch.sparse_coo:
                A = A.to_sparse()
            akey = (tensor_key(A._indices()), tensor_key(A._values()), tuple(A.shape))
        out = []
        for sl in waves:
            crow, mrow = conditions[sl.start], key_padding_mask[sl.start]
            key = (id(eng), akey, tensor_key(b), tensor_key(c), tensor_key(crow), tensor_key(mrow))
            prep = cache.get(key)
            if prep is None:
                if A is None:
                    prep = eng.prepare_conditions(crow, mrow)
                else:
                    if coo is None:
                        idx = A._indices().detach().cpu().numpy()
                        coo = (idx[0], idx[1], A._values().detach().to(torch.float32).cpu().numpy(), int(A.shape[0]))
                    prep = eng.prepare_instance(crow, mrow, coo[0], coo[1], coo[2], coo[3], b, c)
                # the entry keeps the keyed tensors alive: a freed tensor's address (and version 0) could otherwise be recycled by a
                # different tensor and hit this entry
                cache[key] = (prep, (A, b, c, conditions, key_padding_mask))
                while len(cache) > 8:
                    cache.popitem(last=False)
            else:
                cache.move_to_end(key)
                prep = prep[0]
            out.append((key, prep))
        return out

    def _use(self, eng, keyed):
        key, prep = keyed
        self.last_instance = prep            # the prepared device-side instance (CSR / CSC / masks) of the latest wave, for the caller's tail
        if getattr(eng, "_current_key", None) != key:
            eng.use_instance(prep)
            eng._current_key = key


class DDPMSampler(_SamplerBase):
    """1000-step ancestral sampler, guided and unguided (reference: diffusion.py:338-510)."""

    def __init__(self, trainer_model, decoder=None, gradient_scale=50000, obj_guided_coef=0, device="cpu"):
        super().__init__()
        self.device = device
        self.model = trainer_model
        self.parameterization = self.model.parameterization
        self.predict_model = self.model.get_predict_model()
        self.decoder = decoder
        self.gradient_scale = gradient_scale
        self.obj_guided_coef = obj_guided_coef
        self.v_posterior = 0
        self.original_elbo_weight = 0.
        self.register_schedule()

    def register_schedule(self):
        """Posterior tables (reference: diffusion.py:356-396), fp32 arithmetic on the trainer's fp32 buffers."""
        m = self.model
        betas, ac, acp, alphas = (t.detach().float().cpu() for t in (m.betas, m.alphas_cumprod, m.alphas_cumprod_prev, m.alphas))
        self.num_timesteps = int(betas.shape[0])
        self.linear_start, self.linear_end = m.linear_start, m.linear_end
        assert ac.shape[0] == self.num_timesteps, 'alphas have to be defined for each timestep'
        post_var = (1 - self.v_posterior) * betas * (1. - acp) / (1. - ac) + self.v_posterior * betas
        tables = {
            "betas": betas, "alphas_cumprod": ac, "alphas_cumprod_prev": acp,
            "sqrt_alphas_cumprod": to_torch(np.sqrt(ac.numpy())), "sqrt_one_minus_alphas_cumprod": to_torch(np.sqrt(1. - ac.numpy())),
            "log_one_minus_alphas_cumprod": to_torch(np.log(1. - ac.numpy())), "sqrt_recip_alphas_cumprod": to_torch(np.sqrt(1. / ac.numpy())),
            "sqrt_recipm1_alphas_cumprod": to_torch(np.sqrt(1. / ac.numpy() - 1)),
            "posterior_variance": post_var,
            "posterior_log_variance_clipped": to_torch(np.log(np.maximum(post_var.numpy(), 1e-20))),
            "posterior_mean_coef1": betas * to_torch(np.sqrt(acp.numpy())) / (1. - ac),
            "posterior_mean_coef2": (1. - acp) * to_torch(np.sqrt(alphas.numpy())) / (1. - ac),
        }
        for k, v in tables.items():
            self.register_buffer(k, v.float().to(self.device))
        self._host = {k: v.float().cpu() for k, v in tables.items()}

    def _coeffs(self, i):
        h = self._host
        std = (0.5 * h["posterior_log_variance_clipped"][i]).exp()
        return _lib.DdpmCoeffs(float(i), float(h["posterior_mean_coef1"][i]), float(h["posterior_mean_coef2"][i]), float(std),
                               0.0 if i == 0 else 1.0, float(h["sqrt_recip_alphas_cumprod"][i]), float(h["sqrt_recipm1_alphas_cumprod"][i]))

    def _loop(self, conditions, key_padding_mask, A, b, c):
        self._check(conditions, key_padding_mask)
        eng = self._engine(conditions.device)
        guided = A is not None
        if guided:
            if self.decoder is None:
                raise _lib.DipError("guided sampling needs a decoder")
            self.decoder.sync_engine(eng)
        self.conditions, self.batch_size, self.key_padding_mask = conditions, conditions.shape[0], key_padding_mask
        x = torch.randn_like(conditions).contiguous()
        eps = self.parameterization == "eps"
        waves = self._waves(conditions, key_padding_mask)
        prepared = self._prepare(eng, conditions, key_padding_mask, waves, A, b, c)
        for i in reversed(range(self.num_timesteps)):
            co = self._coeffs(i)
            noise = noise_like(x.shape, x.device)          # drawn every step, like the reference (:447, :458)
            for sl, prep in zip(waves, prepared):
                self._use(eng, prep)
                if guided:
                    eng.guided_ddpm_step(x[sl], co, self.gradient_scale, self.obj_guided_coef, eps, noise[sl])
                else:
                    eng.ddpm_step(x[sl], co, eps, noise[sl])
        return x

    @torch.no_grad()
    def sample(self, conditions, key_padding_mask):
        return self._loop(conditions, key_padding_mask, None, None, None)

    @torch.no_grad()
    def constraint_guided_sample(self, conditions, key_padding_mask, A, b, c):
        return self._loop(conditions, key_padding_mask, A, b, c)


class DDIMSampler(_SamplerBase):
    """S-step DDIM sampler, guided and unguided (reference: diffusion.py:512-699)."""

    def __init__(self, trainer_model, decoder=None, gradient_scale=1000, obj_guided_coef=0, device="cpu"):
        super().__init__()
        self.ddim_num_steps = None
        self.device = device
        self.model = trainer_model
        self.ddpm_num_timesteps = self.model.num_timesteps
        self.parameterization = self.model.parameterization
        self.decoder = decoder
        self.gradient_scale = gradient_scale
        self.initial_noise = None
        self.obj_guided_coef = obj_guided_coef
        self.v_posterior = 0
        self.original_elbo_weight = 0.
        self.schedule = "linear"
        self.keep_intermediates = True      # reference behaviour: S+1 latents returned (diffusion.py:587-591)
        self.strict_rng = True              # draw the (sigma = 0) per-step noise so torch's generator advances like the reference's

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=False):
        """DDIM tables (reference: diffusion.py:530-567): rebuilt on every call, host side."""
        self.ddim_timesteps = make_ddim_timesteps(ddim_discr_method=ddim_discretize, num_ddim_timesteps=ddim_num_steps,
                                                  num_ddpm_timesteps=self.ddpm_num_timesteps, verbose=verbose)
        ac = self.model.alphas_cumprod.detach().float().cpu()
        assert ac.shape[0] == self.ddpm_num_timesteps, 'alphas have to be defined for each timestep'
        if int(self.ddim_timesteps.max()) >= ac.shape[0]:
            raise IndexError("ddim timestep %d out of range (SURVEY section 7 quirk 10): choose S with max(range(0,T,T//S))+1 < T"
                             % int(self.ddim_timesteps.max()))
        sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(alphacums=ac, ddim_timesteps=self.ddim_timesteps, eta=ddim_eta,
                                                                    verbose=verbose)
        self.register_buffer('ddim_sigmas', to_torch(sigmas))
        self.register_buffer('ddim_alphas', to_torch(alphas))
        self.register_buffer('ddim_alphas_prev', torch.from_numpy(alphas_prev))
        self.register_buffer('ddim_sqrt_one_minus_alphas', to_torch(np.sqrt(1. - alphas.numpy())))

    def _coeffs(self, step, index):
        """fp32 scalars exactly as torch.full + the tensor arithmetic of diffusion.py:601-604,640-643 produce them."""
        a_t = self.ddim_alphas[index].float()
        a_prev = self.ddim_alphas_prev[index].float()
        sigma = self.ddim_sigmas[index].float()
        s1m = self.ddim_sqrt_one_minus_alphas[index].float()
        return Engine.ddim_coeffs(step, a_t.sqrt(), s1m, (1. - a_prev - sigma ** 2).sqrt(), a_prev.sqrt(), sigma)

    def _loop(self, conditions, key_padding_mask, S, A, b, c):
        self._check(conditions, key_padding_mask)
        self.make_schedule(ddim_num_steps=S)
        eng = self._engine(conditions.device)
        guided = A is not None
        if guided:
            if self.decoder is None:
                raise _lib.DipError("guided sampling needs a decoder")
            self.decoder.sync_engine(eng)
        self.batch_size, self.key_padding_mask, self.conditions = conditions.shape[0], key_padding_mask, conditions
        x = torch.randn_like(conditions)
        if self.initial_noise is not None:
            x = self.initial_noise            # cached across calls, like the reference (:583-586)
        self.initial_noise = x
        intermediates = [x]
        x = x.detach().to(torch.float32).clone().contiguous()
        eps = self.parameterization == "eps"
        waves = self._waves(conditions, key_padding_mask)
        prepared = self._prepare(eng, conditions, key_padding_mask, waves, A, b, c)
        for step in reversed(range(S)):
            index = S - step - 1
            co = self._coeffs(step, index)
            noise = None
            if co.sigma != 0.0 or self.strict_rng:
                noise = noise_like(x.shape, x.device)
            for sl, prep in zip(waves, prepared):
                self._use(eng, prep)
                nz = noise[sl] if (noise is not None and co.sigma != 0.0) else None
                if guided:
                    eng.guided_ddim_step(x[sl], co, self.gradient_scale, self.obj_guided_coef, eps, nz)
                else:
                    eng.ddim_step(x[sl], co, eps, nz)
            if self.keep_intermediates:
                intermediates.append(x.clone())
        if not self.keep_intermediates:
            intermediates.append(x)
        return x, intermediates

    @torch.no_grad()
    def constraint_guided_sample(self, conditions, key_padding_mask, A, b, c, S=100):
        return self._loop(conditions, key_padding_mask, S, A, b, c)

    @torch.no_grad()
    def sample(self, conditions, key_padding_mask, S=100):
        return self._loop(conditions, key_padding_mask, S, None, None, None)
