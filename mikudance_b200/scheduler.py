"""DDIMScheduler — host-side tables of the DDIM sampler the reference configures
(configs/inference/mikudance_config.yaml:24-33; `diffusers.DDIMScheduler` in the reference, built at
scripts/inference_video.py:101-102).  Duck-types what the pipelines call: set_timesteps, timesteps,
init_noise_sigma, scale_model_input, step(...).prev_sample, order.  The tables (betas, cumulative
alphas, timestep lists) are tiny host computations exactly as in the reference; the per-step tensor
update runs in the fused CFG+DDIM CUDA kernel (mdk_cfg_ddim_step) when driven by the pipeline, or
through `step()` on CUDA tensors for callers that use the scheduler directly.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


def _rescale_zero_terminal_snr(betas: torch.Tensor) -> torch.Tensor:
    alphas_bar_sqrt = torch.cumprod(1.0 - betas, dim=0).sqrt()
    a0, aT = alphas_bar_sqrt[0].clone(), alphas_bar_sqrt[-1].clone()
    alphas_bar_sqrt = (alphas_bar_sqrt - aT) * (a0 / (a0 - aT))
    alphas_bar = alphas_bar_sqrt ** 2
    alphas = torch.cat([alphas_bar[0:1], alphas_bar[1:] / alphas_bar[:-1]])
    return 1 - alphas


class DDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02,
                 beta_schedule="linear", trained_betas=None, clip_sample=True, set_alpha_to_one=True,
                 steps_offset=0, prediction_type="epsilon", thresholding=False,
                 dynamic_thresholding_ratio=0.995, clip_sample_range=1.0, sample_max_value=1.0,
                 timestep_spacing="leading", rescale_betas_zero_snr=False):
        if thresholding or clip_sample:
            raise NotImplementedError("clip_sample / thresholding are not used by the reference config")
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                   dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        if rescale_betas_zero_snr:
            betas = _rescale_zero_terminal_snr(betas)
        self.betas = betas
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps,
                                      prediction_type=prediction_type, steps_offset=steps_offset,
                                      timestep_spacing=timestep_spacing, clip_sample=clip_sample)
        if prediction_type not in ("v_prediction", "epsilon"):
            raise NotImplementedError(prediction_type)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        sp = self.config.timestep_spacing
        if sp == "trailing":
            ts = np.round(np.arange(T, 0, -T / num_inference_steps)).astype(np.int64) - 1
        elif sp == "leading":
            ratio = T // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
            ts += self.config.steps_offset
        elif sp == "linspace":
            ts = np.linspace(0, T - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        else:
            raise ValueError(sp)
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step_coefficients(self, timestep: int):
        """(sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)) as fp32, and t_prev."""
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        coef = torch.stack([a_t.sqrt(), (1 - a_t).sqrt(), a_prev.sqrt(), (1 - a_prev).sqrt()]).float()
        return coef, prev_t

    def step(self, model_output, timestep, sample, eta: float = 0.0, generator=None, **kw):
        """model_output / sample: CUDA fp16 [1, c, F, h, w].  eta must be 0 (deterministic DDIM)."""
        if eta != 0.0:
            raise NotImplementedError("eta > 0 (stochastic DDIM) is not used by the reference")
        from . import ops
        if not sample.is_cuda:
            raise RuntimeError("DDIMScheduler.step runs in the CUDA kernel only (no CPU path)")
        coef, _ = self.step_coefficients(int(timestep))
        coef = coef.to(sample.device)
        acc = model_output.float().contiguous()
        F_ = sample.shape[2]
        counter = torch.ones(F_, dtype=torch.float32, device=sample.device)
        prev = sample.to(torch.float16).contiguous().clone()
        ops.cfg_ddim_step(acc, counter, prev, coef, 1.0, self.config.prediction_type == "v_prediction")
        return SimpleNamespace(prev_sample=prev.to(sample.dtype))
