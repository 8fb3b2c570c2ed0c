"""ORACLE (test infrastructure).  DDIM scheduler of `diffusers==0.24.0` restated in numpy/torch fp32
for the one configuration the reference uses (configs/inference/mikudance_config.yaml:24-33:
linear betas 0.00085..0.012, zero-terminal-SNR rescale, trailing spacing, v-prediction,
clip_sample false, eta 0; built at scripts/inference_video.py:101-102, called at
src/pipelines/pipeline_mikudance.py:400-401,631-633,676-678).

The diffusers source is not under /root/reference and not installed: these formulas restate the
published 0.24.0 `DDIMScheduler` (parity "pinned to a restatement"); the integer tables are checked
against the values SURVEY.md §8a16 lists.
"""
import numpy as np
import torch


def rescale_zero_terminal_snr(betas: torch.Tensor) -> torch.Tensor:
    alphas = 1.0 - betas
    alphas_cumprod = torch.cumprod(alphas, dim=0)
    alphas_bar_sqrt = alphas_cumprod.sqrt()
    a0 = alphas_bar_sqrt[0].clone()
    aT = alphas_bar_sqrt[-1].clone()
    alphas_bar_sqrt -= aT
    alphas_bar_sqrt *= a0 / (a0 - aT)
    alphas_bar = alphas_bar_sqrt ** 2
    alphas = alphas_bar[1:] / alphas_bar[:-1]
    alphas = torch.cat([alphas_bar[0:1], alphas])
    return 1 - alphas


class DDIMOracle:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 beta_schedule="linear", clip_sample=False, steps_offset=1,
                 prediction_type="v_prediction", rescale_betas_zero_snr=True,
                 timestep_spacing="trailing", set_alpha_to_one=True):
        assert beta_schedule == "linear" and not clip_sample
        self.T = num_train_timesteps
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        if rescale_betas_zero_snr:
            betas = rescale_zero_terminal_snr(betas)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.prediction_type = prediction_type
        self.timestep_spacing = timestep_spacing
        self.steps_offset = steps_offset
        self.init_noise_sigma = 1.0
        self.order = 1

    def set_timesteps(self, n):
        self.num_inference_steps = n
        if self.timestep_spacing == "trailing":
            step_ratio = self.T / n
            ts = np.round(np.arange(self.T, 0, -step_ratio)).astype(np.int64) - 1
        elif self.timestep_spacing == "leading":
            step_ratio = self.T // n
            ts = (np.arange(0, n) * step_ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        else:
            raise NotImplementedError(self.timestep_spacing)
        self.timesteps = ts
        return ts

    def coefficients(self, t: int):
        prev_t = t - self.T // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        return prev_t, a_t, a_prev

    def step(self, model_output: torch.Tensor, t: int, sample: torch.Tensor) -> torch.Tensor:
        """eta = 0, no clipping / thresholding."""
        _, a_t, a_prev = self.coefficients(int(t))
        b_t = 1 - a_t
        if self.prediction_type == "v_prediction":
            x0 = a_t ** 0.5 * sample - b_t ** 0.5 * model_output
            eps = a_t ** 0.5 * model_output + b_t ** 0.5 * sample
        elif self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        else:
            raise NotImplementedError(self.prediction_type)
        return a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps
