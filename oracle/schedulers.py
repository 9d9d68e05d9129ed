"""CPU restatement of DDPMScheduler / DDIMScheduler (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/mustango/diffusers/src/diffusers/schedulers/scheduling_ddpm.py and scheduling_ddim.py
operation by operation in torch fp32 on the CPU, so the results are bit-identical to the reference run on CPU.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

# stabilityai/stable-diffusion-2-1 scheduler_config.json (not in the reference tree)
SD21_CONFIG = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                   prediction_type="v_prediction", clip_sample=False, set_alpha_to_one=False, steps_offset=1)


def make_betas(num_train_timesteps, beta_start, beta_end, beta_schedule):
    """scheduling_ddpm.py:138-152."""
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    raise NotImplementedError(beta_schedule)


class OracleDDPM:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 variance_type="fixed_small", clip_sample=True, prediction_type="epsilon", clip_sample_range=1.0,
                 **_ignored):
        self.T = num_train_timesteps
        self.betas = make_betas(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)  # :154-155
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.variance_type = variance_type
        self.clip_sample, self.clip_range, self.prediction_type = clip_sample, clip_sample_range, prediction_type
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def set_timesteps(self, n):
        """:184-204."""
        if n > self.T:
            raise ValueError("num_inference_steps > num_train_timesteps")
        self.num_inference_steps = n
        ratio = self.T // n
        self.timesteps = torch.from_numpy((np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64))

    def _get_variance(self, t):
        """:206-224 (fixed_small)."""
        n = self.num_inference_steps if self.num_inference_steps else self.T
        prev_t = t - self.T // n
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        cur_beta = 1 - a_t / a_prev
        var = (1 - a_prev) / (1 - a_t) * cur_beta
        return torch.clamp(var, min=1e-20)

    def step(self, model_output, t, sample, noise: Optional[torch.Tensor] = None):
        """:254-349. `noise` replaces the randn_tensor draw of :331-335 (same shape as model_output)."""
        t = int(t)
        n = self.num_inference_steps if self.num_inference_steps else self.T
        prev_t = t - self.T // n
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        if self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        elif self.prediction_type == "sample":
            x0 = model_output
        elif self.prediction_type == "v_prediction":
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
        else:
            raise ValueError(self.prediction_type)
        if self.clip_sample:
            x0 = x0.clamp(-self.clip_range, self.clip_range)
        c0 = (a_prev ** 0.5 * cur_beta) / b_t
        ct = cur_alpha ** 0.5 * b_prev / b_t
        prev = c0 * x0 + ct * sample
        variance = 0
        if t > 0:
            assert noise is not None
            variance = (self._get_variance(t) ** 0.5) * noise
        return prev + variance


class OracleDDIM:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon",
                 clip_sample_range=1.0, **_ignored):
        self.T = num_train_timesteps
        self.betas = make_betas(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.steps_offset = steps_offset
        self.clip_sample, self.clip_range, self.prediction_type = clip_sample, clip_sample_range, prediction_type
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, n):
        """scheduling_ddim.py:214-236."""
        if n > self.T:
            raise ValueError("num_inference_steps > num_train_timesteps")
        self.num_inference_steps = n
        ratio = self.T // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts) + self.steps_offset

    def step(self, model_output, t, sample, eta: float = 0.0):
        """scheduling_ddim.py:238-359 with eta = 0."""
        t = int(t)
        prev_t = t - self.T // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        if self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        elif self.prediction_type == "v_prediction":
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
            eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        else:
            raise ValueError(self.prediction_type)
        if self.clip_sample:
            x0 = x0.clamp(-self.clip_range, self.clip_range)
        b_prev = 1 - a_prev
        variance = (b_prev / b_t) * (1 - a_t / a_prev)
        std = eta * variance ** 0.5
        direction = (1 - a_prev - std ** 2) ** 0.5 * eps
        return a_prev ** 0.5 * x0 + direction
