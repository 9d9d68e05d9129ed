"""DDPM / DDIM schedulers with the reference's interface; the update itself runs in one fused CUDA kernel.

Mirrors diffusers' DDPMScheduler / DDIMScheduler as Tango uses them
(/root/reference/mustango/diffusers/src/diffusers/schedulers/scheduling_ddpm.py:122-349,
scheduling_ddim.py:132-359; call sites models.py:224-249, tango.py:36): `set_timesteps`, `timesteps`,
`init_noise_sigma`, `order`, `scale_model_input`, `step(...).prev_sample`, `config`.

All per-step scalars are computed on the host with the reference's own fp32 torch ops (same association order),
packed into a [num_steps, 10] coefficient table and shipped to the device once per `set_timesteps`; the kernel
(tng_sched_step) then evaluates  x0 = (c0*s + c1*v)/c9, prev = c2*x0 + c3*s + c7*(c5*s + c6*v) + c4*noise  with
un-fused multiplies/adds, which reproduces the reference CPU arithmetic bit for bit and removes the two host syncs
per step of the reference.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import lib as L

# stabilityai/stable-diffusion-2-1 `scheduler/scheduler_config.json` — what Tango loads (tango.py:36, models.py:80-81).
# The JSON is not in the reference tree; these are its published values and every field can be overridden.
SD21_SCHEDULER_CONFIG = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                             beta_schedule="scaled_linear", prediction_type="v_prediction", clip_sample=False,
                             set_alpha_to_one=False, steps_offset=1, skip_prk_steps=True, trained_betas=None)

NCOEF = 10


class SchedulerOutput(SimpleNamespace):
    pass


class _Config(dict):
    __getattr__ = dict.__getitem__


def _betas(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas=None):
    if trained_betas is not None:
        return torch.tensor(trained_betas, dtype=torch.float32)
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    raise NotImplementedError(f"{beta_schedule} is not implemented")


class _SchedulerBase:
    order = 1

    def __init__(self, **cfg):
        self.config = _Config(cfg)
        self.betas = _betas(cfg["num_train_timesteps"], cfg["beta_start"], cfg["beta_end"], cfg["beta_schedule"],
                            cfg.get("trained_betas"))
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, cfg["num_train_timesteps"])[::-1].copy().astype(np.int64))
        self._coef_host: Optional[torch.Tensor] = None   # [num_steps, NCOEF] fp32 (CPU)
        self._coef_dev: Optional[torch.Tensor] = None
        self._t_index: dict = {}

    @classmethod
    def from_pretrained(cls, name: str = "stabilityai/stable-diffusion-2-1", subfolder: str = "scheduler", **overrides):
        """The reference downloads the SD-2.1 scheduler JSON from the hub (tango.py:36, models.py:80-81). Offline:
        a local directory is read (`<name>/<subfolder>/scheduler_config.json` or `<name>/scheduler_config.json`);
        `stabilityai/stable-diffusion-2-1` (or None) maps to its published values; any other name is refused rather
        than silently sampled with the wrong betas / prediction type."""
        import json
        import os
        base = dict(SD21_SCHEDULER_CONFIG)
        if name is not None and os.path.isdir(str(name)):
            for cand in (os.path.join(name, subfolder or "", "scheduler_config.json"), os.path.join(name, "scheduler_config.json")):
                if os.path.exists(cand):
                    with open(cand) as f:
                        base.update({k: v for k, v in json.load(f).items() if not k.startswith("_")})
                    break
            else:
                raise FileNotFoundError(f"no scheduler_config.json under '{name}'")
        elif name not in (None, "stabilityai/stable-diffusion-2-1"):
            raise ValueError(f"scheduler '{name}' is not reachable offline: pass a local directory holding its "
                             "scheduler_config.json (only stabilityai/stable-diffusion-2-1 is built in)")
        base.update(overrides)
        return cls(**{k: v for k, v in base.items() if k in cls._ACCEPTED})

    def __len__(self):
        return self.config["num_train_timesteps"]

    def scale_model_input(self, sample, timestep=None):
        return sample

    # ---------------------------------------------------------------------------------------------------------
    def _grid(self, n: int) -> np.ndarray:
        T = self.config["num_train_timesteps"]
        if n > T:
            raise ValueError(
                f"`num_inference_steps`: {n} cannot be larger than `self.config.train_timesteps`: {T} as the unet"
                f" model trained with this scheduler can only handle maximal {T} timesteps.")
        ratio = T // n
        return (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64)

    def _finish_set_timesteps(self, device):
        self._t_list = [int(t) for t in self.timesteps.tolist()]
        key = tuple(self._t_list)
        cache = self.__dict__.setdefault("_table_cache", {})
        if key not in cache:   # ~40 tiny fp32 torch ops per step: computed once per grid, reused by later calls
            cache[key] = torch.stack([self._coefficients(t) for t in self._t_list]).contiguous()
        self._coef_host = cache[key]
        self._t_index = {t: i for i, t in enumerate(self._t_list)}
        self._coef_dev = None
        if device is not None:
            self.timesteps = self.timesteps.to(device)
            if torch.device(device).type == "cuda":
                self._coef_dev = self._coef_host.to(device)

    def timestep_at(self, i: int) -> int:
        """Host copy of timesteps[i] (no device sync)."""
        return self._t_list[i]

    def coefficient_table(self, device=None) -> torch.Tensor:
        """[num_steps, 10] fp32 table (row i belongs to timesteps[i])."""
        if self._coef_host is None:
            self._finish_set_timesteps(None)
        if device is None:
            return self._coef_host
        if self._coef_dev is None or self._coef_dev.device != torch.device(device):
            self._coef_dev = self._coef_host.to(device)
        return self._coef_dev

    def _row(self, timestep) -> int:
        t = int(timestep)
        if self._coef_host is None or t not in self._t_index:
            # arbitrary timestep outside the current grid (the reference allows it): one-row table
            self._coef_host = self._coefficients(t)[None].contiguous()
            self._t_index = {t: 0}
            self._coef_dev = None
        return self._t_index[t]

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, generator=None,
             variance_noise: Optional[torch.Tensor] = None, return_dict: bool = True, **_unused):
        """x_t -> x_{t-1} for NCHW fp32 CUDA tensors (reference layout). Noise comes from the torch RNG exactly as
        in the reference (randn of model_output's shape when t > 0) unless `variance_noise` is given."""
        L.require_cuda(model_output)   # no CPU fallback
        i = self._row(timestep)
        coef = self.coefficient_table(sample.device)[i]
        B, Cc, H, W = sample.shape
        noise = None
        if self._needs_noise(int(timestep)):
            noise = variance_noise
            if noise is None:
                noise = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                                    dtype=model_output.dtype)
            noise = noise.contiguous().float()
        mo = model_output.float().permute(0, 2, 3, 1).contiguous().view(B * H * W, Cc)  # channels-last rows
        prev = torch.empty_like(sample, dtype=torch.float32)
        L.sched_step(mo, False, 1.0, sample.contiguous().float(), noise, coef.contiguous(), prev, None, B=B, Cc=Cc,
                     HW=H * W)
        if not return_dict:
            return (prev,)
        return SchedulerOutput(prev_sample=prev)


class DDPMScheduler(_SchedulerBase):
    _ACCEPTED = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "trained_betas", "variance_type",
                 "clip_sample", "prediction_type", "clip_sample_range")

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, variance_type="fixed_small", clip_sample=True, prediction_type="epsilon",
                 clip_sample_range=1.0):
        if variance_type != "fixed_small":
            raise NotImplementedError("only variance_type='fixed_small' (Tango's) is implemented")
        super().__init__(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                         beta_schedule=beta_schedule, trained_betas=trained_betas, variance_type=variance_type,
                         clip_sample=clip_sample, prediction_type=prediction_type,
                         clip_sample_range=clip_sample_range)
        self.one = torch.tensor(1.0)
        self.variance_type = variance_type

    def set_timesteps(self, num_inference_steps: int, device=None):
        """scheduling_ddpm.py:184-204: t_i = (i * (T // N)) reversed, int64 (no steps_offset in this version)."""
        self.timesteps = torch.from_numpy(self._grid(num_inference_steps))
        self.num_inference_steps = num_inference_steps
        self._finish_set_timesteps(device)

    def _needs_noise(self, t: int) -> bool:
        return t > 0

    def _coefficients(self, t: int) -> torch.Tensor:
        """scheduling_ddpm.py:283-344 scalar arithmetic, same fp32 torch ops in the same order."""
        cfg = self.config
        n = self.num_inference_steps if self.num_inference_steps else cfg["num_train_timesteps"]
        prev_t = t - cfg["num_train_timesteps"] // n
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        one, zero = torch.tensor(1.0), torch.tensor(0.0)
        if cfg["prediction_type"] == "epsilon":
            c_x0_s, c_x0_m, c_div = one, -(b_t ** 0.5), a_t ** 0.5
        elif cfg["prediction_type"] == "sample":
            c_x0_s, c_x0_m, c_div = zero, one, one
        elif cfg["prediction_type"] == "v_prediction":
            c_x0_s, c_x0_m, c_div = a_t ** 0.5, -(b_t ** 0.5), one
        else:
            raise ValueError(f"prediction_type given as {cfg['prediction_type']} must be one of `epsilon`, `sample` or"
                             " `v_prediction`  for the DDPMScheduler.")
        c_prev_x0 = (a_prev ** 0.5 * cur_beta) / b_t
        c_prev_s = cur_alpha ** 0.5 * b_prev / b_t
        c_noise = zero
        if t > 0:
            var = (1 - a_prev) / (1 - a_t) * cur_beta   # _get_variance :206-224
            var = torch.clamp(var, min=1e-20)
            c_noise = var ** 0.5
        clip = torch.tensor(float(cfg["clip_sample_range"]) if cfg["clip_sample"] else 0.0)
        return torch.stack([c_x0_s, c_x0_m, c_prev_x0, c_prev_s, c_noise, zero, zero, zero, clip, c_div]).float()


class DDIMScheduler(_SchedulerBase):
    _ACCEPTED = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "trained_betas", "clip_sample",
                 "set_alpha_to_one", "steps_offset", "prediction_type", "clip_sample_range")

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, clip_sample=True, set_alpha_to_one=True, steps_offset=0,
                 prediction_type="epsilon", clip_sample_range=1.0):
        super().__init__(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                         beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
                         set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                         prediction_type=prediction_type, clip_sample_range=clip_sample_range)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]

    def set_timesteps(self, num_inference_steps: int, device=None):
        """scheduling_ddim.py:214-236: same grid as DDPM plus steps_offset."""
        self.timesteps = torch.from_numpy(self._grid(num_inference_steps)) + self.config["steps_offset"]
        self.num_inference_steps = num_inference_steps
        self._finish_set_timesteps(device)

    def _needs_noise(self, t: int) -> bool:
        return False  # eta = 0 (deterministic DDIM)

    def _coefficients(self, t: int) -> torch.Tensor:
        """scheduling_ddim.py:292-354 with eta = 0."""
        cfg = self.config
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the"
                             " scheduler")
        prev_t = t - cfg["num_train_timesteps"] // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        one, zero = torch.tensor(1.0), torch.tensor(0.0)
        if cfg["prediction_type"] == "epsilon":
            c_x0_s, c_x0_m, c_div = one, -(b_t ** 0.5), a_t ** 0.5
            c_eps_s, c_eps_m = zero, one
        elif cfg["prediction_type"] == "v_prediction":
            c_x0_s, c_x0_m, c_div = a_t ** 0.5, -(b_t ** 0.5), one
            c_eps_s, c_eps_m = b_t ** 0.5, a_t ** 0.5
        else:
            raise ValueError(f"prediction_type given as {cfg['prediction_type']} must be one of `epsilon` or"
                             " `v_prediction` for the fused DDIM step")
        b_prev = 1 - a_prev
        variance = (b_prev / b_t) * (1 - a_t / a_prev)
        std = 0.0 * variance ** 0.5
        c_prev_eps = (1 - a_prev - std ** 2) ** 0.5
        c_prev_x0 = a_prev ** 0.5
        clip = torch.tensor(float(cfg["clip_sample_range"]) if cfg["clip_sample"] else 0.0)
        return torch.stack([c_x0_s, c_x0_m, c_prev_x0, zero, zero, c_eps_s, c_eps_m, c_prev_eps, clip,
                            c_div]).float()
