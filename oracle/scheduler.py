"""CPU restatement of diffusers 0.29.0 EulerDiscreteScheduler as shipped with stabilityai/sd-turbo
(schedulers/scheduling_euler_discrete.py: __init__, set_timesteps, scale_model_input, step with s_churn = 0).
TEST INFRASTRUCTURE — see oracle/__init__.  The reference never overrides pipe.scheduler on the eval path
(controller/agent/sd_controlnet_agent.py:31-42), so this is the scheduler its `pipe(...)` call runs.
"""
from __future__ import annotations

import numpy as np
import torch


class EulerDiscreteOracle:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, timestep_spacing="trailing",
                 steps_offset=0):
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.train_sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        self.timestep_spacing = timestep_spacing
        self.timesteps = None
        self.sigmas = None

    def set_timesteps(self, n: int):
        T = self.num_train_timesteps
        if self.timestep_spacing == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)).astype(np.float64) - 1
        elif self.timestep_spacing == "leading":
            # upstream: timesteps = (arange(n) * step_ratio).round()[::-1]; timesteps += self.config.steps_offset
            ts = (np.arange(0, n) * (T // n)).round()[::-1].astype(np.float64) + self.steps_offset
        elif self.timestep_spacing == "linspace":
            ts = np.linspace(0, T - 1, n, dtype=np.float64)[::-1].copy()
        else:
            raise ValueError(self.timestep_spacing)
        sig = np.interp(ts, np.arange(0, T), self.train_sigmas)
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32)
        self.timesteps = ts.astype(np.float32)
        return self.timesteps, self.sigmas

    @property
    def init_noise_sigma(self) -> float:
        m = float(self.sigmas.max())
        if self.timestep_spacing in ("linspace", "trailing"):
            return m
        return float((m ** 2 + 1) ** 0.5)

    def scale_model_input(self, x: torch.Tensor, i: int) -> torch.Tensor:
        s = float(self.sigmas[i])
        return x / ((s ** 2 + 1) ** 0.5)

    def step(self, eps: torch.Tensor, i: int, x: torch.Tensor) -> torch.Tensor:
        """Upstream arithmetic order, in fp32: x0 = x - sigma*eps; d = (x - x0)/sigma; x + d*(sigma_next - sigma)."""
        sigma = float(self.sigmas[i])
        sigma_next = float(self.sigmas[i + 1])
        x = x.to(torch.float32)
        x0 = x - sigma * eps.to(torch.float32)
        d = (x - x0) / sigma
        return x + d * (sigma_next - sigma)


class EulerAncestralOracle(EulerDiscreteOracle):
    """diffusers 0.29.0 EulerAncestralDiscreteScheduler (schedulers/scheduling_euler_ancestral_discrete.py) as shipped
    with stabilityai/sdxl-turbo: same sigma tables / scaling as EulerDiscrete; step() moves to sigma_down and adds
    sigma_up * noise.  [upstream, from memory]  `noise` is passed in (upstream draws randn_tensor(model_output.shape,
    dtype=model_output.dtype, generator=generator) inside step)."""

    def step(self, eps: torch.Tensor, i: int, x: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        sigma_from, sigma_to = float(self.sigmas[i]), float(self.sigmas[i + 1])
        x = x.to(torch.float32)
        x0 = x - sigma_from * eps.to(torch.float32)
        sigma_up = (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5
        sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
        d = (x - x0) / sigma_from
        return x + d * (sigma_down - sigma_from) + noise.to(torch.float32) * sigma_up


class DDIMOracle:
    """diffusers 0.29.0 DDIMScheduler (schedulers/scheduling_ddim.py), eta = 0, epsilon prediction, clip_sample and
    thresholding off (what Stable Diffusion snapshots that ship DDIM configure) [upstream, from memory]: identity
    scale_model_input, init_noise_sigma 1, prev_timestep = t - num_train_timesteps // num_inference_steps."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, timestep_spacing="leading",
                 steps_offset=0, set_alpha_to_one=True):
        self.T = num_train_timesteps
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0).double()
        self.final_alpha_cumprod = 1.0 if set_alpha_to_one else float(self.alphas_cumprod[0])
        self.timestep_spacing, self.steps_offset = timestep_spacing, steps_offset
        self.init_noise_sigma = 1.0

    def set_timesteps(self, n: int):
        T = self.T
        self.n = n
        if self.timestep_spacing == "leading":
            ts = (np.arange(0, n) * (T // n)).round()[::-1].copy().astype(np.int64) + self.steps_offset
        elif self.timestep_spacing == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)).astype(np.int64) - 1
        elif self.timestep_spacing == "linspace":
            ts = np.linspace(0, T - 1, n).round()[::-1].copy().astype(np.int64)
        else:
            raise ValueError(self.timestep_spacing)
        self.timesteps = ts
        self.sigmas = np.zeros(n + 1, dtype=np.float32)      # (interface parity with the Euler oracles)
        return ts.astype(np.float32), self.sigmas

    def scale_model_input(self, x: torch.Tensor, i: int) -> torch.Tensor:
        return x

    def step(self, eps: torch.Tensor, i: int, x: torch.Tensor) -> torch.Tensor:
        t = int(self.timesteps[i])
        prev = t - self.T // self.n
        a_t = float(self.alphas_cumprod[t])
        a_p = float(self.alphas_cumprod[prev]) if prev >= 0 else self.final_alpha_cumprod
        x, eps = x.to(torch.float64), eps.to(torch.float64)
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return (a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps).to(torch.float32)
