"""Host-side scheduler tables for the denoise loop (diffusers EulerDiscreteScheduler as shipped by sd-turbo).

Only scalar, data-independent host logic lives here (timesteps, sigmas); the per-element update runs in
gn_euler_step.  The reference keeps whatever scheduler `stabilityai/sd-turbo` ships (it never assigns pipe.scheduler on
the eval path, controller/agent/sd_controlnet_agent.py:31-42): EulerDiscreteScheduler, trailing spacing, epsilon
prediction.  Unknown scheduler classes raise instead of silently running the wrong ODE.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .configs import SchedulerConfig


class EulerDiscreteSchedule:
    def __init__(self, cfg: SchedulerConfig = SchedulerConfig()):
        if cfg.class_name != "EulerDiscreteScheduler":
            raise NotImplementedError(f"scheduler {cfg.class_name!r} is not implemented (EulerDiscreteScheduler only)")
        if cfg.prediction_type != "epsilon" or cfg.beta_schedule != "scaled_linear":
            raise NotImplementedError("only epsilon prediction with the scaled_linear beta schedule is implemented")
        self.cfg = cfg
        T = cfg.num_train_timesteps
        # float32 torch ops in the same order as diffusers' __init__ (torch.linspace(...)**2, cumprod), so the tables
        # are bit-identical to the upstream scheduler's; host scalars only, computed once
        import torch

        betas = torch.linspace(cfg.beta_start ** 0.5, cfg.beta_end ** 0.5, T, dtype=torch.float32) ** 2
        alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.train_sigmas = (((1 - alphas_cumprod) / alphas_cumprod) ** 0.5).numpy()
        self.timesteps = None
        self.sigmas = None

    def set_timesteps(self, n: int) -> Tuple[np.ndarray, np.ndarray]:
        T = self.cfg.num_train_timesteps
        sp = self.cfg.timestep_spacing
        if sp == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)).astype(np.float64) - 1
        elif sp == "leading":
            ts = (np.arange(0, n) * (T // n)).round()[::-1].astype(np.float64)
        elif sp == "linspace":
            ts = np.linspace(0, T - 1, n, dtype=np.float64)[::-1].copy()
        else:
            raise NotImplementedError(f"timestep_spacing {sp!r}")
        sig = np.interp(ts, np.arange(0, T), self.train_sigmas)
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32)
        self.timesteps = ts.astype(np.float32)
        return self.timesteps, self.sigmas

    @property
    def init_noise_sigma(self) -> float:
        m = float(self.sigmas.max())
        if self.cfg.timestep_spacing in ("linspace", "trailing"):
            return m
        return float((m * m + 1.0) ** 0.5)
