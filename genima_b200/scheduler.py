"""Host-side scheduler tables for the denoise loop (diffusers EulerDiscreteScheduler as shipped by sd-turbo, and
EulerAncestralDiscreteScheduler as shipped by sdxl-turbo).

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
        if cfg.class_name not in ("EulerDiscreteScheduler", "EulerAncestralDiscreteScheduler", "DDIMScheduler"):
            raise NotImplementedError(f"scheduler {cfg.class_name!r} is not implemented (EulerDiscreteScheduler, "
                                      "EulerAncestralDiscreteScheduler and DDIMScheduler only)")
        # DDIM (eta = 0): the same ODE step in the variance-preserving variable — no input
        # scaling, unit initial noise, x' = a x + b eps with host scalars (a, b) per step
        self.ddim = cfg.class_name == "DDIMScheduler"
        if self.ddim and cfg.clip_sample:
            raise NotImplementedError("DDIMScheduler with clip_sample=True (clamping the predicted x0) is not "
                                      "implemented; Stable Diffusion snapshots ship clip_sample=false")
        # stabilityai/sdxl-turbo ships the ancestral variant: same tables, but every step re-injects noise
        self.ancestral = cfg.class_name == "EulerAncestralDiscreteScheduler"
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
        self.alphas_cumprod = alphas_cumprod.numpy().astype(np.float64)
        self.timesteps = None
        self.sigmas = None

    def set_timesteps(self, n: int) -> Tuple[np.ndarray, np.ndarray]:
        T = self.cfg.num_train_timesteps
        sp = self.cfg.timestep_spacing
        if self.ddim:
            return self._set_timesteps_ddim(n)
        if sp == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)).astype(np.float64) - 1
        elif sp == "leading":
            # EulerDiscreteScheduler.set_timesteps: (arange(n) * (T // n)).round()[::-1] + steps_offset
            ts = (np.arange(0, n) * (T // n)).round()[::-1].astype(np.float64) + self.cfg.steps_offset
        elif sp == "linspace":
            ts = np.linspace(0, T - 1, n, dtype=np.float64)[::-1].copy()
        else:
            raise NotImplementedError(f"timestep_spacing {sp!r}")
        sig = np.interp(ts, np.arange(0, T), self.train_sigmas)
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32)
        self.timesteps = ts.astype(np.float32)
        return self.timesteps, self.sigmas

    def _set_timesteps_ddim(self, n: int) -> Tuple[np.ndarray, np.ndarray]:
        """DDIMScheduler.set_timesteps.  `sigmas` is all zeros here: scale_model_input is the identity and
        init_noise_sigma is 1, which is exactly what the Euler-form plumbing computes for sigma = 0."""
        T = self.cfg.num_train_timesteps
        sp = self.cfg.timestep_spacing
        if sp == "leading":
            ts = (np.arange(0, n) * (T // n)).round()[::-1].astype(np.int64) + self.cfg.steps_offset
        elif sp == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)).astype(np.int64) - 1
        elif sp == "linspace":
            ts = np.linspace(0, T - 1, n).round()[::-1].astype(np.int64)
        else:
            raise NotImplementedError(f"timestep_spacing {sp!r}")
        self._ddim_stride = T // n
        self.timesteps = ts.astype(np.float32)
        self.sigmas = np.zeros(n + 1, dtype=np.float32)
        return self.timesteps, self.sigmas

    def ddim_coeffs(self, i: int) -> Tuple[float, float]:
        """(a, b) of step i: x_prev = sqrt(ab_prev) (x - sqrt(1 - ab_t) eps) / sqrt(ab_t) + sqrt(1 - ab_prev) eps
        = a x + b eps (DDIMScheduler.step, eta = 0, no clipping / thresholding); prev_timestep = t - T // n upstream."""
        t = int(self.timesteps[i])
        prev = t - self._ddim_stride
        ab_t = float(self.alphas_cumprod[t])
        ab_p = float(self.alphas_cumprod[prev]) if prev >= 0 else (1.0 if self.cfg.set_alpha_to_one
                                                                   else float(self.alphas_cumprod[0]))
        a = (ab_p / ab_t) ** 0.5
        b = (1.0 - ab_p) ** 0.5 - (ab_p * (1.0 - ab_t) / ab_t) ** 0.5
        return a, b

    @property
    def init_noise_sigma(self) -> float:
        if self.ddim:
            return 1.0
        m = float(self.sigmas.max())
        if self.cfg.timestep_spacing in ("linspace", "trailing"):
            return m
        return float((m * m + 1.0) ** 0.5)

    def ancestral_sigmas(self, i: int) -> Tuple[float, float]:
        """(sigma_up, sigma_down) of step i (EulerAncestralDiscreteScheduler.step): the noise re-injected after the
        deterministic move from sigmas[i] down to sigma_down.  The last step (sigma_to = 0) has sigma_up = 0."""
        s_from, s_to = float(self.sigmas[i]), float(self.sigmas[i + 1])
        up = (s_to ** 2 * (s_from ** 2 - s_to ** 2) / s_from ** 2) ** 0.5
        down = (s_to ** 2 - up ** 2) ** 0.5
        return up, down
