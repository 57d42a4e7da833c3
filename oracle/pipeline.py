"""fp32 CPU restatement of one Genima agent step (TEST INFRASTRUCTURE — see oracle/__init__).

  controlnet_pipeline()  diffusers 0.29.0 StableDiffusionControlNetPipeline.__call__ exactly as the reference exercises it
                         (controller/agent/sd_controlnet_agent.py:67-76): no CFG (guidance 0.0,
                         controller/cfgs/eval_genima.yaml:31), control image in [0, 1] without normalisation, explicit
                         `latents` multiplied by init_noise_sigma, Euler-trailing loop with ControlNet residuals added to
                         the U-Net skips, VAE decode of latents / scaling_factor, postprocess to uint8.
  sdxl_controlnet_pipeline()  the SDXL-ControlNet sibling (controller/agent/sdxl_controlnet_agent.py): two text encoders'
                         penultimate states, text_time added conditioning, Euler-ancestral steps.
  pix2pix_pipeline()     the InstructPix2Pix sibling (controller/agent/sd_pix2pix_agent.py): VAE-encoded image latents
                         concatenated to the U-Net input, no ControlNet.
  agent_step()           controller/eval_genima.py:163-249 between `obs` and `actions`: tile_images -> pipeline ->
                         untile_images -> GenimaACT.act (policy forward on the four generated views).
PARITY UNPINNED (no upstream package, weights or golden vectors offline); pinned sub-pieces are listed in oracle/__init__.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

ACTConfig = UNetConfig = VAEConfig = object   # duck-typed configuration objects (oracle/configs.py)

from . import act as act_oracle
from . import sd_models, tiling
from .scheduler import EulerDiscreteOracle


def controlnet_pipeline(unet_sd, cn_sd, vae_sd, ucfg: UNetConfig, vcfg: VAEConfig, cond_u8: np.ndarray,
                        ctx: torch.Tensor, latents: torch.Tensor, n_steps: int, conditioning_scale: float = 1.0,
                        return_intermediates: bool = False, scheduler=None):
    """cond_u8 [B, H, W, 3] uint8; ctx [B, 77, D] fp32 prompt embeddings; latents [B, 4, H/8, W/8] unit-variance noise.
    `scheduler`: an oracle scheduler object (default: sd-turbo's EulerDiscrete; oracle.scheduler.DDIMOracle for
    snapshots that ship DDIMScheduler).
    Returns dict(latents=final latents fp32, image=decoded image in [-1, 1] NCHW fp32, u8=[B, H, W, 3] uint8)."""
    sched = scheduler if scheduler is not None else EulerDiscreteOracle()
    ts, sig = sched.set_timesteps(n_steps)
    cond = torch.from_numpy(cond_u8.astype(np.float32) / 255.0).permute(0, 3, 1, 2)      # VaeImageProcessor.preprocess
    x = latents.to(torch.float32) * sched.init_noise_sigma
    inter = []
    for i, t in enumerate(ts):
        xs = sched.scale_model_input(x, i)
        tt = torch.tensor([float(t)])
        down, mid = sd_models.controlnet_forward(cn_sd, ucfg, xs, tt, ctx, cond, conditioning_scale)
        eps = sd_models.unet_forward(unet_sd, ucfg, xs, tt, ctx, down, mid)
        x = sched.step(eps, i, x)
        if return_intermediates:
            inter.append(dict(eps=eps, x=x))
    # AutoencoderKL, or AutoencoderTiny when the caller selected TAESD (sd_controlnet_agent.py:45-49)
    decode = sd_models.taesd_decode if hasattr(vcfg, "num_blocks") else sd_models.vae_decode
    img = decode(vae_sd, vcfg, x / vcfg.scaling_factor)
    den = (img / 2 + 0.5).clamp(0, 1)                                                      # postprocess, denormalize
    u8 = (den.permute(0, 2, 3, 1).numpy() * 255).round().astype(np.uint8)
    out = dict(latents=x, image=img, u8=u8)
    if return_intermediates:
        out["steps"] = inter
    return out


def sdxl_controlnet_pipeline(unet_sd, cn_sd, vae_sd, ucfg: UNetConfig, vcfg: VAEConfig, cond_u8: np.ndarray,
                             ctx: torch.Tensor, pooled: torch.Tensor, latents: torch.Tensor, step_noise, n_steps: int,
                             conditioning_scale: float = 1.0):
    """diffusers 0.29.0 StableDiffusionXLControlNetPipeline.__call__ as controller/agent/sdxl_controlnet_agent.py:66-75
    exercises it (guidance 0.0 -> no CFG) [upstream, from memory]: ctx = cat of the two encoders' penultimate hidden
    states [B, 77, 2048]; added_cond_kwargs = {text_embeds: pooled, time_ids: (H, W, 0, 0, H, W)} for both ControlNet
    and U-Net; EulerAncestralDiscreteScheduler (sdxl-turbo) with the per-step noise passed in as step_noise[i]
    [B, 4, h, w] (upstream draws it from `generator` inside scheduler.step)."""
    from .scheduler import EulerAncestralOracle

    sched = EulerAncestralOracle()
    ts, sig = sched.set_timesteps(n_steps)
    B, H, W, _ = cond_u8.shape
    cond = torch.from_numpy(cond_u8.astype(np.float32) / 255.0).permute(0, 3, 1, 2)
    added = dict(text_embeds=pooled.to(torch.float32).expand(B, -1),
                 time_ids=torch.tensor([[H, W, 0, 0, H, W]], dtype=torch.float32).expand(B, -1))
    x = latents.to(torch.float32) * sched.init_noise_sigma
    for i, t in enumerate(ts):
        xs = sched.scale_model_input(x, i)
        tt = torch.tensor([float(t)])
        down, mid = sd_models.controlnet_forward(cn_sd, ucfg, xs, tt, ctx, cond, conditioning_scale, added)
        eps = sd_models.unet_forward(unet_sd, ucfg, xs, tt, ctx, down, mid, added)
        x = sched.step(eps, i, x, step_noise[i])
    decode = sd_models.taesd_decode if hasattr(vcfg, "num_blocks") else sd_models.vae_decode
    img = decode(vae_sd, vcfg, x / vcfg.scaling_factor)
    den = (img / 2 + 0.5).clamp(0, 1)
    u8 = (den.permute(0, 2, 3, 1).numpy() * 255).round().astype(np.uint8)
    return dict(latents=x, image=img, u8=u8)


def pix2pix_pipeline(unet_sd, vae_sd, ucfg: UNetConfig, vcfg: VAEConfig, image_u8: np.ndarray, ctx: torch.Tensor,
                     latents: torch.Tensor, n_steps: int):
    """diffusers 0.29.0 StableDiffusionInstructPix2PixPipeline.__call__ as controller/agent/sd_pix2pix_agent.py:52-60
    exercises it: guidance_scale 0.0 (controller/cfgs/eval_genima.yaml:31) -> do_classifier_free_guidance is False, so
    one U-Net evaluation per step on cat([scale_model_input(latents), image_latents], dim=1) with the 8-channel conv_in;
    image_latents = vae.encode(preprocess(image)).latent_dist.mode() (image normalised to [-1, 1], latents NOT scaled).
    [upstream, from memory.  Older diffusers releases converted eps to x0 and back around the guidance step for
    sigma-space schedulers; without guidance that round trip is the identity, so either variant computes this.]
    image_u8 [B, H, W, 3]; ctx [B, 77, D]; latents [B, 4, H/8, W/8] unit-variance noise."""
    sched = EulerDiscreteOracle()
    ts, sig = sched.set_timesteps(n_steps)
    img = torch.from_numpy(image_u8.astype(np.float32) / 255.0).permute(0, 3, 1, 2) * 2.0 - 1.0
    image_latents = sd_models.vae_encode_mean(vae_sd, vcfg, img)
    x = latents.to(torch.float32) * sched.init_noise_sigma
    for i, t in enumerate(ts):
        xs = torch.cat([sched.scale_model_input(x, i), image_latents], dim=1)
        eps = sd_models.unet_forward(unet_sd, ucfg, xs, torch.tensor([float(t)]), ctx)
        x = sched.step(eps, i, x)
    out = sd_models.vae_decode(vae_sd, vcfg, x / vcfg.scaling_factor)
    den = (out / 2 + 0.5).clamp(0, 1)
    u8 = (den.permute(0, 2, 3, 1).numpy() * 255).round().astype(np.uint8)
    return dict(latents=x, image=out, u8=u8, image_latents=image_latents)


def agent_step(weights: Dict[str, dict], ucfg: UNetConfig, vcfg: VAEConfig, acfg: ACTConfig, views_u8: np.ndarray,
               ctx: torch.Tensor, latents: torch.Tensor, qpos: torch.Tensor, task_emb: torch.Tensor, n_steps: int):
    """views_u8 [4, 256, 256, 3] (camera order) -> dict(a_hat [1, nq, A], tile_u8, gen_views_u8 [4, 256, 256, 3]).
    `weights`: dict(unet=, controlnet=, vae=, act=) state dicts."""
    tile = tiling.tile_views(views_u8)[None]
    out = controlnet_pipeline(weights["unet"], weights["controlnet"], weights["vae"], ucfg, vcfg, tile, ctx, latents,
                              n_steps)
    gen = tiling.untile_views(out["u8"][0])                                               # [4, 256, 256, 3]
    image = torch.from_numpy(gen).permute(0, 3, 1, 2)[None].float()                       # [1, V, 3, H, W] 0..255
    a_hat, is_pad = act_oracle.act_forward(weights["act"], acfg, qpos, image, task_emb)
    return dict(a_hat=a_hat, is_pad=is_pad, tile_u8=out["u8"], gen_views_u8=gen, latents=out["latents"],
                image=out["image"])
