"""Seeded tiny-configuration cases of the three pipelines and of the whole agent step, shared by
  * make_pipeline_golden.py  (runs the fp32 CPU oracle on them and writes pipelines_tiny.npz — committed),
  * tests/test_golden_pipelines.py  (CPU: the oracle still reproduces the committed vectors; GPU: the device pipelines,
    through the reference-facing calls and the C ABI, match the committed vectors).
No upstream golden vectors exist for this path (the reference has no tests, diffusers / RoboBase cannot be
installed offline), so these are generated from the oracle by the committed script, as §8c prescribes."""
import dataclasses

import numpy as np
import torch

from genima_b200 import weights as W
from genima_b200.configs import ACTConfig, UNetConfig, VAEConfig

N_STEPS = 2
CASES = ("controlnet", "pix2pix", "sdxl", "agent_step")


def _common(ucfg, seed):
    g = torch.Generator().manual_seed(seed)
    ctx = torch.randn(1, 77, ucfg.cross_attention_dim, generator=g).half().float()
    cond = torch.randint(0, 256, (1, 128, 128, 3), generator=g, dtype=torch.uint8)
    return g, ctx, cond


def build(name: str) -> dict:
    """-> dict of configs, host state dicts and inputs for one case (all tensors on the CPU)."""
    if name == "controlnet":
        ucfg, vcfg = UNetConfig.tiny(), VAEConfig.tiny()
        g, ctx, cond = _common(ucfg, 101)
        return dict(ucfg=ucfg, vcfg=vcfg, ctx=ctx, cond=cond,
                    unet=W.synth_state_dict(W.unet_shapes(ucfg)), controlnet=W.synth_state_dict(W.controlnet_shapes(ucfg), 1),
                    vae=W.synth_state_dict(W.vae_decoder_shapes(vcfg), 2),
                    lat=torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(2)).half().float())
    if name == "pix2pix":
        ucfg, vcfg = dataclasses.replace(UNetConfig.tiny(), in_channels=8), VAEConfig.tiny()
        g, ctx, cond = _common(ucfg, 102)
        vae = W.synth_state_dict(W.vae_decoder_shapes(vcfg), 2)
        vae.update(W.synth_state_dict(W.vae_encoder_shapes(vcfg), 2))
        return dict(ucfg=ucfg, vcfg=vcfg, ctx=ctx, cond=cond, unet=W.synth_state_dict(W.unet_shapes(ucfg), 5), vae=vae,
                    lat=torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(2)).half().float())
    if name == "sdxl":
        ucfg, vcfg = UNetConfig.sdxl_tiny(), dataclasses.replace(VAEConfig.tiny(), scaling_factor=0.13025)
        g, ctx, cond = _common(ucfg, 103)
        pooled = torch.randn(1, ucfg.projection_input_dim - 6 * ucfg.addition_time_embed_dim, generator=g).half().float()
        # the draws the pipeline makes from a CPU generator seeded with 2: latents, then one noise tensor per step (fp16)
        gen = torch.Generator().manual_seed(2)
        lat = torch.randn(1, 4, 16, 16, generator=gen, dtype=torch.float16)
        noises = [torch.randn(1, 4, 16, 16, generator=gen, dtype=torch.float16).float() for _ in range(N_STEPS)]
        return dict(ucfg=ucfg, vcfg=vcfg, ctx=ctx, cond=cond, pooled=pooled, lat=lat.float(), noises=noises,
                    unet=W.synth_state_dict(W.unet_shapes(ucfg)), controlnet=W.synth_state_dict(W.controlnet_shapes(ucfg), 1),
                    vae=W.synth_state_dict(W.vae_decoder_shapes(vcfg), 2))
    if name == "agent_step":
        ucfg, vcfg, acfg = UNetConfig.tiny(), VAEConfig.tiny(), ACTConfig.tiny()
        g = torch.Generator().manual_seed(104)
        S = acfg.image_size
        return dict(ucfg=ucfg, vcfg=vcfg, acfg=acfg,
                    views=torch.randint(0, 256, (4, S, S, 3), generator=g, dtype=torch.uint8),
                    ctx=torch.randn(1, 77, ucfg.cross_attention_dim, generator=g).half().float(),
                    lat=torch.randn(1, 4, S // 4, S // 4, generator=torch.Generator().manual_seed(2)).half().float(),
                    qpos=torch.randn(1, acfg.state_dim, generator=g), task=torch.randn(1, acfg.task_emb_dim, generator=g),
                    unet=W.synth_state_dict(W.unet_shapes(ucfg)), controlnet=W.synth_state_dict(W.controlnet_shapes(ucfg), 1),
                    vae=W.synth_state_dict(W.vae_decoder_shapes(vcfg), 2), act=W.synth_state_dict(W.act_shapes(acfg), 3))
    raise KeyError(name)


def run_oracle(name: str) -> dict:
    """fp32 CPU oracle outputs of one case as numpy arrays (what make_pipeline_golden.py stores)."""
    from oracle import pipeline as P

    c = build(name)
    if name == "controlnet":
        r = P.controlnet_pipeline(c["unet"], c["controlnet"], c["vae"], c["ucfg"], c["vcfg"], c["cond"].numpy(), c["ctx"],
                                  c["lat"], N_STEPS)
    elif name == "pix2pix":
        r = P.pix2pix_pipeline(c["unet"], c["vae"], c["ucfg"], c["vcfg"], c["cond"].numpy(), c["ctx"], c["lat"], N_STEPS)
    elif name == "sdxl":
        r = P.sdxl_controlnet_pipeline(c["unet"], c["controlnet"], c["vae"], c["ucfg"], c["vcfg"], c["cond"].numpy(),
                                       c["ctx"], c["pooled"], c["lat"], c["noises"], N_STEPS)
    else:
        r = P.agent_step(dict(unet=c["unet"], controlnet=c["controlnet"], vae=c["vae"], act=c["act"]), c["ucfg"],
                         c["vcfg"], c["acfg"], c["views"].numpy(), c["ctx"], c["lat"], c["qpos"], c["task"], N_STEPS)
        return {"agent_step.a_hat": r["a_hat"].numpy().astype(np.float32),
                "agent_step.latents": r["latents"].numpy().astype(np.float32),
                "agent_step.tile_u8": np.asarray(r["tile_u8"], dtype=np.uint8)}
    return {f"{name}.latents": r["latents"].numpy().astype(np.float32), f"{name}.u8": np.asarray(r["u8"], dtype=np.uint8)}
