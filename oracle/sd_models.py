"""fp32 CPU restatement of the diffusers 0.29.0 graphs on Genima's hot path (TEST INFRASTRUCTURE — see oracle/__init__).

diffusers is pinned by the reference (pyproject.toml:24, poetry.lock:595-596) but is not vendored and cannot be
installed offline, so these functions restate its published architecture (SURVEY.md Appendices A-C), operating on the
upstream state-dict key names.  PARITY UNPINNED: no upstream golden vector exists.

Upstream modules restated (file paths inside diffusers 0.29.0):
  models/embeddings.py            get_timestep_embedding, TimestepEmbedding
  models/resnet.py                ResnetBlock2D, Downsample2D, Upsample2D
  models/attention.py             BasicTransformerBlock, FeedForward(GEGLU)
  models/transformers/transformer_2d.py   Transformer2DModel (use_linear_projection=True)
  models/unets/unet_2d_condition.py       UNet2DConditionModel.forward
  models/controlnet.py            ControlNetModel.forward, ControlNetConditioningEmbedding
  models/autoencoders/vae.py      Decoder, Encoder; autoencoder_kl.py AutoencoderKL.decode / .encode
Reference call sites: controller/agent/sd_controlnet_agent.py:31-42 (model construction), :67-76 (pipe call);
controller/agent/sd_pix2pix_agent.py:29-41, :52-60 (InstructPix2Pix sibling).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .configs import taesd_layer_plan

UNetConfig = VAEConfig = object   # duck-typed configuration objects (oracle/configs.py)

SD = Dict[str, torch.Tensor]


def _w(sd: SD, key: str) -> torch.Tensor:
    return sd[key].to(torch.float32)


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0, max_period=10000)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t.to(torch.float32)[:, None] * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def time_embed(sd: SD, cfg: UNetConfig, t: torch.Tensor, added: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """time_embedding(time_proj(t)) [+ SDXL get_aug_embed: add_embedding(cat([text_embeds, add_time_proj(time_ids)]))
    when the model has addition_embed_type='text_time'; added = dict(text_embeds [B, P], time_ids [B, 6])]."""
    emb = timestep_embedding(t, cfg.block_out_channels[0])
    w1 = _w(sd, "time_embedding.linear_1.weight")
    emb = F.linear(emb.to(w1.dtype), w1, _w(sd, "time_embedding.linear_1.bias"))  # sinusoid in fp32, then model dtype
    emb = F.silu(emb)
    emb = F.linear(emb, _w(sd, "time_embedding.linear_2.weight"), _w(sd, "time_embedding.linear_2.bias"))
    if cfg.addition_embed:
        if added is None:
            raise ValueError("this model needs added_cond_kwargs (text_embeds, time_ids)")
        ids = added["time_ids"].to(torch.float32)
        tid = timestep_embedding(ids.flatten(), cfg.addition_time_embed_dim).reshape(ids.shape[0], -1)
        a = torch.cat([added["text_embeds"].to(tid.dtype), tid], dim=-1).to(w1.dtype)
        a = F.silu(F.linear(a, _w(sd, "add_embedding.linear_1.weight"), _w(sd, "add_embedding.linear_1.bias")))
        emb = emb + F.linear(a, _w(sd, "add_embedding.linear_2.weight"), _w(sd, "add_embedding.linear_2.bias"))
    return emb


def resnet_block(sd: SD, p: str, x: torch.Tensor, temb: Optional[torch.Tensor], groups: int, eps: float):
    """ResnetBlock2D (output_scale_factor=1, dropout=0, time_embedding_norm='default')."""
    h = F.group_norm(x, groups, _w(sd, f"{p}.norm1.weight"), _w(sd, f"{p}.norm1.bias"), eps)
    h = F.silu(h)
    h = F.conv2d(h, _w(sd, f"{p}.conv1.weight"), _w(sd, f"{p}.conv1.bias"), padding=1)
    if temb is not None and f"{p}.time_emb_proj.weight" in sd:
        t = F.linear(F.silu(temb), _w(sd, f"{p}.time_emb_proj.weight"), _w(sd, f"{p}.time_emb_proj.bias"))
        h = h + t[:, :, None, None]
    h = F.group_norm(h, groups, _w(sd, f"{p}.norm2.weight"), _w(sd, f"{p}.norm2.bias"), eps)
    h = F.silu(h)
    h = F.conv2d(h, _w(sd, f"{p}.conv2.weight"), _w(sd, f"{p}.conv2.bias"), padding=1)
    if f"{p}.conv_shortcut.weight" in sd:
        x = F.conv2d(x, _w(sd, f"{p}.conv_shortcut.weight"), _w(sd, f"{p}.conv_shortcut.bias"))
    return x + h


def attention(sd: SD, p: str, x: torch.Tensor, ctx: torch.Tensor, heads: int) -> torch.Tensor:
    """Attention (AttnProcessor2_0): to_q/k/v without bias, to_out.0 with bias, scale 1/sqrt(head_dim)."""
    b, n, c = x.shape
    q = F.linear(x, _w(sd, f"{p}.to_q.weight"))
    k = F.linear(ctx, _w(sd, f"{p}.to_k.weight"))
    v = F.linear(ctx, _w(sd, f"{p}.to_v.weight"))
    d = c // heads
    q = q.reshape(b, n, heads, d).transpose(1, 2)
    k = k.reshape(b, -1, heads, d).transpose(1, 2)
    v = v.reshape(b, -1, heads, d).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    o = torch.softmax(s, dim=-1) @ v
    o = o.transpose(1, 2).reshape(b, n, c)
    return F.linear(o, _w(sd, f"{p}.to_out.0.weight"), _w(sd, f"{p}.to_out.0.bias"))


def transformer2d(sd: SD, p: str, x: torch.Tensor, ctx: torch.Tensor, heads: int, groups: int) -> torch.Tensor:
    """Transformer2DModel with its BasicTransformerBlocks (one for SD-2.x, `transformer_layers_per_block` for SDXL; the
    count is read off the state dict), linear projections, GN eps 1e-6, LN eps 1e-5."""
    b, c, hh, ww = x.shape
    res = x
    h = F.group_norm(x, groups, _w(sd, f"{p}.norm.weight"), _w(sd, f"{p}.norm.bias"), 1e-6)
    h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
    h = F.linear(h, _w(sd, f"{p}.proj_in.weight"), _w(sd, f"{p}.proj_in.bias"))
    k = 0
    while f"{p}.transformer_blocks.{k}.norm1.weight" in sd:
        t = f"{p}.transformer_blocks.{k}"
        n = F.layer_norm(h, (c,), _w(sd, f"{t}.norm1.weight"), _w(sd, f"{t}.norm1.bias"), 1e-5)
        h = h + attention(sd, f"{t}.attn1", n, n, heads)
        n = F.layer_norm(h, (c,), _w(sd, f"{t}.norm2.weight"), _w(sd, f"{t}.norm2.bias"), 1e-5)
        h = h + attention(sd, f"{t}.attn2", n, ctx, heads)
        n = F.layer_norm(h, (c,), _w(sd, f"{t}.norm3.weight"), _w(sd, f"{t}.norm3.bias"), 1e-5)
        proj = F.linear(n, _w(sd, f"{t}.ff.net.0.proj.weight"), _w(sd, f"{t}.ff.net.0.proj.bias"))
        hidden, gate = proj.chunk(2, dim=-1)
        h = h + F.linear(hidden * F.gelu(gate), _w(sd, f"{t}.ff.net.2.weight"), _w(sd, f"{t}.ff.net.2.bias"))
        k += 1
    h = F.linear(h, _w(sd, f"{p}.proj_out.weight"), _w(sd, f"{p}.proj_out.bias"))
    return h.reshape(b, hh, ww, c).permute(0, 3, 1, 2) + res


def _encoder(sd: SD, cfg: UNetConfig, h: torch.Tensor, temb: torch.Tensor, ctx: torch.Tensor):
    """conv_in output -> (mid-block output, skip list S0..S11)."""
    g, eps = cfg.norm_num_groups, cfg.norm_eps
    skips = [h]
    n_levels = len(cfg.block_out_channels)
    for i in range(n_levels):
        for j in range(cfg.layers_per_block):
            h = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", h, temb, g, eps)
            if cfg.attn_levels[i]:
                h = transformer2d(sd, f"down_blocks.{i}.attentions.{j}", h, ctx, cfg.num_heads[i], g)
            skips.append(h)
        if i < n_levels - 1:
            p = f"down_blocks.{i}.downsamplers.0.conv"
            h = F.conv2d(h, _w(sd, f"{p}.weight"), _w(sd, f"{p}.bias"), stride=2, padding=1)
            skips.append(h)
    h = resnet_block(sd, "mid_block.resnets.0", h, temb, g, eps)
    h = transformer2d(sd, "mid_block.attentions.0", h, ctx, cfg.num_heads[-1], g)
    h = resnet_block(sd, "mid_block.resnets.1", h, temb, g, eps)
    return h, skips


def cond_embedding(sd: SD, cfg: UNetConfig, cond: torch.Tensor) -> torch.Tensor:
    """ControlNetConditioningEmbedding: conv_in, SiLU, [conv, SiLU, conv(stride 2), SiLU] x3, conv_out."""
    p = "controlnet_cond_embedding"
    e = F.silu(F.conv2d(cond, _w(sd, f"{p}.conv_in.weight"), _w(sd, f"{p}.conv_in.bias"), padding=1))
    for k in range(2 * (len(cfg.cond_embed_channels) - 1)):
        stride = 2 if k % 2 == 1 else 1
        e = F.silu(F.conv2d(e, _w(sd, f"{p}.blocks.{k}.weight"), _w(sd, f"{p}.blocks.{k}.bias"), stride=stride,
                            padding=1))
    return F.conv2d(e, _w(sd, f"{p}.conv_out.weight"), _w(sd, f"{p}.conv_out.bias"), padding=1)


def controlnet_forward(sd: SD, cfg: UNetConfig, x: torch.Tensor, t: torch.Tensor, ctx: torch.Tensor,
                       cond: torch.Tensor, conditioning_scale: float = 1.0, added=None):
    """ControlNetModel.forward(guess_mode=False) -> (12 down residuals, mid residual)."""
    temb = time_embed(sd, cfg, t.expand(x.shape[0]), added)
    h = F.conv2d(x, _w(sd, "conv_in.weight"), _w(sd, "conv_in.bias"), padding=1)
    h = h + cond_embedding(sd, cfg, cond)
    mid, skips = _encoder(sd, cfg, h, temb, ctx)
    down = []
    for i, s in enumerate(skips):
        p = f"controlnet_down_blocks.{i}"
        down.append(F.conv2d(s, _w(sd, f"{p}.weight"), _w(sd, f"{p}.bias")) * conditioning_scale)
    mid = F.conv2d(mid, _w(sd, "controlnet_mid_block.weight"), _w(sd, "controlnet_mid_block.bias")) * conditioning_scale
    return down, mid


def unet_forward(sd: SD, cfg: UNetConfig, x: torch.Tensor, t: torch.Tensor, ctx: torch.Tensor,
                 down_residuals: Optional[List[torch.Tensor]] = None, mid_residual: Optional[torch.Tensor] = None,
                 added=None):
    """UNet2DConditionModel.forward with ControlNet residuals added to the skips and to the mid-block output."""
    g, eps = cfg.norm_num_groups, cfg.norm_eps
    temb = time_embed(sd, cfg, t.expand(x.shape[0]), added)
    h = F.conv2d(x, _w(sd, "conv_in.weight"), _w(sd, "conv_in.bias"), padding=1)
    h, skips = _encoder(sd, cfg, h, temb, ctx)
    if down_residuals is not None:
        skips = [s + r for s, r in zip(skips, down_residuals)]
    if mid_residual is not None:
        h = h + mid_residual
    n_levels = len(cfg.block_out_channels)
    for i in range(n_levels):
        level = n_levels - 1 - i
        for j in range(cfg.layers_per_block + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", h, temb, g, eps)
            if cfg.attn_levels[level]:
                h = transformer2d(sd, f"up_blocks.{i}.attentions.{j}", h, ctx, cfg.num_heads[level], g)
        if i < n_levels - 1:
            p = f"up_blocks.{i}.upsamplers.0.conv"
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, _w(sd, f"{p}.weight"), _w(sd, f"{p}.bias"), padding=1)
    h = F.group_norm(h, g, _w(sd, "conv_norm_out.weight"), _w(sd, "conv_norm_out.bias"), eps)
    h = F.silu(h)
    return F.conv2d(h, _w(sd, "conv_out.weight"), _w(sd, "conv_out.bias"), padding=1)


def _vae_mid_attention(sd: SD, a: str, h: torch.Tensor, g: int, eps: float) -> torch.Tensor:
    """UNetMidBlock2D's Attention in the VAE: one head over all channels, linear projections WITH bias, residual."""
    b, c, hh, ww = h.shape
    n = F.group_norm(h, g, _w(sd, f"{a}.group_norm.weight"), _w(sd, f"{a}.group_norm.bias"), eps)
    n = n.reshape(b, c, hh * ww).transpose(1, 2)
    q = F.linear(n, _w(sd, f"{a}.to_q.weight"), _w(sd, f"{a}.to_q.bias"))
    k = F.linear(n, _w(sd, f"{a}.to_k.weight"), _w(sd, f"{a}.to_k.bias"))
    v = F.linear(n, _w(sd, f"{a}.to_v.weight"), _w(sd, f"{a}.to_v.bias"))
    o = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(c), dim=-1) @ v
    o = F.linear(o, _w(sd, f"{a}.to_out.0.weight"), _w(sd, f"{a}.to_out.0.bias"))
    return h + o.transpose(1, 2).reshape(b, c, hh, ww)


def vae_encode_mean(sd: SD, cfg: VAEConfig, image: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.encode(image).latent_dist.mode(): Encoder -> quant_conv -> first half of the moments (the mean of
    the diagonal Gaussian; its mode).  image: [B, 3, H, W] in [-1, 1] -> [B, latent_channels, H/8, W/8], NOT multiplied
    by scaling_factor (StableDiffusionInstructPix2PixPipeline.prepare_image_latents uses the raw mode; reached from
    controller/agent/sd_pix2pix_agent.py:52-60).  [upstream, from memory: diffusers 0.29.0 models/autoencoders/vae.py
    Encoder, DownEncoderBlock2D; Downsample2D(padding=0) pads (0, 1, 0, 1) before its stride-2 convolution]"""
    g, eps = cfg.norm_num_groups, cfg.norm_eps
    ch = cfg.block_out_channels
    h = F.conv2d(image, _w(sd, "encoder.conv_in.weight"), _w(sd, "encoder.conv_in.bias"), padding=1)
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block):
            h = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, None, g, eps)
        if i < len(ch) - 1:
            p = f"encoder.down_blocks.{i}.downsamplers.0.conv"
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0.0)
            h = F.conv2d(h, _w(sd, f"{p}.weight"), _w(sd, f"{p}.bias"), stride=2)
    h = resnet_block(sd, "encoder.mid_block.resnets.0", h, None, g, eps)
    h = _vae_mid_attention(sd, "encoder.mid_block.attentions.0", h, g, eps)
    h = resnet_block(sd, "encoder.mid_block.resnets.1", h, None, g, eps)
    h = F.group_norm(h, g, _w(sd, "encoder.conv_norm_out.weight"), _w(sd, "encoder.conv_norm_out.bias"), eps)
    h = F.conv2d(F.silu(h), _w(sd, "encoder.conv_out.weight"), _w(sd, "encoder.conv_out.bias"), padding=1)
    moments = F.conv2d(h, _w(sd, "quant_conv.weight"), _w(sd, "quant_conv.bias"))
    return moments[:, :cfg.latent_channels]


def vae_decode(sd: SD, cfg: VAEConfig, z: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.decode(z): post_quant_conv -> Decoder.  The caller divides latents by scaling_factor first."""
    g, eps = cfg.norm_num_groups, cfg.norm_eps
    h = F.conv2d(z, _w(sd, "post_quant_conv.weight"), _w(sd, "post_quant_conv.bias"))
    h = F.conv2d(h, _w(sd, "decoder.conv_in.weight"), _w(sd, "decoder.conv_in.bias"), padding=1)
    h = resnet_block(sd, "decoder.mid_block.resnets.0", h, None, g, eps)
    h = _vae_mid_attention(sd, "decoder.mid_block.attentions.0", h, g, eps)
    h = resnet_block(sd, "decoder.mid_block.resnets.1", h, None, g, eps)
    n_levels = len(cfg.block_out_channels)
    for i in range(n_levels):
        for j in range(cfg.layers_per_block + 1):
            h = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, None, g, eps)
        if i < n_levels - 1:
            p = f"decoder.up_blocks.{i}.upsamplers.0.conv"
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, _w(sd, f"{p}.weight"), _w(sd, f"{p}.bias"), padding=1)
    h = F.group_norm(h, g, _w(sd, "decoder.conv_norm_out.weight"), _w(sd, "decoder.conv_norm_out.bias"), eps)
    h = F.silu(h)
    return F.conv2d(h, _w(sd, "decoder.conv_out.weight"), _w(sd, "decoder.conv_out.bias"), padding=1)


def taesd_decode(sd: SD, cfg, z: torch.Tensor) -> torch.Tensor:
    """diffusers AutoencoderTiny.decode(z) = DecoderTiny.forward: tanh(z / 3) * 3 -> conv / ReLU stack with nearest x2
    upsampling -> x * 2 - 1 (images in [-1, 1], like AutoencoderKL.decode).  Reached from
    controller/agent/sd_controlnet_agent.py:45-49 when eval_cfg.autoencoder names a TAESD checkpoint; the pipeline divides
    the latents by scaling_factor (1.0) first.  [upstream, from memory: diffusers 0.29.0 models/autoencoders/vae.py]"""
    h = torch.tanh(z / cfg.latent_magnitude) * cfg.latent_magnitude
    for kind, i in taesd_layer_plan(cfg):
        p = f"decoder.layers.{i}"
        if kind in ("conv_in", "conv_out"):
            h = F.conv2d(h, _w(sd, f"{p}.weight"), _w(sd, f"{p}.bias"), padding=1)
        elif kind == "relu":
            h = F.relu(h)
        elif kind == "block":
            y = F.relu(F.conv2d(h, _w(sd, f"{p}.conv.0.weight"), _w(sd, f"{p}.conv.0.bias"), padding=1))
            y = F.relu(F.conv2d(y, _w(sd, f"{p}.conv.2.weight"), _w(sd, f"{p}.conv.2.bias"), padding=1))
            y = F.conv2d(y, _w(sd, f"{p}.conv.4.weight"), _w(sd, f"{p}.conv.4.bias"), padding=1)
            h = F.relu(y + h)
        elif kind == "up":
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
        elif kind == "conv":
            h = F.conv2d(h, _w(sd, f"{p}.weight"), None, padding=1)
    return h * 2.0 - 1.0
