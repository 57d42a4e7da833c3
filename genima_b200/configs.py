"""Architecture hyper-parameters of the networks on Genima's per-step hot path.

Defaults are the shipped configuration of the reference (SURVEY.md Appendices B, C, E, F):
  * `stabilityai/sd-turbo` U-Net / ControlNet (SD-2.1-base topology) and KL-VAE decoder, loaded by
    controller/agent/sd_controlnet_agent.py:31-42,
  * OpenCLIP-H text encoder (SD text tower) and CLIP ViT-B/32 text tower (controller/method/genima_act.py:314-346),
  * RoboBase ACT as configured by controller/cfgs/method/genima_act.yaml.
`tiny()` variants keep every structural feature (all block types, skip concat, stride-2, cross attention, FiLM) at
sizes the CPU oracle runs in well under a second; they are what the whole-network parity tests use.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    # diffusers calls this `attention_head_dim` but for SD-2.x it holds the HEAD COUNT per level (head_dim = 64)
    num_heads: Tuple[int, ...] = (5, 10, 20, 20)
    # levels whose down/up blocks carry Transformer2D layers (CrossAttn*Block2D); the last level is plain
    attn_levels: Tuple[bool, ...] = (True, True, True, False)
    cross_attention_dim: int = 1024
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    cond_embed_channels: Tuple[int, ...] = (16, 32, 96, 256)  # ControlNet conditioning embedding
    sample_size: int = 64
    # BasicTransformerBlocks per Transformer2DModel at each level (diffusers transformer_layers_per_block; the mid block
    # uses the last entry).  () = one everywhere (SD-2.x); SDXL: (1, 2, 10)
    transformer_layers: Tuple[int, ...] = ()
    # SDXL addition_embed_type="text_time": emb += add_embedding(cat(pooled text embeds, sinusoid(time_ids)))
    addition_embed: bool = False
    addition_time_embed_dim: int = 256
    projection_input_dim: int = 2816          # pooled text embeds (1280) + 6 time ids x 256

    def tf_layers(self, level: int) -> int:
        return self.transformer_layers[level] if self.transformer_layers else 1

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4

    @staticmethod
    def tiny() -> "UNetConfig":
        return UNetConfig(block_out_channels=(64, 128, 128, 128), num_heads=(1, 2, 2, 2), cross_attention_dim=128,
                          cond_embed_channels=(16, 32, 64, 64), sample_size=16)

    @staticmethod
    def sdxl() -> "UNetConfig":
        """stabilityai/sdxl-turbo (and SDXL base) U-Net / ControlNet topology: three levels, no attention at level 0,
        1 / 2 / 10 transformer blocks, 2048-wide context (two text encoders), text_time added conditioning."""
        return UNetConfig(block_out_channels=(320, 640, 1280), num_heads=(5, 10, 20), attn_levels=(False, True, True),
                          cross_attention_dim=2048, transformer_layers=(1, 2, 10), addition_embed=True, sample_size=64)

    @staticmethod
    def sdxl_tiny() -> "UNetConfig":
        return UNetConfig(block_out_channels=(64, 128, 128), num_heads=(1, 2, 2), attn_levels=(False, True, True),
                          cross_attention_dim=256, cond_embed_channels=(16, 32, 64, 64), transformer_layers=(1, 2, 3),
                          addition_embed=True, addition_time_embed_dim=32, projection_input_dim=64 + 6 * 32,
                          sample_size=16)


@dataclass(frozen=True)
class VAEConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2  # decoder uses layers_per_block + 1 resnets per up block
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.18215
    # AutoencoderKL config.json `force_upcast` (stabilityai/sdxl-turbo's VAE ships true: diffusers then decodes in fp32
    # because that VAE overflows fp16).  The fp16 tensor-core decoder here checks its output for non-finite values when
    # the flag is set (pipeline.py) instead of silently returning a black image.
    force_upcast: bool = False

    @staticmethod
    def tiny() -> "VAEConfig":
        return VAEConfig(block_out_channels=(64, 64, 128, 128))


@dataclass(frozen=True)
class TAESDConfig:
    """diffusers AutoencoderTiny (madebyollin/taesd) decoder, selected by `autoencoder: <...taesd...>` in the reference's
    eval config (controller/agent/sd_controlnet_agent.py:45-49): conv / ReLU only, 1.22 M parameters."""
    latent_channels: int = 4
    out_channels: int = 3
    channels: int = 64
    num_blocks: Tuple[int, ...] = (3, 3, 3, 1)   # decoder_block_out_channels = (64, 64, 64, 64)
    latent_magnitude: float = 3.0                # DecoderTiny.forward: tanh(x / 3) * 3
    scaling_factor: float = 1.0

    @staticmethod
    def tiny() -> "TAESDConfig":
        return TAESDConfig(num_blocks=(1, 1, 1, 1))

    @property
    def block_out_channels(self) -> Tuple[int, ...]:   # one entry per resolution level (pipeline's vae_scale_factor)
        return (self.channels,) * len(self.num_blocks)


@dataclass(frozen=True)
class CLIPTextConfig:
    vocab_size: int = 49408
    hidden_size: int = 1024
    intermediate_size: int = 4096
    num_layers: int = 23
    num_heads: int = 16
    max_positions: int = 77
    act: str = "gelu"          # OpenCLIP-H text tower (SD-2.x): exact GELU; OpenAI ViT-B/32: quick_gelu
    eps: float = 1e-5
    projection_dim: int = 0    # > 0: EOT-pooled output times text_projection (OpenAI CLIP encode_text)

    @staticmethod
    def sd_turbo() -> "CLIPTextConfig":
        return CLIPTextConfig()

    @staticmethod
    def sdxl_clip_l() -> "CLIPTextConfig":
        """SDXL text_encoder: OpenAI CLIP ViT-L/14 text tower (12 layers, d 768, quick_gelu)."""
        return CLIPTextConfig(hidden_size=768, intermediate_size=3072, num_layers=12, num_heads=12, act="quick_gelu")

    @staticmethod
    def sdxl_open_clip_bigg() -> "CLIPTextConfig":
        """SDXL text_encoder_2: OpenCLIP ViT-bigG/14 text tower with projection (32 layers, d 1280, GELU)."""
        return CLIPTextConfig(hidden_size=1280, intermediate_size=5120, num_layers=32, num_heads=20, act="gelu",
                              projection_dim=1280)

    @staticmethod
    def vit_b32() -> "CLIPTextConfig":
        return CLIPTextConfig(hidden_size=512, intermediate_size=2048, num_layers=12, num_heads=8, act="quick_gelu",
                              projection_dim=512)

    @staticmethod
    def tiny(projection_dim: int = 0) -> "CLIPTextConfig":
        return CLIPTextConfig(vocab_size=1000, hidden_size=128, intermediate_size=256, num_layers=2, num_heads=2,
                              act="quick_gelu" if projection_dim else "gelu", projection_dim=projection_dim)


@dataclass(frozen=True)
class ACTConfig:
    hidden_dim: int = 256
    enc_layers: int = 4
    dec_layers: int = 6
    dim_feedforward: int = 2048
    nheads: int = 8
    num_queries: int = 20
    state_dim: int = 8
    action_dim: int = 8
    latent_dim: int = 32
    num_views: int = 4
    image_size: int = 256
    task_emb_dim: int = 512
    resnet_widths: Tuple[int, ...] = (64, 128, 256, 512)
    bn_eps: float = 1e-5
    ln_eps: float = 1e-5

    @staticmethod
    def tiny() -> "ACTConfig":
        return ACTConfig(hidden_dim=64, enc_layers=1, dec_layers=2, dim_feedforward=128, nheads=2, num_queries=4,
                         image_size=64, task_emb_dim=64, resnet_widths=(64, 64, 64, 64))


@dataclass(frozen=True)
class SchedulerConfig:
    """stabilityai/sd-turbo scheduler_config.json: EulerDiscreteScheduler, trailing spacing, epsilon prediction."""
    class_name: str = "EulerDiscreteScheduler"
    num_train_timesteps: int = 1000
    beta_start: float = 0.00085
    beta_end: float = 0.012
    beta_schedule: str = "scaled_linear"
    timestep_spacing: str = "trailing"
    prediction_type: str = "epsilon"
    # DDIMScheduler only (diffusers defaults; Stable Diffusion snapshots ship clip_sample / set_alpha_to_one = false)
    steps_offset: int = 0
    set_alpha_to_one: bool = True
    clip_sample: bool = False
