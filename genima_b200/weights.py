"""State-dict schemas (upstream key names -> shapes) and seeded synthetic weights.

The key names are the ones the upstream checkpoints use (diffusers `diffusion_pytorch_model.safetensors`, transformers
`CLIPTextModel`, torchvision ResNet, DETR-style ACT transformer), so a real checkpoint binds
through the same table.  No weights are available offline, so tests and benchmarks use `synth_state_dict`: every tensor
is drawn from a generator seeded by crc32(name), rounded to fp16 once, and shared by the CPU oracle and the device
path — weight quantisation is therefore not part of any parity error.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .configs import ACTConfig, CLIPTextConfig, UNetConfig, VAEConfig

Shapes = "OrderedDict[str, Tuple[int, ...]]"


# --------------------------------------------------------------------------------------------------- schema helpers
def _conv(s, name, cout, cin, k):
    s[f"{name}.weight"] = (cout, cin, k, k)
    s[f"{name}.bias"] = (cout,)


def _lin(s, name, cout, cin, bias=True):
    s[f"{name}.weight"] = (cout, cin)
    if bias:
        s[f"{name}.bias"] = (cout,)


def _norm(s, name, c):
    s[f"{name}.weight"] = (c,)
    s[f"{name}.bias"] = (c,)


def _resnet(s, p, cin, cout, temb_dim):
    _norm(s, f"{p}.norm1", cin)
    _conv(s, f"{p}.conv1", cout, cin, 3)
    if temb_dim:
        _lin(s, f"{p}.time_emb_proj", cout, temb_dim)
    _norm(s, f"{p}.norm2", cout)
    _conv(s, f"{p}.conv2", cout, cout, 3)
    if cin != cout:
        _conv(s, f"{p}.conv_shortcut", cout, cin, 1)


def _transformer2d(s, p, c, ctx_dim, depth=1):
    _norm(s, f"{p}.norm", c)
    _lin(s, f"{p}.proj_in", c, c)
    for k in range(depth):
        t = f"{p}.transformer_blocks.{k}"
        _norm(s, f"{t}.norm1", c)
        for n in ("to_q", "to_k", "to_v"):
            _lin(s, f"{t}.attn1.{n}", c, c, bias=False)
        _lin(s, f"{t}.attn1.to_out.0", c, c)
        _norm(s, f"{t}.norm2", c)
        _lin(s, f"{t}.attn2.to_q", c, c, bias=False)
        _lin(s, f"{t}.attn2.to_k", c, ctx_dim, bias=False)
        _lin(s, f"{t}.attn2.to_v", c, ctx_dim, bias=False)
        _lin(s, f"{t}.attn2.to_out.0", c, c)
        _norm(s, f"{t}.norm3", c)
        _lin(s, f"{t}.ff.net.0.proj", 8 * c, c)
        _lin(s, f"{t}.ff.net.2", c, 4 * c)
    _lin(s, f"{p}.proj_out", c, c)


def _unet_encoder(s, cfg: UNetConfig):
    ch = cfg.block_out_channels
    temb = cfg.time_embed_dim
    _conv(s, "conv_in", ch[0], cfg.in_channels, 3)
    _lin(s, "time_embedding.linear_1", temb, ch[0])
    _lin(s, "time_embedding.linear_2", temb, temb)
    if cfg.addition_embed:      # SDXL text_time conditioning (add_time_proj is a parameter-free sinusoid)
        _lin(s, "add_embedding.linear_1", temb, cfg.projection_input_dim)
        _lin(s, "add_embedding.linear_2", temb, temb)
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(cfg.layers_per_block):
            _resnet(s, f"down_blocks.{i}.resnets.{j}", cin, cout, temb)
            if cfg.attn_levels[i]:
                _transformer2d(s, f"down_blocks.{i}.attentions.{j}", cout, cfg.cross_attention_dim, cfg.tf_layers(i))
            cin = cout
        if i < len(ch) - 1:
            _conv(s, f"down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
    _resnet(s, "mid_block.resnets.0", ch[-1], ch[-1], temb)
    _transformer2d(s, "mid_block.attentions.0", ch[-1], cfg.cross_attention_dim, cfg.tf_layers(len(ch) - 1))
    _resnet(s, "mid_block.resnets.1", ch[-1], ch[-1], temb)


def unet_skip_channels(cfg: UNetConfig):
    """Channel count of each of the 12 skip tensors S0..S11 the encoder pushes."""
    ch = cfg.block_out_channels
    skips = [ch[0]]
    for i, cout in enumerate(ch):
        skips += [cout] * cfg.layers_per_block
        if i < len(ch) - 1:
            skips.append(cout)
    return skips


def unet_shapes(cfg: UNetConfig) -> Shapes:
    """diffusers UNet2DConditionModel state-dict schema."""
    s: Shapes = OrderedDict()
    _unet_encoder(s, cfg)
    ch = cfg.block_out_channels
    temb = cfg.time_embed_dim
    skips = unet_skip_channels(cfg)
    rev = list(reversed(ch))
    prev = ch[-1]
    for i, cout in enumerate(rev):
        level = len(ch) - 1 - i
        for j in range(cfg.layers_per_block + 1):
            skip = skips.pop()
            _resnet(s, f"up_blocks.{i}.resnets.{j}", prev + skip, cout, temb)
            if cfg.attn_levels[level]:
                _transformer2d(s, f"up_blocks.{i}.attentions.{j}", cout, cfg.cross_attention_dim, cfg.tf_layers(level))
            prev = cout
        if i < len(ch) - 1:
            _conv(s, f"up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    _norm(s, "conv_norm_out", ch[0])
    _conv(s, "conv_out", cfg.out_channels, ch[0], 3)
    return s


def controlnet_shapes(cfg: UNetConfig) -> Shapes:
    """diffusers ControlNetModel state-dict schema (ControlNetModel.from_unet mirrors the U-Net encoder)."""
    s: Shapes = OrderedDict()
    _unet_encoder(s, cfg)
    ce = cfg.cond_embed_channels
    _conv(s, "controlnet_cond_embedding.conv_in", ce[0], 3, 3)
    k = 0
    for i in range(len(ce) - 1):
        _conv(s, f"controlnet_cond_embedding.blocks.{k}", ce[i], ce[i], 3)
        _conv(s, f"controlnet_cond_embedding.blocks.{k + 1}", ce[i + 1], ce[i], 3)  # stride 2
        k += 2
    _conv(s, "controlnet_cond_embedding.conv_out", cfg.block_out_channels[0], ce[-1], 3)
    for i, c in enumerate(unet_skip_channels(cfg)):
        _conv(s, f"controlnet_down_blocks.{i}", c, c, 1)
    _conv(s, "controlnet_mid_block", cfg.block_out_channels[-1], cfg.block_out_channels[-1], 1)
    return s


def vae_decoder_shapes(cfg: VAEConfig) -> Shapes:
    """diffusers AutoencoderKL: post_quant_conv + decoder.* keys."""
    s: Shapes = OrderedDict()
    ch = cfg.block_out_channels
    top = ch[-1]
    _conv(s, "post_quant_conv", cfg.latent_channels, cfg.latent_channels, 1)
    _conv(s, "decoder.conv_in", top, cfg.latent_channels, 3)
    _resnet(s, "decoder.mid_block.resnets.0", top, top, 0)
    a = "decoder.mid_block.attentions.0"
    _norm(s, f"{a}.group_norm", top)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        _lin(s, f"{a}.{n}", top, top)
    _resnet(s, "decoder.mid_block.resnets.1", top, top, 0)
    prev = top
    for i, cout in enumerate(reversed(ch)):
        for j in range(cfg.layers_per_block + 1):
            _resnet(s, f"decoder.up_blocks.{i}.resnets.{j}", prev, cout, 0)
            prev = cout
        if i < len(ch) - 1:
            _conv(s, f"decoder.up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    _norm(s, "decoder.conv_norm_out", ch[0])
    _conv(s, "decoder.conv_out", cfg.out_channels, ch[0], 3)
    return s


def vae_encoder_shapes(cfg: VAEConfig) -> Shapes:
    """diffusers AutoencoderKL: encoder.* + quant_conv keys (used by StableDiffusionInstructPix2PixPipeline's
    prepare_image_latents; controller/agent/sd_pix2pix_agent.py:36-41 loads the full VAE from sd_ckpt)."""
    s: Shapes = OrderedDict()
    ch = cfg.block_out_channels
    top = ch[-1]
    _conv(s, "encoder.conv_in", ch[0], cfg.out_channels, 3)
    prev = ch[0]
    for i, cout in enumerate(ch):
        for j in range(cfg.layers_per_block):
            _resnet(s, f"encoder.down_blocks.{i}.resnets.{j}", prev, cout, 0)
            prev = cout
        if i < len(ch) - 1:
            _conv(s, f"encoder.down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
    _resnet(s, "encoder.mid_block.resnets.0", top, top, 0)
    a = "encoder.mid_block.attentions.0"
    _norm(s, f"{a}.group_norm", top)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        _lin(s, f"{a}.{n}", top, top)
    _resnet(s, "encoder.mid_block.resnets.1", top, top, 0)
    _norm(s, "encoder.conv_norm_out", top)
    _conv(s, "encoder.conv_out", 2 * cfg.latent_channels, top, 3)
    _conv(s, "quant_conv", 2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
    return s


def taesd_layer_plan(cfg):
    """DecoderTiny's nn.Sequential as [(kind, index)]: kind in conv_in / relu / block / up / conv / conv_out; index = the
    position in `decoder.layers` (the state-dict key prefix)."""
    plan = [("conv_in", 0), ("relu", 1)]
    idx = 2
    n = len(cfg.num_blocks)
    for i, nb in enumerate(cfg.num_blocks):
        for _ in range(nb):
            plan.append(("block", idx))
            idx += 1
        if i < n - 1:
            plan.append(("up", idx))
            idx += 1
            plan.append(("conv", idx))
        else:
            plan.append(("conv_out", idx))
        idx += 1
    return plan


def taesd_decoder_shapes(cfg) -> Shapes:
    """diffusers AutoencoderTiny: decoder.layers.{i}[.conv.{0,2,4}].{weight,bias} (the 64 -> 64 convolutions that follow an
    Upsample have no bias; AutoencoderTinyBlock.skip is an Identity because in == out channels)."""
    s: Shapes = OrderedDict()
    c = cfg.channels
    for kind, i in taesd_layer_plan(cfg):
        p = f"decoder.layers.{i}"
        if kind == "conv_in":
            _conv(s, p, c, cfg.latent_channels, 3)
        elif kind == "block":
            for j in (0, 2, 4):
                _conv(s, f"{p}.conv.{j}", c, c, 3)
        elif kind == "conv":
            s[f"{p}.weight"] = (c, c, 3, 3)
        elif kind == "conv_out":
            _conv(s, p, cfg.out_channels, c, 3)
    return s


def clip_text_shapes(cfg: CLIPTextConfig) -> Shapes:
    """transformers CLIPTextModel schema (OpenAI clip weights map 1:1 onto it; text_projection is [proj, hidden])."""
    s: Shapes = OrderedDict()
    d = cfg.hidden_size
    s["text_model.embeddings.token_embedding.weight"] = (cfg.vocab_size, d)
    s["text_model.embeddings.position_embedding.weight"] = (cfg.max_positions, d)
    for i in range(cfg.num_layers):
        p = f"text_model.encoder.layers.{i}"
        _norm(s, f"{p}.layer_norm1", d)
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            _lin(s, f"{p}.self_attn.{n}", d, d)
        _norm(s, f"{p}.layer_norm2", d)
        _lin(s, f"{p}.mlp.fc1", cfg.intermediate_size, d)
        _lin(s, f"{p}.mlp.fc2", d, cfg.intermediate_size)
    _norm(s, "text_model.final_layer_norm", d)
    if cfg.projection_dim:
        s["text_projection.weight"] = (cfg.projection_dim, d)
    return s


def act_shapes(cfg: ACTConfig) -> Shapes:
    """GenimaACTPolicy (RoboBase ACT) schema: torchvision resnet18 trunk with FrozenBatchNorm2d + FiLM, DETR-style
    transformer (torch.nn.MultiheadAttention packing).  RoboBase's exact key names are unknown offline;
    these follow torchvision / DETR naming under the prefixes the survey expects."""
    s: Shapes = OrderedDict()
    w = cfg.resnet_widths
    d = cfg.hidden_dim
    b = "encoder_model.backbone"
    s[f"{b}.conv1.weight"] = (w[0], 3, 7, 7)

    def bn(name, c):
        for k in ("weight", "bias", "running_mean", "running_var"):
            s[f"{name}.{k}"] = (c,)

    bn(f"{b}.bn1", w[0])
    cin = w[0]
    for li, cout in enumerate(w):
        for bi in range(2):
            p = f"{b}.layer{li + 1}.{bi}"
            stride = 2 if (li > 0 and bi == 0) else 1
            s[f"{p}.conv1.weight"] = (cout, cin, 3, 3)
            bn(f"{p}.bn1", cout)
            s[f"{p}.conv2.weight"] = (cout, cout, 3, 3)
            bn(f"{p}.bn2", cout)
            _lin(s, f"{p}.film", 2 * cout, cfg.task_emb_dim)
            if stride != 1 or cin != cout:
                s[f"{p}.downsample.0.weight"] = (cout, cin, 1, 1)
                bn(f"{p}.downsample.1", cout)
            cin = cout
    _conv(s, "encoder_model.input_proj", d, w[-1], 1)
    a = "actor_model"

    def mha(name):
        s[f"{name}.in_proj_weight"] = (3 * d, d)
        s[f"{name}.in_proj_bias"] = (3 * d,)
        _lin(s, f"{name}.out_proj", d, d)

    for i in range(cfg.enc_layers):
        p = f"{a}.transformer.encoder.layers.{i}"
        mha(f"{p}.self_attn")
        _lin(s, f"{p}.linear1", cfg.dim_feedforward, d)
        _lin(s, f"{p}.linear2", d, cfg.dim_feedforward)
        _norm(s, f"{p}.norm1", d)
        _norm(s, f"{p}.norm2", d)
    for i in range(cfg.dec_layers):
        p = f"{a}.transformer.decoder.layers.{i}"
        mha(f"{p}.self_attn")
        mha(f"{p}.multihead_attn")
        _lin(s, f"{p}.linear1", cfg.dim_feedforward, d)
        _lin(s, f"{p}.linear2", d, cfg.dim_feedforward)
        for n in ("norm1", "norm2", "norm3"):
            _norm(s, f"{p}.{n}", d)
    _norm(s, f"{a}.transformer.decoder.norm", d)
    s[f"{a}.query_embed.weight"] = (cfg.num_queries, d)
    s[f"{a}.additional_pos_embed.weight"] = (2, d)
    _lin(s, f"{a}.input_proj_robot_state.0", d, cfg.state_dim)
    _lin(s, f"{a}.input_proj_robot_state.2", d, d)
    _lin(s, f"{a}.latent_out_proj", d, cfg.latent_dim)
    _lin(s, f"{a}.action_head", cfg.action_dim, d)
    _lin(s, f"{a}.is_pad_head", 1, d)
    return s


def count_params(shapes: Shapes, exclude=("running_mean", "running_var")) -> int:
    n = 0
    for k, shp in shapes.items():
        if any(k.endswith(e) for e in exclude):
            continue
        p = 1
        for v in shp:
            p *= v
        n += p
    return n


# --------------------------------------------------------------------------------------------------- synthetic weights
def _seed(name: str, salt: int) -> int:
    return (zlib.crc32(name.encode()) ^ (salt * 0x9E3779B1)) & 0x7FFFFFFF


def synth_tensor(name: str, shape: Tuple[int, ...], salt: int = 0) -> torch.Tensor:
    """Deterministic fp16-rounded tensor for key `name` (returned as fp16 on CPU)."""
    g = torch.Generator().manual_seed(_seed(name, salt))
    leaf = name.rsplit(".", 1)[-1]
    parent = name.rsplit(".", 1)[0]
    if leaf == "running_var":
        t = 1.0 + 0.1 * torch.randn(shape, generator=g).abs()
    elif leaf == "running_mean":
        t = 0.1 * torch.randn(shape, generator=g)
    elif leaf in ("bias", "in_proj_bias"):
        t = 0.02 * torch.randn(shape, generator=g)
    elif leaf == "weight" and len(shape) == 1:
        t = 1.0 + 0.02 * torch.randn(shape, generator=g)  # norm scales
    elif parent.rsplit(".", 1)[-1] in ("token_embedding", "position_embedding"):
        t = 0.02 * torch.randn(shape, generator=g)        # CLIP embedding tables
    elif parent.rsplit(".", 1)[-1] in ("query_embed", "additional_pos_embed"):
        t = torch.randn(shape, generator=g)               # nn.Embedding default init (DETR / ACT learned positions)
    else:
        fan_in = 1
        for v in shape[1:]:
            fan_in *= v
        t = torch.randn(shape, generator=g) * (fan_in ** -0.5)
        if parent.endswith("film"):
            t = t * 0.2
    return t.to(torch.float16)


def synth_state_dict(shapes: Shapes, salt: int = 0) -> Dict[str, torch.Tensor]:
    return OrderedDict((k, synth_tensor(k, shp, salt)) for k, shp in shapes.items())
