"""The oracle's OWN statement of the published architectures on Genima's hot path (TEST INFRASTRUCTURE — see
oracle/__init__).  Nothing here imports the product package: a topology or schema mistake in genima_b200/configs.py or
genima_b200/weights.py can therefore not hide by being shared with the checker.

Each `*_CONFIG_JSON` literal is the upstream `config.json` of the snapshot the reference loads
(controller/agent/sd_controlnet_agent.py:31-42 -> stabilityai/sd-turbo; controller/agent/sdxl_controlnet_agent.py ->
stabilityai/sdxl-turbo; controller/cfgs/method/genima_act.yaml:13-39 for ACT) [upstream, from memory — SURVEY.md
Appendices B, C, E, F], and `from_json` builds the plain namespace objects the oracle graph functions read (the same
attribute names as the product's dataclasses: the graph functions are duck-typed).

Independent check (tests/test_oracle_pins.py): `count_*` below are CLOSED-FORM parameter counts written from the
architecture description, not from any key table.  They must reproduce the published model sizes — U-Net 865,910,724,
ControlNet 364,228,240 (SD-2.1 topology), VAE decoder 49,490,199, OpenCLIP-H text 340,387,840 — and the oracle's graph
functions must touch exactly that many parameters when traced with a recording state dict.
"""
from __future__ import annotations

from types import SimpleNamespace

# stabilityai/sd-turbo/unet/config.json (SD-2.1-base topology)
SD_TURBO_UNET_JSON = {
    "in_channels": 4, "out_channels": 4, "block_out_channels": [320, 640, 1280, 1280], "layers_per_block": 2,
    "attention_head_dim": [5, 10, 20, 20],     # SD-2.x: these are HEAD COUNTS (head_dim = 64)
    "down_block_types": ["CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"],
    "cross_attention_dim": 1024, "norm_num_groups": 32, "norm_eps": 1e-5, "use_linear_projection": True,
    "sample_size": 64,
}
# ControlNetModel.from_unet (diffusion/train_controlnet_genima.py:1071): conditioning_embedding_out_channels default
CONTROLNET_COND_CHANNELS = [16, 32, 96, 256]
# stabilityai/sd-turbo/vae/config.json
SD_TURBO_VAE_JSON = {"latent_channels": 4, "out_channels": 3, "block_out_channels": [128, 256, 512, 512],
                     "layers_per_block": 2, "norm_num_groups": 32, "scaling_factor": 0.18215}
# stabilityai/sd-turbo/text_encoder/config.json (OpenCLIP ViT-H/14 text tower, penultimate-layer trick baked in: 23 layers)
SD_TURBO_TEXT_JSON = {"vocab_size": 49408, "hidden_size": 1024, "intermediate_size": 4096, "num_hidden_layers": 23,
                      "num_attention_heads": 16, "max_position_embeddings": 77, "hidden_act": "gelu",
                      "layer_norm_eps": 1e-5}
# OpenAI CLIP ViT-B/32 text tower (clip.load("ViT-B/32"), controller/method/genima_act.py:315-321)
CLIP_VIT_B32_TEXT = {"vocab_size": 49408, "hidden_size": 512, "intermediate_size": 2048, "num_hidden_layers": 12,
                     "num_attention_heads": 8, "max_position_embeddings": 77, "hidden_act": "quick_gelu",
                     "layer_norm_eps": 1e-5, "projection_dim": 512}
# stabilityai/sdxl-turbo/unet/config.json
SDXL_UNET_JSON = {
    "in_channels": 4, "out_channels": 4, "block_out_channels": [320, 640, 1280], "layers_per_block": 2,
    "attention_head_dim": [5, 10, 20], "down_block_types": ["DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"],
    "cross_attention_dim": 2048, "norm_num_groups": 32, "norm_eps": 1e-5, "use_linear_projection": True,
    "transformer_layers_per_block": [1, 2, 10], "addition_embed_type": "text_time", "addition_time_embed_dim": 256,
    "projection_class_embeddings_input_dim": 2816, "sample_size": 128,
}
# controller/cfgs/method/genima_act.yaml:13-39 + RoboBase ACT defaults (latent_dim 32 [U]); 4 cameras x 256^2
GENIMA_ACT = {"hidden_dim": 256, "enc_layers": 4, "dec_layers": 6, "dim_feedforward": 2048, "nheads": 8,
              "num_queries": 20, "state_dim": 8, "action_dim": 8, "latent_dim": 32, "num_views": 4, "image_size": 256,
              "task_emb_dim": 512, "resnet_widths": [64, 128, 256, 512], "bn_eps": 1e-5, "ln_eps": 1e-5}


def unet_from_json(j: dict, cond_channels=CONTROLNET_COND_CHANNELS) -> SimpleNamespace:
    ch = tuple(j["block_out_channels"])
    tl = tuple(j.get("transformer_layers_per_block", ()))
    if isinstance(j.get("transformer_layers_per_block"), int):
        tl = ()
    ns = SimpleNamespace(
        in_channels=j["in_channels"], out_channels=j["out_channels"], block_out_channels=ch,
        layers_per_block=j["layers_per_block"], num_heads=tuple(j["attention_head_dim"]),
        attn_levels=tuple(t.startswith("CrossAttn") for t in j["down_block_types"]),
        cross_attention_dim=j["cross_attention_dim"], norm_num_groups=j["norm_num_groups"], norm_eps=j["norm_eps"],
        cond_embed_channels=tuple(cond_channels), sample_size=j["sample_size"], transformer_layers=tl,
        addition_embed=j.get("addition_embed_type") == "text_time",
        addition_time_embed_dim=j.get("addition_time_embed_dim", 256),
        projection_input_dim=j.get("projection_class_embeddings_input_dim", 2816))
    ns.tf_layers = lambda level: ns.transformer_layers[level] if ns.transformer_layers else 1
    ns.time_embed_dim = ch[0] * 4
    return ns


def vae_from_json(j: dict) -> SimpleNamespace:
    return SimpleNamespace(latent_channels=j["latent_channels"], out_channels=j["out_channels"],
                           block_out_channels=tuple(j["block_out_channels"]), layers_per_block=j["layers_per_block"],
                           norm_num_groups=j["norm_num_groups"], norm_eps=1e-6, scaling_factor=j["scaling_factor"],
                           force_upcast=bool(j.get("force_upcast", False)))


def text_from_json(j: dict) -> SimpleNamespace:
    return SimpleNamespace(vocab_size=j["vocab_size"], hidden_size=j["hidden_size"],
                           intermediate_size=j["intermediate_size"], num_layers=j["num_hidden_layers"],
                           num_heads=j["num_attention_heads"], max_positions=j["max_position_embeddings"],
                           act=j["hidden_act"], eps=j["layer_norm_eps"], projection_dim=j.get("projection_dim", 0))


def act_from_dict(d: dict) -> SimpleNamespace:
    return SimpleNamespace(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in d.items()})


def sd_turbo():
    """(unet / controlnet cfg, vae cfg, text cfg) of the snapshot the reference evaluates with."""
    return unet_from_json(SD_TURBO_UNET_JSON), vae_from_json(SD_TURBO_VAE_JSON), text_from_json(SD_TURBO_TEXT_JSON)


def taesd_layer_plan(cfg):
    """DecoderTiny's nn.Sequential (diffusers 0.29.0 models/autoencoders/vae.py [U]) as [(kind, index in
    `decoder.layers`)]: conv_in, ReLU, then per stage `num_blocks[i]` AutoencoderTinyBlocks followed by Upsample + conv
    (no bias), the last stage ending in conv_out."""
    plan = [("conv_in", 0), ("relu", 1)]
    idx = 2
    last = len(cfg.num_blocks) - 1
    for i, nb in enumerate(cfg.num_blocks):
        plan += [("block", idx + k) for k in range(nb)]
        idx += nb
        if i < last:
            plan += [("up", idx), ("conv", idx + 1)]
            idx += 2
        else:
            plan.append(("conv_out", idx))
            idx += 1
    return plan


# ---------------------------------------------------------------------------------------------- closed-form counts
def _conv(cin, cout, k, bias=True):
    return cout * cin * k * k + (cout if bias else 0)


def _lin(cin, cout, bias=True):
    return cout * cin + (cout if bias else 0)


def _resnet(cin, cout, temb):
    n = 2 * cin + _conv(cin, cout, 3) + 2 * cout + _conv(cout, cout, 3)          # norm1, conv1, norm2, conv2
    if temb:
        n += _lin(temb, cout)
    if cin != cout:
        n += _conv(cin, cout, 1)                                                  # conv_shortcut
    return n


def _basic_transformer_block(c, ctx):
    attn1 = 3 * _lin(c, c, bias=False) + _lin(c, c)
    attn2 = _lin(c, c, bias=False) + 2 * _lin(ctx, c, bias=False) + _lin(c, c)
    ff = _lin(c, 8 * c) + _lin(4 * c, c)                                          # GEGLU proj + out
    return attn1 + attn2 + ff + 3 * 2 * c                                         # three LayerNorms


def _transformer2d(c, ctx, depth):
    return 2 * c + 2 * _lin(c, c) + depth * _basic_transformer_block(c, ctx)      # GroupNorm, proj_in / proj_out (linear)


def _encoder_params(cfg, with_conv_in=True):
    """conv_in + time embedding (+ SDXL add_embedding) + down blocks + mid block: shared by U-Net and ControlNet."""
    ch, temb, ctx = cfg.block_out_channels, cfg.block_out_channels[0] * 4, cfg.cross_attention_dim
    n = _conv(cfg.in_channels, ch[0], 3) if with_conv_in else 0
    n += _lin(ch[0], temb) + _lin(temb, temb)
    if cfg.addition_embed:
        n += _lin(cfg.projection_input_dim, temb) + _lin(temb, temb)
    cin = ch[0]
    for i, cout in enumerate(ch):
        for _ in range(cfg.layers_per_block):
            n += _resnet(cin, cout, temb)
            if cfg.attn_levels[i]:
                n += _transformer2d(cout, ctx, cfg.tf_layers(i))
            cin = cout
        if i < len(ch) - 1:
            n += _conv(cout, cout, 3)                                             # Downsample2D
    n += 2 * _resnet(ch[-1], ch[-1], temb) + _transformer2d(ch[-1], ctx, cfg.tf_layers(len(ch) - 1))
    return n


def skip_channels(cfg):
    ch = cfg.block_out_channels
    out = [ch[0]]
    for i, c in enumerate(ch):
        out += [c] * cfg.layers_per_block
        if i < len(ch) - 1:
            out.append(c)
    return out


def count_unet(cfg) -> int:
    ch, temb, ctx = cfg.block_out_channels, cfg.block_out_channels[0] * 4, cfg.cross_attention_dim
    n = _encoder_params(cfg)
    skips = skip_channels(cfg)
    prev = ch[-1]
    for i, cout in enumerate(reversed(ch)):
        level = len(ch) - 1 - i
        for _ in range(cfg.layers_per_block + 1):
            n += _resnet(prev + skips.pop(), cout, temb)
            if cfg.attn_levels[level]:
                n += _transformer2d(cout, ctx, cfg.tf_layers(level))
            prev = cout
        if i < len(ch) - 1:
            n += _conv(cout, cout, 3)                                             # Upsample2D conv
    return n + 2 * ch[0] + _conv(ch[0], cfg.out_channels, 3)                      # conv_norm_out, conv_out


def count_controlnet(cfg) -> int:
    ch, ce = cfg.block_out_channels, cfg.cond_embed_channels
    n = _encoder_params(cfg)
    n += _conv(3, ce[0], 3)                                                       # cond embedding conv_in
    for a, b in zip(ce[:-1], ce[1:]):
        n += _conv(a, a, 3) + _conv(a, b, 3)                                      # (same, stride-2) pairs
    n += _conv(ce[-1], ch[0], 3)                                                  # zero-initialised conv_out
    n += sum(_conv(c, c, 1) for c in skip_channels(cfg)) + _conv(ch[-1], ch[-1], 1)   # zero convs
    return n


def count_vae_decoder(cfg) -> int:
    ch, lc = cfg.block_out_channels, cfg.latent_channels
    top = ch[-1]
    n = _conv(lc, lc, 1) + _conv(lc, top, 3)                                      # post_quant_conv, conv_in
    n += 2 * _resnet(top, top, 0) + 2 * top + 4 * _lin(top, top)                  # mid: 2 resnets + 1-head attention
    prev = top
    for i, cout in enumerate(reversed(ch)):
        for _ in range(cfg.layers_per_block + 1):
            n += _resnet(prev, cout, 0)
            prev = cout
        if i < len(ch) - 1:
            n += _conv(cout, cout, 3)
    return n + 2 * ch[0] + _conv(ch[0], cfg.out_channels, 3)


def count_clip_text(cfg) -> int:
    d, ff = cfg.hidden_size, cfg.intermediate_size
    layer = 4 * _lin(d, d) + _lin(d, ff) + _lin(ff, d) + 2 * 2 * d
    n = cfg.vocab_size * d + cfg.max_positions * d + cfg.num_layers * layer + 2 * d
    return n + (cfg.projection_dim * d if cfg.projection_dim else 0)
