"""On-disk weight formats of the reference -> plain {name: tensor} state dicts for the device graphs.

  * diffusers directories as read by controller/agent/sd_controlnet_agent.py:19-42:
      <diffusion_ckpt>/checkpoint-<N>/controlnet/{config.json, diffusion_pytorch_model.safetensors}   (last N, natural sort)
      <sd_ckpt>/{unet,vae,text_encoder}/...(.fp16).safetensors, <sd_ckpt>/scheduler/scheduler_config.json
  * RoboBase controller snapshots written by controller/train_act.py:262-279 and read by
    controller/eval_genima.py:91-103: torch.save({... "agent": state_dict minus clip_model.*}) as latest.pt / <N>.pt.
No checkpoint is available offline, so everything else in this repo runs on `synthetic_weights`
(genima_b200/weights.py); these loaders exist so that a real checkpoint binds through the same schema tables and are
exercised on round-tripped synthetic checkpoints in tests/test_checkpoint.py.
"""
from __future__ import annotations

import dataclasses
import json
import os
import re
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch

from . import weights as W
from .configs import ACTConfig, CLIPTextConfig, SchedulerConfig, TAESDConfig, UNetConfig, VAEConfig


def _natural_key(s: str):
    return [int(t) if t.isdigit() else t.lower() for t in re.split(r"(\d+)", s)]


def find_controlnet_dir(diffusion_ckpt: str) -> str:
    """controller/agent/sd_controlnet_agent.py:19-29: the last `*checkpoint*` sub-directory in natural order, else the
    directory itself."""
    dirs = sorted((d for d in os.listdir(diffusion_ckpt) if "checkpoint" in d), key=_natural_key)
    if dirs:
        return os.path.join(diffusion_ckpt, dirs[-1], "controlnet")
    return diffusion_ckpt


def load_safetensors_dir(path: str, prefer_fp16: bool = True) -> Dict[str, torch.Tensor]:
    """Reads diffusion_pytorch_model[.fp16].safetensors / model[.fp16].safetensors from a diffusers component dir."""
    from safetensors.torch import load_file

    names = []
    for stem in ("diffusion_pytorch_model", "model"):
        if prefer_fp16:
            names.append(f"{stem}.fp16.safetensors")
        names.append(f"{stem}.safetensors")
    for n in names:
        f = os.path.join(path, n)
        if os.path.exists(f):
            return load_file(f)
    raise FileNotFoundError(f"no safetensors weight file in {path} (looked for {names})")


def _read_json(path: str) -> dict:
    with open(path) as f:
        return json.load(f)


def unet_config_from_json(cfg: dict) -> UNetConfig:
    """diffusers unet/config.json or controlnet/config.json -> UNetConfig; rejects topologies the graphs do not cover."""
    boc = tuple(cfg.get("block_out_channels", (320, 640, 1280, 1280)))
    down = cfg.get("down_block_types", ("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",))
    heads = cfg.get("attention_head_dim", (5, 10, 20, 20))
    if isinstance(heads, int):
        heads = (heads,) * len(boc)
    if not cfg.get("use_linear_projection", True):
        raise NotImplementedError("use_linear_projection=False (conv proj_in/out, SD-1.x) is not implemented")
    if cfg.get("upcast_attention", False):
        raise NotImplementedError("upcast_attention=True (SD-2.1-768) is not implemented")
    if cfg.get("class_embed_type") or cfg.get("addition_embed_type") not in (None, "text_time"):
        raise NotImplementedError("class embeddings / addition_embed_type other than SDXL's 'text_time' are not implemented")
    sdxl = cfg.get("addition_embed_type") == "text_time"
    tl = cfg.get("transformer_layers_per_block", 1)
    tl = () if tl == 1 else tuple((tl,) * len(boc) if isinstance(tl, int) else tl)
    ce = tuple(cfg.get("conditioning_embedding_out_channels", (16, 32, 96, 256)))
    return UNetConfig(transformer_layers=tl, addition_embed=sdxl,
                      addition_time_embed_dim=cfg.get("addition_time_embed_dim") or 256,
                      projection_input_dim=cfg.get("projection_class_embeddings_input_dim") or 2816,
                      in_channels=cfg.get("in_channels", 4), out_channels=cfg.get("out_channels", 4),
                      block_out_channels=boc, layers_per_block=cfg.get("layers_per_block", 2),
                      num_heads=tuple(heads), attn_levels=tuple("CrossAttn" in t for t in down),
                      cross_attention_dim=cfg.get("cross_attention_dim", 1024),
                      norm_num_groups=cfg.get("norm_num_groups", 32), norm_eps=cfg.get("norm_eps", 1e-5),
                      cond_embed_channels=ce, sample_size=cfg.get("sample_size", 64))


# scheduler_config.json keys that change the ODE and are NOT implemented: a non-default value raises instead of being
# silently dropped (diffusers 0.29.0 defaults in the right column)
_UNSUPPORTED_SCHEDULER_FLAGS = {
    "use_karras_sigmas": False, "interpolation_type": "linear", "rescale_betas_zero_snr": False,
    "timestep_type": "discrete", "final_sigmas_type": "zero", "sigma_min": None, "sigma_max": None,
    "thresholding": False, "trained_betas": None,
}


def scheduler_config_from_json(cfg: dict) -> SchedulerConfig:
    cls = cfg.get("_class_name", "EulerDiscreteScheduler")
    for k, default in _UNSUPPORTED_SCHEDULER_FLAGS.items():
        if k in cfg and cfg[k] != default:
            raise NotImplementedError(f"scheduler_config.json: {k}={cfg[k]!r} is not implemented (only {default!r})")
    # diffusers' own defaults when the key is absent: DDIM 'leading'; EulerDiscrete / EulerAncestral 'linspace'
    # (the sd-turbo / sdxl-turbo snapshots the reference loads say 'trailing' explicitly)
    default_spacing = "leading" if cls == "DDIMScheduler" else "linspace"
    return SchedulerConfig(class_name=cls,
                           num_train_timesteps=cfg.get("num_train_timesteps", 1000),
                           beta_start=cfg.get("beta_start", 0.00085),
                           beta_end=cfg.get("beta_end", 0.012),
                           beta_schedule=cfg.get("beta_schedule", "scaled_linear"),
                           timestep_spacing=cfg.get("timestep_spacing", default_spacing),
                           prediction_type=cfg.get("prediction_type", "epsilon"),
                           steps_offset=cfg.get("steps_offset", 0),
                           set_alpha_to_one=cfg.get("set_alpha_to_one", True),
                           # diffusers' DDIMScheduler clips the predicted x0 by default; SD snapshots switch it off
                           clip_sample=bool(cfg.get("clip_sample", cls == "DDIMScheduler")))


def check_schema(sd: Dict[str, torch.Tensor], shapes, what: str, allow_extra: Tuple[str, ...] = ()) -> None:
    """Every key of the schema must be present with the expected shape — a wrong architecture fails at load time."""
    missing = [k for k in shapes if k not in sd]
    bad = [k for k, shp in shapes.items() if k in sd and tuple(sd[k].shape) != tuple(shp)]
    if missing or bad:
        raise ValueError(f"{what}: {len(missing)} missing keys (e.g. {missing[:3]}), "
                         f"{len(bad)} shape mismatches (e.g. {[(k, tuple(sd[k].shape), shapes[k]) for k in bad[:3]]})")


def find_pix2pix_unet_dir(diffusion_ckpt: str) -> str:
    """controller/agent/sd_pix2pix_agent.py:19-33: the last `*checkpoint*` sub-directory in natural order (else the
    directory itself), sub-folder `unet`."""
    dirs = sorted((d for d in os.listdir(diffusion_ckpt) if "checkpoint" in d), key=_natural_key)
    return os.path.join(os.path.join(diffusion_ckpt, dirs[-1]) if dirs else diffusion_ckpt, "unet")


def load_sd_pix2pix(sd_ckpt: str, diffusion_ckpt: str):
    """-> dict(unet=, vae=, text=, unet_cfg=, vae_cfg=, text_cfg=, scheduler_cfg=): the fine-tuned 8-channel U-Net from
    `<diffusion_ckpt>/checkpoint-*/unet`, everything else (full VAE with its encoder, text encoder, scheduler) from
    sd_ckpt — what StableDiffusionInstructPix2PixPipeline.from_pretrained(sd_ckpt, unet=unet) assembles
    (controller/agent/sd_pix2pix_agent.py:29-41)."""
    out = load_sd_turbo(sd_ckpt, None)
    udir = find_pix2pix_unet_dir(diffusion_ckpt)
    ucfg = unet_config_from_json(_read_json(os.path.join(udir, "config.json")))
    if ucfg.in_channels != 2 * out["vae_cfg"].latent_channels:
        raise ValueError(f"InstructPix2Pix U-Net must take {2 * out['vae_cfg'].latent_channels} input channels, "
                         f"{udir} has {ucfg.in_channels}")
    out["unet"], out["unet_cfg"] = load_safetensors_dir(udir), ucfg
    check_schema(out["unet"], W.unet_shapes(ucfg), "InstructPix2Pix U-Net")
    check_schema(out["vae"], W.vae_encoder_shapes(out["vae_cfg"]), "VAE encoder")
    return out


def load_clip_tokenizer(path: str):
    """`<sd_ckpt>/tokenizer` (vocab.json + merges.txt [+ tokenizer_config.json]) -> callable list[str] -> ids [B, 77] int64,
    i.e. what diffusers' encode_prompt does with `self.tokenizer(prompt, padding="max_length", max_length=
    tokenizer.model_max_length, truncation=True, return_tensors="pt").input_ids`.  Uses transformers' CLIPTokenizer on
    local files only; returns None when the directory (or transformers) is missing — string prompts then need an explicit
    `tokenizer=` callable (no CLIP vocabulary ships with this repository)."""
    if not (os.path.isfile(os.path.join(path, "vocab.json")) and os.path.isfile(os.path.join(path, "merges.txt"))):
        return None
    try:
        from transformers import CLIPTokenizer
    except Exception:  # pragma: no cover
        return None
    tok = CLIPTokenizer.from_pretrained(path, local_files_only=True)
    max_len = min(int(getattr(tok, "model_max_length", 77) or 77), 77)

    def encode(prompts):
        return tok(list(prompts), padding="max_length", max_length=max_len, truncation=True,
                   return_tensors="pt").input_ids.to(torch.int64)

    return encode


def text_config_from_json(tj: dict) -> CLIPTextConfig:
    return CLIPTextConfig(vocab_size=tj.get("vocab_size", 49408), hidden_size=tj.get("hidden_size", 1024),
                          intermediate_size=tj.get("intermediate_size", 4096),
                          num_layers=tj.get("num_hidden_layers", 23), num_heads=tj.get("num_attention_heads", 16),
                          max_positions=tj.get("max_position_embeddings", 77),
                          act="quick_gelu" if tj.get("hidden_act", "gelu") == "quick_gelu" else "gelu",
                          eps=tj.get("layer_norm_eps", 1e-5),
                          projection_dim=(tj.get("projection_dim", 0)
                                          if "WithProjection" in str(tj.get("architectures", "")) else 0))


def load_sdxl(sd_ckpt: str, diffusion_ckpt: str):
    """load_sd_turbo for an SDXL snapshot (stabilityai/sdxl-turbo layout; controller/agent/sdxl_controlnet_agent.py:19-42):
    adds text_encoder_2 (CLIPTextModelWithProjection) -> out["text2"], out["text2_cfg"]."""
    out = load_sd_turbo(sd_ckpt, diffusion_ckpt)
    if not out["unet_cfg"].addition_embed:
        raise ValueError(f"{sd_ckpt} is not an SDXL snapshot (unet addition_embed_type != 'text_time')")
    t2 = os.path.join(sd_ckpt, "text_encoder_2")
    out["text2_cfg"] = text_config_from_json(_read_json(os.path.join(t2, "config.json")))
    if not out["text2_cfg"].projection_dim:
        raise ValueError("text_encoder_2 must be a CLIPTextModelWithProjection")
    out["tokenizer_2"] = load_clip_tokenizer(os.path.join(sd_ckpt, "tokenizer_2"))
    out["text2"] = load_safetensors_dir(t2)
    check_schema(out["text2"], W.clip_text_shapes(out["text2_cfg"]), "text encoder 2")
    return out


def load_sd_turbo(sd_ckpt: str, diffusion_ckpt: Optional[str]):
    """-> dict(unet=, controlnet=, vae=, text=, unet_cfg=, vae_cfg=, text_cfg=, scheduler_cfg=) from local directories
    (diffusion_ckpt None: the base components only, no ControlNet)."""
    if not os.path.isdir(sd_ckpt):
        raise FileNotFoundError(
            f"sd_ckpt {sd_ckpt!r} is not a local directory; hub ids cannot be resolved offline — point it at a local "
            "snapshot of stabilityai/sd-turbo (unet/, vae/, text_encoder/, scheduler/)")
    ucfg = unet_config_from_json(_read_json(os.path.join(sd_ckpt, "unet", "config.json")))
    cn_dir = find_controlnet_dir(diffusion_ckpt) if diffusion_ckpt is not None else None
    if cn_dir is not None:
        ccfg = unet_config_from_json(_read_json(os.path.join(cn_dir, "config.json")))
        if (ccfg.block_out_channels, ccfg.num_heads) != (ucfg.block_out_channels, ucfg.num_heads):
            raise ValueError("ControlNet and U-Net configurations do not match")
        ucfg = dataclasses.replace(ucfg, cond_embed_channels=ccfg.cond_embed_channels)
    vj = _read_json(os.path.join(sd_ckpt, "vae", "config.json"))
    vcfg = VAEConfig(latent_channels=vj.get("latent_channels", 4), out_channels=vj.get("out_channels", 3),
                     block_out_channels=tuple(vj.get("block_out_channels", (128, 256, 512, 512))),
                     layers_per_block=vj.get("layers_per_block", 2), norm_num_groups=vj.get("norm_num_groups", 32),
                     scaling_factor=vj.get("scaling_factor", 0.18215),
                     force_upcast=bool(vj.get("force_upcast", False)) and ucfg.addition_embed)   # (SDXL VAEs only: the
    #                SD-2.x VAE config also carries force_upcast=true by default but is fp16-safe and diffusers only
    #                acts on the flag in the SDXL pipelines)
    tj = _read_json(os.path.join(sd_ckpt, "text_encoder", "config.json"))
    tcfg = text_config_from_json(tj)
    scfg = scheduler_config_from_json(_read_json(os.path.join(sd_ckpt, "scheduler", "scheduler_config.json")))
    out = dict(tokenizer=load_clip_tokenizer(os.path.join(sd_ckpt, "tokenizer")),
               unet=load_safetensors_dir(os.path.join(sd_ckpt, "unet")),
               controlnet=load_safetensors_dir(cn_dir) if cn_dir is not None else None,
               vae=load_safetensors_dir(os.path.join(sd_ckpt, "vae")),
               text=load_safetensors_dir(os.path.join(sd_ckpt, "text_encoder")),
               unet_cfg=ucfg, vae_cfg=vcfg, text_cfg=tcfg, scheduler_cfg=scfg)
    check_schema(out["unet"], W.unet_shapes(ucfg), "U-Net")
    if cn_dir is not None:
        check_schema(out["controlnet"], W.controlnet_shapes(ucfg), "ControlNet")
    check_schema(out["vae"], W.vae_decoder_shapes(vcfg), "VAE decoder")
    check_schema(out["text"], W.clip_text_shapes(tcfg), "text encoder")
    return out


def load_taesd(path: str):
    """Local snapshot of an AutoencoderTiny checkpoint (madebyollin/taesd: config.json + diffusion_pytorch_model.safetensors)
    -> (decoder state dict, TAESDConfig).  The reference passes a hub id (sd_controlnet_agent.py:45-49); offline it must be
    a directory."""
    if not os.path.isdir(path):
        raise FileNotFoundError(f"autoencoder {path!r} is not a local directory; hub ids cannot be resolved offline — "
                                "point it at a local snapshot of madebyollin/taesd")
    j = _read_json(os.path.join(path, "config.json"))
    chans = j.get("decoder_block_out_channels", (64, 64, 64, 64))
    cfg = TAESDConfig(latent_channels=j.get("latent_channels", 4), out_channels=j.get("out_channels", 3),
                      channels=int(chans[0]), num_blocks=tuple(j.get("num_decoder_blocks", (3, 3, 3, 1))),
                      latent_magnitude=float(j.get("latent_magnitude", 3)), scaling_factor=float(j.get("scaling_factor", 1.0)))
    sd = {k: v for k, v in load_safetensors_dir(path).items() if k.startswith("decoder.")}
    check_schema(sd, W.taesd_decoder_shapes(cfg), "TAESD decoder")
    return sd, cfg


def load_controller_snapshot(path: str, cfg: ACTConfig = ACTConfig(), prefix: str = "actor.") -> Dict[str, torch.Tensor]:
    """RoboBase snapshot (controller/train_act.py:262-279) -> ACT state dict in the act_shapes schema.
    `payload["agent"]` holds the agent's state dict without clip_model.* keys; the policy lives under `actor.`."""
    payload = torch.load(path, map_location="cpu", weights_only=False)
    agent = payload["agent"] if "agent" in payload else payload
    sd = {k[len(prefix):]: v for k, v in agent.items() if k.startswith(prefix)}
    check_schema(sd, W.act_shapes(cfg), f"controller snapshot {path}")
    return sd


def openai_clip_text_to_hf(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Text tower of an OpenAI CLIP state dict (`clip.load("ViT-B/32")`, controller/method/genima_act.py:315-321; names as
    the reference reads them at :324-341: token_embedding, positional_embedding, transformer.resblocks, ln_final,
    text_projection) -> the transformers CLIPTextModelWithProjection names DeviceCLIPText binds.  OpenAI's
    `x @ text_projection` ([d, proj]) becomes `text_projection.weight` = its transpose ([proj, d])."""
    out: Dict[str, torch.Tensor] = {}
    out["text_model.embeddings.token_embedding.weight"] = sd["token_embedding.weight"]
    out["text_model.embeddings.position_embedding.weight"] = sd["positional_embedding"]
    i = 0
    while f"transformer.resblocks.{i}.attn.in_proj_weight" in sd:
        src, dst = f"transformer.resblocks.{i}", f"text_model.encoder.layers.{i}"
        w, b = sd[f"{src}.attn.in_proj_weight"], sd[f"{src}.attn.in_proj_bias"]
        d = w.shape[1]
        for j, n in enumerate("qkv"):
            out[f"{dst}.self_attn.{n}_proj.weight"] = w[j * d:(j + 1) * d]
            out[f"{dst}.self_attn.{n}_proj.bias"] = b[j * d:(j + 1) * d]
        for a, bname in (("attn.out_proj", "self_attn.out_proj"), ("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"),
                         ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
            out[f"{dst}.{bname}.weight"] = sd[f"{src}.{a}.weight"]
            out[f"{dst}.{bname}.bias"] = sd[f"{src}.{a}.bias"]
        i += 1
    if i == 0:
        raise ValueError("not an OpenAI CLIP state dict: no transformer.resblocks.* keys")
    out["text_model.final_layer_norm.weight"] = sd["ln_final.weight"]
    out["text_model.final_layer_norm.bias"] = sd["ln_final.bias"]
    out["text_projection.weight"] = sd["text_projection"].t().contiguous()
    return out


def load_openai_clip_text(path: str, cfg: CLIPTextConfig = CLIPTextConfig.vit_b32()) -> Dict[str, torch.Tensor]:
    """Local copy of the file `clip.load("ViT-B/32")` downloads (`~/.cache/clip/ViT-B-32.pt`: a TorchScript archive) or a
    plain state dict saved with torch.save -> text-tower state dict in transformers naming, schema-checked."""
    if not os.path.isfile(path):
        raise FileNotFoundError(f"CLIP checkpoint {path!r} not found; hub downloads are impossible offline")
    try:
        sd = torch.jit.load(path, map_location="cpu").state_dict()
    except RuntimeError:
        sd = torch.load(path, map_location="cpu", weights_only=False)
        if hasattr(sd, "state_dict"):
            sd = sd.state_dict()
    sd = {k: v for k, v in sd.items() if not k.startswith("visual.")}
    if "token_embedding.weight" in sd:
        sd = openai_clip_text_to_hf(sd)
    sd = {k: v for k, v in sd.items() if k in W.clip_text_shapes(cfg)}
    check_schema(sd, W.clip_text_shapes(cfg), f"CLIP text tower {path}")
    return sd


def save_synthetic_checkpoints(root: str, ucfg: UNetConfig, vcfg: VAEConfig, tcfg: CLIPTextConfig, acfg: ACTConfig,
                               text2_cfg: Optional[CLIPTextConfig] = None):
    """Writes seeded synthetic weights in the reference's on-disk layouts (for loader tests and offline demos).  With an
    SDXL `ucfg` (addition_embed) and `text2_cfg` the snapshot has the stabilityai/sdxl-turbo layout: text_encoder_2/ and an
    EulerAncestralDiscreteScheduler."""
    from safetensors.torch import save_file

    sd_ckpt = os.path.join(root, "sd-turbo")
    dif = os.path.join(root, "diffusion_ckpt")
    ctl = os.path.join(root, "controller_ckpt")
    heads = list(ucfg.num_heads)
    down = ["CrossAttnDownBlock2D" if a else "DownBlock2D" for a in ucfg.attn_levels]
    ujson = dict(_class_name="UNet2DConditionModel", in_channels=ucfg.in_channels, out_channels=ucfg.out_channels,
                 block_out_channels=list(ucfg.block_out_channels), layers_per_block=ucfg.layers_per_block,
                 attention_head_dim=heads, down_block_types=down, cross_attention_dim=ucfg.cross_attention_dim,
                 norm_num_groups=ucfg.norm_num_groups, norm_eps=ucfg.norm_eps, use_linear_projection=True,
                 sample_size=ucfg.sample_size)
    if ucfg.addition_embed:
        ujson.update(addition_embed_type="text_time", addition_time_embed_dim=ucfg.addition_time_embed_dim,
                     projection_class_embeddings_input_dim=ucfg.projection_input_dim,
                     transformer_layers_per_block=list(ucfg.transformer_layers))
    parts = [
        (os.path.join(sd_ckpt, "unet"), "diffusion_pytorch_model.fp16.safetensors", W.unet_shapes(ucfg), 0, ujson),
        (os.path.join(sd_ckpt, "vae"), "diffusion_pytorch_model.fp16.safetensors",
         OrderedDict(list(W.vae_decoder_shapes(vcfg).items()) + list(W.vae_encoder_shapes(vcfg).items())), 2,
         dict(_class_name="AutoencoderKL", latent_channels=vcfg.latent_channels,
              block_out_channels=list(vcfg.block_out_channels), layers_per_block=vcfg.layers_per_block,
              norm_num_groups=vcfg.norm_num_groups, scaling_factor=vcfg.scaling_factor)),
        (os.path.join(sd_ckpt, "text_encoder"), "model.fp16.safetensors", W.clip_text_shapes(tcfg), 0,
         dict(vocab_size=tcfg.vocab_size, hidden_size=tcfg.hidden_size, intermediate_size=tcfg.intermediate_size,
              num_hidden_layers=tcfg.num_layers, num_attention_heads=tcfg.num_heads,
              max_position_embeddings=tcfg.max_positions, hidden_act=tcfg.act, layer_norm_eps=tcfg.eps)),
        (os.path.join(dif, "checkpoint-500", "controlnet"), "diffusion_pytorch_model.safetensors",
         W.controlnet_shapes(ucfg), 7, ujson),     # an older checkpoint with different weights: must NOT be picked
        (os.path.join(dif, "checkpoint-1000", "controlnet"), "diffusion_pytorch_model.safetensors",
         W.controlnet_shapes(ucfg), 1,
         dict(ujson, _class_name="ControlNetModel",
              conditioning_embedding_out_channels=list(ucfg.cond_embed_channels))),
    ]
    if text2_cfg is not None:
        parts.append((os.path.join(sd_ckpt, "text_encoder_2"), "model.fp16.safetensors", W.clip_text_shapes(text2_cfg), 4,
                      dict(architectures=["CLIPTextModelWithProjection"], vocab_size=text2_cfg.vocab_size,
                           hidden_size=text2_cfg.hidden_size, intermediate_size=text2_cfg.intermediate_size,
                           num_hidden_layers=text2_cfg.num_layers, num_attention_heads=text2_cfg.num_heads,
                           max_position_embeddings=text2_cfg.max_positions, hidden_act=text2_cfg.act,
                           layer_norm_eps=text2_cfg.eps, projection_dim=text2_cfg.projection_dim)))
    parts += [
        # InstructPix2Pix fine-tune (diffusion/train_instruct_pix2pix_genima.py output): 8-channel conv_in
        (os.path.join(root, "pix2pix_ckpt", "checkpoint-200", "unet"), "diffusion_pytorch_model.safetensors",
         W.unet_shapes(dataclasses.replace(ucfg, in_channels=2 * vcfg.latent_channels)), 5,
         dict(ujson, in_channels=2 * vcfg.latent_channels)),
    ]
    for d, fname, shapes, salt, cfg in parts:
        os.makedirs(d, exist_ok=True)
        save_file(W.synth_state_dict(shapes, salt=salt), os.path.join(d, fname))
        with open(os.path.join(d, "config.json"), "w") as f:
            json.dump(cfg, f)
    os.makedirs(os.path.join(sd_ckpt, "scheduler"), exist_ok=True)
    with open(os.path.join(sd_ckpt, "scheduler", "scheduler_config.json"), "w") as f:
        json.dump(dict(_class_name="EulerAncestralDiscreteScheduler" if text2_cfg is not None else "EulerDiscreteScheduler",
                       num_train_timesteps=1000, beta_start=0.00085,
                       beta_end=0.012, beta_schedule="scaled_linear", timestep_spacing="trailing",
                       prediction_type="epsilon"), f)
    os.makedirs(ctl, exist_ok=True)
    act_sd = W.synth_state_dict(W.act_shapes(acfg), salt=3)
    torch.save({"cfg": {}, "_epoch": 0, "_num_iters": 0, "agent": {f"actor.{k}": v for k, v in act_sd.items()}},
               os.path.join(ctl, "latest.pt"))
    return dict(sd_ckpt=sd_ckpt, diffusion_ckpt=dif, controller_ckpt=ctl, pix2pix_ckpt=os.path.join(root, "pix2pix_ckpt"))
