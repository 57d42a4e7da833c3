"""Reference-facing plugin classes: the drop-in boundary of SURVEY.md §8(b).

  B200ControlNetAgent      hydra `_target_` replacement for `agent.SDControlNetAgent`
                           (controller/cfgs/eval_genima.yaml:27-28; controller/agent/sd_controlnet_agent.py:11-76 over
                           controller/agent/diffusion_agent.py:5-65): same constructor (eval_cfg), same
                           load_checkpoint / set_optimizations / common_setup / infer methods, same `.pipe` attribute and
                           `transform_to_half_resolution`; `.pipe` is a B200ControlNetPipeline.
  B200Pix2PixAgent         the same for `agent.SDPix2PixAgent` (controller/agent/sd_pix2pix_agent.py:11-60); `.pipe` is a
                           B200Pix2PixPipeline (InstructPix2Pix: VAE-encoded image latents, no ControlNet).
  B200SDXLControlNetAgent  the same for `agent.SDXLControlNetAgent` (controller/agent/sdxl_controlnet_agent.py:11-75).
  B200GenimaACTPolicy      replacement for `GenimaACTPolicy` (controller/method/genima_act.py:142-214): forward(qpos, image,
                           actions=None, is_pad=None, task_emb=None) -> a_hat [B, 20, 8]; inference only.
  B200GenimaACT            the `act` / `encode_clip_text` surface of `GenimaACT` (genima_act.py:273-346) on the obs-dict
                           conventions of the eval loop (controller/eval_genima.py:237-248).
Everything arithmetic goes to libgenima_b200.so; these classes only translate arguments.
"""
from __future__ import annotations

import dataclasses
import os
import re
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np
import torch

from . import checkpoint as ckpt
from . import weights as W
from .act_policy import DeviceACT
from .configs import ACTConfig, CLIPTextConfig, SchedulerConfig, TAESDConfig, UNetConfig, VAEConfig
from .ops import Ops
from .pipeline import B200ControlNetPipeline, B200Pix2PixPipeline, B200SDXLControlNetPipeline
from .text_encoder import DeviceCLIPText
from .unet import tensor_key

_OPS: Dict[int, Ops] = {}


def get_ops(device="cuda") -> Ops:
    """One Ops (gn_handle + workspace) per device, shared by the diffusion agent and the controller."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
    if idx not in _OPS:
        _OPS[idx] = Ops(idx)
    return _OPS[idx]


def _cfg_get(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    try:
        v = getattr(cfg, name)
    except Exception:
        return default
    return default if v is None and default is not None else v


def _preset(name: str):
    if name == "tiny":
        return UNetConfig.tiny(), VAEConfig.tiny(), CLIPTextConfig.tiny(), ACTConfig.tiny()
    if name in ("sd-turbo", "full", True):
        return UNetConfig(), VAEConfig(), CLIPTextConfig.sd_turbo(), ACTConfig()
    raise ValueError(f"unknown synthetic_weights preset {name!r} (use 'sd-turbo' or 'tiny')")


class _CenterResize:
    """transforms.Compose([Resize(r, BILINEAR), CenterCrop(r)]) on PIL images (controller/agent/diffusion_agent.py:44-62)
    without the torchvision dependency; an identity when the image is already r x r."""

    def __init__(self, resolution: int):
        self.r = int(resolution)

    def is_identity_for(self, size) -> bool:
        return tuple(size) == (self.r, self.r)

    def __call__(self, im):
        from PIL import Image

        w, h = im.size
        if (w, h) == (self.r, self.r):
            return im
        if w <= h:
            nw, nh = self.r, int(self.r * h / w)
        else:
            nw, nh = int(self.r * w / h), self.r
        im = im.resize((nw, nh), Image.BILINEAR)
        left, top = int(round((nw - self.r) / 2.0)), int(round((nh - self.r) / 2.0))
        return im.crop((left, top, left + self.r, top + self.r))


class B200ControlNetAgent:
    """Drop-in for agent.SDControlNetAgent.  Extra eval_cfg keys (all optional): `synthetic_weights` ('sd-turbo' | 'tiny';
    seeded synthetic weights when no checkpoint exists offline), `use_cuda_graph` (default True), `tokenizer` (callable
    list[str] -> ids [B, 77]; no CLIP BPE vocabulary is available offline)."""

    def __init__(self, eval_cfg, ops: Optional[Ops] = None):
        self.eval_cfg = eval_cfg
        self.pipe = None
        self._ops = ops
        self.load_checkpoint()
        self.set_optimizations()
        self.common_setup()

    # ---- controller/agent/sd_controlnet_agent.py:19-65
    def load_checkpoint(self):
        cfg = self.eval_cfg
        ops = self._ops or get_ops(_cfg_get(cfg, "device", "cuda"))
        autoenc = _cfg_get(cfg, "autoencoder", "") or ""
        taesd = "taesd" in autoenc          # same test as the reference (sd_controlnet_agent.py:45): AutoencoderTiny
        synth = _cfg_get(cfg, "synthetic_weights", None)
        tok = _cfg_get(cfg, "tokenizer", None)
        graph = bool(_cfg_get(cfg, "use_cuda_graph", True))
        if synth:
            ucfg, vcfg, tcfg, _ = _preset(synth)
            with_text = bool(_cfg_get(cfg, "synthetic_text_encoder", True))
            if taesd:
                vcfg = TAESDConfig.tiny() if synth == "tiny" else TAESDConfig()
            vae_sd = W.synth_state_dict(W.taesd_decoder_shapes(vcfg) if taesd else W.vae_decoder_shapes(vcfg), salt=2)
            self.pipe = B200ControlNetPipeline(
                ops, W.synth_state_dict(W.unet_shapes(ucfg)), W.synth_state_dict(W.controlnet_shapes(ucfg), salt=1),
                vae_sd,
                W.synth_state_dict(W.clip_text_shapes(tcfg)) if with_text else None,
                ucfg, vcfg, tcfg, SchedulerConfig(), tokenizer=tok, use_cuda_graph=graph)
            return
        loaded = ckpt.load_sd_turbo(_cfg_get(cfg, "sd_ckpt"), _cfg_get(cfg, "diffusion_ckpt"))
        if taesd:
            loaded["vae"], loaded["vae_cfg"] = ckpt.load_taesd(autoenc)
        # string prompts: the snapshot's own tokenizer files (sd_ckpt/tokenizer), unless the caller supplied a callable
        self.pipe = B200ControlNetPipeline(ops, loaded["unet"], loaded["controlnet"], loaded["vae"], loaded["text"],
                                           loaded["unet_cfg"], loaded["vae_cfg"], loaded["text_cfg"],
                                           loaded["scheduler_cfg"], tokenizer=tok or loaded["tokenizer"],
                                           use_cuda_graph=graph)

    # ---- controller/agent/diffusion_agent.py:21-42 (the toggles are accepted; the kernels are always fused)
    def set_optimizations(self):
        cfg = self.eval_cfg
        if _cfg_get(cfg, "vae_slicing", False):
            self.pipe.vae.enable_slicing()
        if _cfg_get(cfg, "upcast_vae", False):
            self.pipe.upcast_vae()
        if _cfg_get(cfg, "fused_projections", False):
            self.pipe.fuse_qkv_projections(vae=False)
        if _cfg_get(cfg, "enable_xformers_memory_efficient_attention", False):
            self.pipe.enable_xformers_memory_efficient_attention()
        self.pipe.set_progress_bar_config(disable=(not _cfg_get(cfg, "show_diffusion_progress", False)))
        self.pipe.to(_cfg_get(cfg, "device", "cuda"))

    # ---- controller/agent/diffusion_agent.py:44-62
    def common_setup(self):
        resolution = int(_cfg_get(self.eval_cfg, "image_resolution", 512))
        self.transform_to_resolution = _CenterResize(resolution)
        self.transform_to_half_resolution = _CenterResize(resolution // 2)

    # ---- controller/agent/sd_controlnet_agent.py:67-76
    def infer(self, *args, **kwargs):
        return self.pipe(
            prompt=kwargs["prompts"],
            image=kwargs["images"],
            negative_prompt=kwargs["negative_prompts"],
            num_inference_steps=kwargs["num_inference_steps"],
            guidance_scale=kwargs["guidance_scale"],
            generator=kwargs["generator"],
            **{k: kwargs[k] for k in ("latents", "prompt_embeds", "output_type") if k in kwargs},
        )


class B200Pix2PixAgent(B200ControlNetAgent):
    """Drop-in for agent.SDPix2PixAgent (controller/agent/sd_pix2pix_agent.py:11-60): same constructor, optimisation
    toggles, transforms and `infer` keywords as the ControlNet agent; `.pipe` is a B200Pix2PixPipeline (fine-tuned
    8-channel U-Net from `<diffusion_ckpt>/checkpoint-*/unet`, VAE encoder + decoder, text encoder, scheduler from
    sd_ckpt)."""

    def load_checkpoint(self):
        cfg = self.eval_cfg
        ops = self._ops or get_ops(_cfg_get(cfg, "device", "cuda"))
        if "taesd" in (_cfg_get(cfg, "autoencoder", "") or ""):
            raise NotImplementedError("the InstructPix2Pix agent needs the AutoencoderKL encoder; TAESD is decode-only")
        synth = _cfg_get(cfg, "synthetic_weights", None)
        tok = _cfg_get(cfg, "tokenizer", None)
        graph = bool(_cfg_get(cfg, "use_cuda_graph", True))
        if synth:
            ucfg, vcfg, tcfg, _ = _preset(synth)
            ucfg = dataclasses.replace(ucfg, in_channels=2 * vcfg.latent_channels)
            with_text = bool(_cfg_get(cfg, "synthetic_text_encoder", True))
            vae_sd = W.synth_state_dict(W.vae_decoder_shapes(vcfg), salt=2)
            vae_sd.update(W.synth_state_dict(W.vae_encoder_shapes(vcfg), salt=2))
            self.pipe = B200Pix2PixPipeline(
                ops, W.synth_state_dict(W.unet_shapes(ucfg), salt=5), vae_sd,
                W.synth_state_dict(W.clip_text_shapes(tcfg)) if with_text else None,
                ucfg, vcfg, tcfg, SchedulerConfig(), tokenizer=tok, use_cuda_graph=graph)
            return
        loaded = ckpt.load_sd_pix2pix(_cfg_get(cfg, "sd_ckpt"), _cfg_get(cfg, "diffusion_ckpt"))
        self.pipe = B200Pix2PixPipeline(ops, loaded["unet"], loaded["vae"], loaded["text"], loaded["unet_cfg"],
                                        loaded["vae_cfg"], loaded["text_cfg"], loaded["scheduler_cfg"],
                                        tokenizer=tok or loaded["tokenizer"], use_cuda_graph=graph)


class B200SDXLControlNetAgent(B200ControlNetAgent):
    """Drop-in for agent.SDXLControlNetAgent (controller/agent/sdxl_controlnet_agent.py:11-75): ControlNet from
    `<diffusion_ckpt>/checkpoint-*/controlnet`, everything else from the SDXL snapshot in sd_ckpt (two text encoders,
    Euler-ancestral scheduler for sdxl-turbo); `autoencoder` containing "taesdxl" selects AutoencoderTiny (:45-49).
    `synthetic_weights`: 'sdxl' (full size) | 'sdxl-tiny'."""

    def load_checkpoint(self):
        cfg = self.eval_cfg
        ops = self._ops or get_ops(_cfg_get(cfg, "device", "cuda"))
        autoenc = _cfg_get(cfg, "autoencoder", "") or ""
        taesd = "taesdxl" in autoenc
        synth = _cfg_get(cfg, "synthetic_weights", None)
        tok, tok2 = _cfg_get(cfg, "tokenizer", None), _cfg_get(cfg, "tokenizer_2", None)
        graph = bool(_cfg_get(cfg, "use_cuda_graph", True))
        if synth:
            if synth == "sdxl-tiny":
                ucfg, vcfg = UNetConfig.sdxl_tiny(), dataclasses.replace(VAEConfig.tiny(), scaling_factor=0.13025)
                t1, t2 = CLIPTextConfig.tiny(), CLIPTextConfig.tiny(projection_dim=64)
            elif synth == "sdxl":
                ucfg, vcfg = UNetConfig.sdxl(), VAEConfig(scaling_factor=0.13025)
                t1, t2 = CLIPTextConfig.sdxl_clip_l(), CLIPTextConfig.sdxl_open_clip_bigg()
            else:
                raise ValueError(f"unknown synthetic_weights preset {synth!r} for the SDXL agent (use 'sdxl' or 'sdxl-tiny')")
            with_text = bool(_cfg_get(cfg, "synthetic_text_encoder", True))
            if taesd:
                vcfg = TAESDConfig.tiny() if synth == "sdxl-tiny" else TAESDConfig()
            vae_sd = W.synth_state_dict(W.taesd_decoder_shapes(vcfg) if taesd else W.vae_decoder_shapes(vcfg), salt=2)
            self.pipe = B200SDXLControlNetPipeline(
                ops, W.synth_state_dict(W.unet_shapes(ucfg)), W.synth_state_dict(W.controlnet_shapes(ucfg), salt=1),
                vae_sd, W.synth_state_dict(W.clip_text_shapes(t1)) if with_text else None,
                W.synth_state_dict(W.clip_text_shapes(t2), salt=4) if with_text else None,
                ucfg, vcfg, t1, t2, tokenizer=tok, tokenizer_2=tok2, use_cuda_graph=graph)
            return
        loaded = ckpt.load_sdxl(_cfg_get(cfg, "sd_ckpt"), _cfg_get(cfg, "diffusion_ckpt"))
        if taesd:
            loaded["vae"], loaded["vae_cfg"] = ckpt.load_taesd(autoenc)
        self.pipe = B200SDXLControlNetPipeline(
            ops, loaded["unet"], loaded["controlnet"], loaded["vae"], loaded["text"], loaded["text2"],
            loaded["unet_cfg"], loaded["vae_cfg"], loaded["text_cfg"], loaded["text2_cfg"], loaded["scheduler_cfg"],
            tokenizer=tok or loaded["tokenizer"], tokenizer_2=tok2 or loaded["tokenizer_2"], use_cuda_graph=graph)

    def infer(self, *args, **kwargs):
        extra = ("latents", "prompt_embeds", "pooled_prompt_embeds", "output_type")
        return self.pipe(
            prompt=kwargs["prompts"],
            image=kwargs["images"],
            negative_prompt=kwargs["negative_prompts"],
            num_inference_steps=kwargs["num_inference_steps"],
            guidance_scale=kwargs["guidance_scale"],
            generator=kwargs["generator"],
            **{k: kwargs[k] for k in extra if k in kwargs},
        )


class B200GenimaACTPolicy:
    """GenimaACTPolicy.forward (controller/method/genima_act.py:165-214), inference branch."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: ACTConfig = ACTConfig(), ops: Optional[Ops] = None,
                 device="cuda", use_cuda_graph: bool = True):
        self.use_cuda_graph = use_cuda_graph
        self.cfg = cfg
        self.ops = ops or get_ops(device)
        self._sd = dict(state_dict)
        self.impl = DeviceACT(self.ops, self._sd, cfg)
        self.training = False

    def state_dict(self):
        return dict(self._sd)

    def load_state_dict(self, sd, strict: bool = True):
        ckpt.check_schema(sd, W.act_shapes(self.cfg), "ACT policy")
        self._sd = dict(sd)
        self.impl = DeviceACT(self.ops, self._sd, self.cfg)

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def forward(self, qpos: torch.Tensor, image: torch.Tensor, actions: torch.Tensor = None,
                is_pad: torch.Tensor = None, task_emb: torch.Tensor = None) -> torch.Tensor:
        if actions is not None:
            raise NotImplementedError("training (actions is not None) is out of scope: this is the inference path")
        if task_emb is None:
            raise ValueError("task_emb is required (genima_act.yaml: use_lang_cond=True)")
        if getattr(self, "use_cuda_graph", True):
            a_hat, _ = self.impl.forward_graphed(qpos, image, task_emb)
        else:
            a_hat, _ = self.impl.forward(qpos, image, task_emb)
        return a_hat

    __call__ = forward


def _hp(spec, name, default):
    """Hyper-parameter `name` of a hydra `_partial_: true` node (functools.partial -> `.keywords`), a DictConfig / dict,
    or None."""
    if spec is None:
        return default
    kw = getattr(spec, "keywords", None)
    if isinstance(kw, dict):
        return kw.get(name, default)
    return _cfg_get(spec, name, default)


def _space_get(space, key):
    """gymnasium.spaces.Dict item (or a plain dict of objects with `.shape`)."""
    try:
        return space[key]
    except Exception:
        return getattr(space, "spaces", {}).get(key)


def _space_keys(space):
    if hasattr(space, "spaces"):
        return list(space.spaces.keys())
    return list(space.keys())


class _IncompatibleKeys(tuple):
    """torch.nn.modules.module._IncompatibleKeys look-alike returned by load_state_dict."""

    def __new__(cls, missing_keys, unexpected_keys):
        self = super().__new__(cls, (missing_keys, unexpected_keys))
        self.missing_keys, self.unexpected_keys = missing_keys, unexpected_keys
        return self


class B200GenimaACT:
    """Drop-in for `method.genima_act.GenimaACT` on the eval path (controller/method/genima_act.py:217-346 over RoboBase
    `ActBCAgent`): hydra target of the checkpoint's `config.yaml` (`controller/cfgs/method/genima_act.yaml:4`),

        hydra.utils.instantiate(train_cfg.method, device=, observation_space=, action_space=, num_train_envs=,
                                replay_alpha=, replay_beta=, frame_stack_on_channel=)       # eval_genima.py:55-64

    followed by `.train(False)` (:66), `.state_dict()` / `.load_state_dict(checkpoint["agent"], strict=False)`
    (:91-103), `robobase_utils.eval_mode(agent)` (:200, reads `.training`, calls `.train(bool)`) and
    `.act(obs, step=, eval_mode=True)` (:243-247).  The network hyper-parameters come from the same places the
    reference reads them: `actor_model` / `encoder_model` (the yaml's `_partial_` nodes), the observation space
    (state size, number of rgb views, image size: genima_act.py:227-231) and the action space.

    State-dict keys are RoboBase's: the policy lives under `actor.` (`actor.encoder_model.*`, `actor.actor_model.*`,
    train_act.py:262-279 saves them minus `clip_model.*`).  No weights exist until `load_state_dict` (the reference
    starts from a random init it immediately overwrites); `act` before that raises.

    The ViT-B/32 text tower the reference fetches with `clip.load` at the first `encode_clip_text` (genima_act.py:315-321)
    comes from `clip_state_dict` (transformers / OpenAI naming) or `clip_ckpt` (a local `ViT-B-32.pt`, TorchScript
    archive or state dict): hub downloads are impossible offline.

    Training-side keywords (lr, weight_decay, replay_*, ...) are accepted and kept in `.hparams`; `update()` raises."""

    def __init__(self, *args, device="cuda", observation_space=None, action_space=None, actor_model=None,
                 encoder_model=None, policy: Optional[B200GenimaACTPolicy] = None,
                 clip_state_dict: Optional[Dict[str, torch.Tensor]] = None, clip_ckpt: Optional[str] = None,
                 clip_cfg: CLIPTextConfig = CLIPTextConfig.vit_b32(), act_cfg: Optional[ACTConfig] = None,
                 ops: Optional[Ops] = None, use_cuda_graph: bool = True, **kwargs):
        if args and isinstance(args[0], B200GenimaACTPolicy):     # round-1 signature: B200GenimaACT(policy, clip_sd, cfg)
            policy, args = args[0], args[1:]
            if args:
                clip_state_dict, args = args[0], args[1:]
            if args:
                clip_cfg, args = args[0], args[1:]
        if args:
            raise TypeError(f"unexpected positional arguments: {args!r}")
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.observation_space, self.action_space = observation_space, action_space
        self.hparams = dict(kwargs)
        self.training = False
        self.use_cuda_graph = bool(use_cuda_graph)
        self._ops = ops if ops is not None else (policy.ops if policy is not None else None)
        self._actor: Optional[B200GenimaACTPolicy] = policy
        self._actor_sd: Optional[Dict[str, torch.Tensor]] = None        # loaded, not yet bound to the device
        self.cfg = policy.cfg if policy is not None else (act_cfg or self._config_from_spaces(actor_model, encoder_model))
        self._clip_sd, self._clip_ckpt, self.clip_cfg = clip_state_dict, clip_ckpt, clip_cfg
        self.clip: Optional[DeviceCLIPText] = None
        self._emb_cache: "OrderedDict[bytes, tuple]" = OrderedDict()     # token content -> (task_emb, last hidden)
        self._ident_cache: "OrderedDict[tuple, tuple]" = OrderedDict()   # same tensor object -> the same, no D2H read

    # ---- genima_act.py:221-249 (what build_actor derives from the spaces and the yaml)
    def _config_from_spaces(self, actor_model, encoder_model) -> ACTConfig:
        base = ACTConfig()
        kw = dict(hidden_dim=int(_hp(actor_model, "hidden_dim", base.hidden_dim)),
                  enc_layers=int(_hp(actor_model, "enc_layers", base.enc_layers)),
                  dec_layers=int(_hp(actor_model, "dec_layers", base.dec_layers)),
                  dim_feedforward=int(_hp(actor_model, "dim_feedforward", base.dim_feedforward)),
                  nheads=int(_hp(actor_model, "nheads", base.nheads)),
                  num_queries=int(_hp(actor_model, "num_queries", base.num_queries)),
                  state_dim=int(_hp(actor_model, "state_dim", base.state_dim)),
                  action_dim=int(_hp(actor_model, "action_dim", base.action_dim)))
        if _hp(actor_model, "pre_norm", False):
            raise NotImplementedError("pre_norm transformer layers are not implemented (genima_act.yaml: pre_norm false)")
        bb = _hp(encoder_model, "backbone", "resnet18")
        if bb != "resnet18":
            raise NotImplementedError(f"backbone {bb!r} is not implemented (genima_act.yaml: resnet18)")
        if _hp(encoder_model, "position_embedding", "sine") != "sine":
            raise NotImplementedError("only the sine position embedding is implemented")
        if encoder_model is not None and not _hp(encoder_model, "use_lang_cond", True):
            raise NotImplementedError("use_lang_cond=False (no FiLM) is not implemented")
        enc_hidden = _hp(encoder_model, "hidden_dim", kw["hidden_dim"])
        if int(enc_hidden) != kw["hidden_dim"]:
            raise ValueError("encoder_model.hidden_dim and actor_model.hidden_dim differ")
        osp = self.observation_space
        if osp is not None:
            low = _space_get(osp, "low_dim_state")
            if low is not None:
                kw["state_dim"] = int(np.prod(low.shape))               # genima_act.py:228
            rgb = [k for k in _space_keys(osp) if re.match(r"rgb.*", k) or "_rgb" in k]
            if rgb:
                shp = tuple(_space_get(osp, rgb[0]).shape)              # (T, 3, H, W) with the frame-stack wrapper
                frames = int(shp[0]) if len(shp) == 4 else 1
                kw["num_views"] = len(rgb) * frames
                kw["image_size"] = int(shp[-1])
        if self.action_space is not None:
            kw["action_dim"] = int(self.action_space.shape[-1])         # genima_act.py:229
        return dataclasses.replace(base, **kw)

    # ---- torch.nn.Module surface the eval workspace touches
    @property
    def ops(self) -> Ops:
        if self._ops is None:
            self._ops = get_ops(self.device if self.device.type == "cuda" else "cuda")
        return self._ops

    @property
    def actor(self) -> Optional[B200GenimaACTPolicy]:
        """The policy (`self.actor` of the reference, genima_act.py:244-249), bound to the device on first use."""
        if self._actor_sd is not None:
            sd, self._actor_sd = self._actor_sd, None
            if self._actor is None:
                self._actor = B200GenimaACTPolicy(sd, self.cfg, ops=self.ops, use_cuda_graph=self.use_cuda_graph)
            else:
                self._actor.load_state_dict(sd)
        return self._actor

    def train(self, mode: bool = True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    def to(self, *a, **k):
        return self

    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        """`actor.*` keys of the RoboBase snapshot schema; shape-only (meta) tensors until weights are loaded."""
        if self._actor_sd is not None:
            return OrderedDict((f"actor.{k}", v) for k, v in self._actor_sd.items())
        if self._actor is not None:
            return OrderedDict((f"actor.{k}", v) for k, v in self._actor.state_dict().items())
        return OrderedDict((f"actor.{k}", torch.empty(shape, device="meta"))
                           for k, shape in W.act_shapes(self.cfg).items())

    def load_state_dict(self, state_dict, strict: bool = True):
        """`checkpoint["agent"]` (eval_genima.py:102).  Binds the device policy to the `actor.*` tensors; other keys
        (`clip_model.*`, optimiser-side duplicates RoboBase may register) are reported as unexpected, and with
        strict=True raise like torch does.  A missing policy tensor always raises: nothing can run without it."""
        want = W.act_shapes(self.cfg)
        sd = {k[len("actor."):]: v for k, v in state_dict.items() if k.startswith("actor.")}
        missing = [f"actor.{k}" for k in want if k not in sd]
        unexpected = [k for k in state_dict if not (k.startswith("actor.") and k[len("actor."):] in want)]
        if missing:
            raise RuntimeError(f"Error(s) in loading state_dict for B200GenimaACT: missing keys {missing[:8]}"
                               f"{' ...' if len(missing) > 8 else ''}")
        if strict and unexpected:
            raise RuntimeError(f"Error(s) in loading state_dict for B200GenimaACT: unexpected keys {unexpected[:8]}")
        sd = {k: sd[k] for k in want}
        ckpt.check_schema(sd, want, "ACT policy")
        self._actor_sd = sd          # bound to the device (DeviceACT: packing, BN / FiLM folds) at the first act()
        if any(k.startswith("clip_model.") for k in state_dict) and self._clip_sd is None and self._clip_ckpt is None:
            self._clip_sd = {k[len("clip_model."):]: v for k, v in state_dict.items() if k.startswith("clip_model.")}
            self.clip = None
        return _IncompatibleKeys([], unexpected)

    def update(self, *a, **k):
        raise NotImplementedError("training (GenimaACT.update, genima_act.py:348-422) is out of scope: inference path only")

    # ---- genima_act.py:314-346
    def _ensure_clip(self) -> DeviceCLIPText:
        if self.clip is None:
            sd = self._clip_sd
            if sd is None and self._clip_ckpt is not None:
                sd = ckpt.load_openai_clip_text(self._clip_ckpt, self.clip_cfg)
            if sd is None:
                raise RuntimeError(
                    "no CLIP text tower bound: the reference downloads ViT-B/32 with clip.load (genima_act.py:315-321), "
                    "which is impossible offline — pass clip_ckpt=<local ViT-B-32.pt> or clip_state_dict=, or supply "
                    "task_emb to the policy yourself")
            if "token_embedding.weight" in sd:                       # OpenAI naming -> transformers naming
                sd = ckpt.openai_clip_text_to_hf(sd)
            self.clip = DeviceCLIPText(self.ops, sd, self.clip_cfg)
        return self.clip

    @staticmethod
    def _lru_put(cache: "OrderedDict", key, value, limit: int) -> None:
        cache[key] = value
        cache.move_to_end(key)
        while len(cache) > limit:
            cache.popitem(last=False)

    def encode_clip_text(self, tokens: torch.Tensor):
        """tokens [B, T, 77] int -> (task_emb [B, proj] fp32, last hidden [B*T, 77, d]).  The text is constant for an
        episode (controller/env/rlbench_utils.py:156), so results are cached by token content; both caches are small
        LRUs (the eval loop hands over a fresh device tensor every step, eval_genima.py:237-240)."""
        shape = tokens.shape
        ident = tensor_key(tokens)
        hit = self._ident_cache.get(ident)
        if hit is not None and hit[2] is tokens:      # the very same tensor object, unmodified: no device-to-host read
            self._ident_cache.move_to_end(ident)
            return hit[0], hit[1]
        tks = tokens.reshape(-1, shape[-1])
        key = tks.cpu().numpy().tobytes() + repr(tuple(shape)).encode()
        hit = self._emb_cache.get(key)
        if hit is None:
            clip = self._ensure_clip()
            emb, pooled = clip(tks.to(self.ops.device, torch.int64))
            x = pooled.reshape(shape[0], shape[1], -1)[:, 0].contiguous()   # text does not change across frames
            hit = (x, emb)
        self._lru_put(self._emb_cache, key, hit, 16)
        self._lru_put(self._ident_cache, ident, (hit[0], hit[1], tokens), 4)
        return hit

    @torch.no_grad()
    def act(self, obs: Dict[str, torch.Tensor], step: int = 0, eval_mode: bool = True) -> torch.Tensor:
        """obs: low_dim_state [B, T, S]; `*rgb*` [B, T, 3, H, W] (dict order = view order); lang_tokens [B, T, 77]."""
        actor = self.actor
        if actor is None:
            raise RuntimeError("B200GenimaACT.act before load_state_dict: no controller weights are bound")
        low = obs["low_dim_state"]
        qpos = low.reshape(low.shape[0], -1).float()
        rgbs = [v for k, v in obs.items() if re.match(r"rgb.*", k) or "_rgb" in k]
        rgb = torch.stack(rgbs, 1)                                    # [B, V, T, 3, H, W]
        image = rgb.reshape(rgb.shape[0], -1, 3, rgb.shape[-2], rgb.shape[-1])
        if image.dtype != torch.uint8:
            image = image.float()                                     # genima_act.py:296 (`rgb.float()`)
        task_emb, _ = self.encode_clip_text(obs["lang_tokens"])
        return actor(qpos, image, task_emb=task_emb)
