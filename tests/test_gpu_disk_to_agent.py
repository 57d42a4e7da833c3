"""Disk -> plugin -> GPU: both plugin points built the way `GenimaEvalWorkspace.__init__` builds them
(controller/eval_genima.py:55-66, 86-89) from checkpoints in the reference's on-disk layouts
(controller/agent/sd_controlnet_agent.py:19-42: `<diffusion_ckpt>/checkpoint-*/controlnet` + the sd-turbo snapshot;
controller/train_act.py:262-279: RoboBase `latest.pt`; `clip.load`'s ViT-B-32.pt), then one iteration of the loop body
(eval_genima.py:163-249) with string prompts, PIL images and a CUDA generator, checked against the CPU oracle fed from
the SAME files through its own readers."""
import dataclasses
import json
import os

import numpy as np
import pytest
import torch

import fake_workspace as fw
from genima_b200 import checkpoint as ckpt
from genima_b200.configs import ACTConfig, CLIPTextConfig, UNetConfig, VAEConfig

pytestmark = pytest.mark.gpu

N_STEPS = 3


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    from test_controller_plugin import _openai_clip_text_sd

    root = str(tmp_path_factory.mktemp("disk"))
    # tiny networks at the reference's real image sizes (4 x 256^2 views, one 512^2 tile): the reference-signature glue
    # (tile_images / untile_images) asserts those sizes like controller/utils/misc.py does
    ucfg, vcfg, tcfg = UNetConfig.tiny(), VAEConfig.tiny(), CLIPTextConfig.tiny()
    acfg = dataclasses.replace(ACTConfig.tiny(), image_size=256)
    dirs = ckpt.save_synthetic_checkpoints(root, ucfg, vcfg, tcfg, acfg)
    fw.write_tiny_clip_tokenizer(os.path.join(dirs["sd_ckpt"], "tokenizer"))
    clip_cfg = CLIPTextConfig(vocab_size=600, hidden_size=64, intermediate_size=128, num_layers=2, num_heads=2,
                              act="quick_gelu", projection_dim=acfg.task_emb_dim)
    clip_path = os.path.join(root, "ViT-B-32.pt")
    torch.save({k: v.half() for k, v in _openai_clip_text_sd(clip_cfg, seed=5).items()}, clip_path)
    return dict(dirs=dirs, cfgs=(ucfg, vcfg, tcfg, acfg), clip_cfg=clip_cfg, clip_path=clip_path)


def _eval_cfg(dirs, S):
    """controller/cfgs/eval_genima.yaml:27-45 with local paths."""
    return dict(diffusion_agent={"_target_": "genima_b200.agents.B200ControlNetAgent"},
                diffusion_ckpt=dirs["diffusion_ckpt"], sd_ckpt=dirs["sd_ckpt"], autoencoder="", device="cuda",
                execution_horizon=20, num_diffusion_steps=N_STEPS, guidance_scale=0.0, diffusion_seed=2,
                image_resolution=2 * S, torch_compile=False, channels_last=False, tf32=False, vae_slicing=False,
                upcast_vae=False, fused_projections=True, enable_xformers_memory_efficient_attention=False,
                show_diffusion_progress=False)


def test_disk_to_agents_to_actions(world):
    from PIL import Image

    from oracle import act as act_oracle
    from oracle import clip_text as clip_oracle
    from oracle import tiling
    from oracle.pipeline import controlnet_pipeline

    ucfg, vcfg, tcfg, acfg = world["cfgs"]
    dirs = world["dirs"]
    S = acfg.image_size
    eval_cfg = _eval_cfg(dirs, S)
    # ---- GenimaEvalWorkspace.__init__ (eval_genima.py:55-66, 86-89)
    fw.install_reference_stub_modules()
    node = dict(fw.GENIMA_ACT_YAML, _target_="genima_b200.agents.B200GenimaACT")
    for part, keys in (("actor_model", ("hidden_dim", "enc_layers", "dec_layers", "dim_feedforward", "nheads",
                                        "num_queries")), ("encoder_model", ("hidden_dim",))):
        node[part] = dict(node[part], **{k: getattr(acfg, k) for k in keys})
    obs_space, act_space = fw.make_spaces(fw.CAMERAS, 1, S, acfg.state_dim, acfg.action_dim)
    controller = fw.instantiate(node, device=torch.device("cuda"), observation_space=obs_space, action_space=act_space,
                                num_train_envs=1, replay_alpha=0.6, replay_beta=0.4, frame_stack_on_channel=False,
                                act_cfg=acfg, clip_ckpt=world["clip_path"], clip_cfg=world["clip_cfg"])
    controller.train(False)
    diffusion_agent = fw.instantiate(eval_cfg["diffusion_agent"], eval_cfg)
    fw.load_controller_ckpt(controller, os.path.join(dirs["controller_ckpt"], "latest.pt"), device="cuda")

    # ---- one iteration of the loop body (eval_genima.py:129-135, 163-249)
    g = torch.Generator().manual_seed(11)
    views = torch.randint(0, 256, (4, 3, S, S), generator=g, dtype=torch.uint8).numpy()     # obs[f"{camera}_rgb"][t]
    qpos = torch.randn(1, acfg.state_dim, generator=g).numpy().astype(np.float32)
    lang = torch.randint(1, 590, (1, 77), generator=g).numpy().astype(np.int32)
    lang[0, 0], lang[0, 10], lang[0, 11:] = 598, 599, 0
    generator = [torch.Generator(device="cuda").manual_seed(eval_cfg["diffusion_seed"])]
    goal = "open the box"
    prompts = [f"tiled perspectives of a robot arm executing '{goal}'"]
    negative_prompts = ["monochrome, lowres, bad anatomy, worst quality, low quality"]
    rgbs = [Image.fromarray(np.transpose(v, (1, 2, 0))) for v in views]
    from genima_b200.host_glue import tile_images, untile_images

    tiled = tile_images(rgbs, 1)
    with torch.inference_mode(), fw.eval_mode(controller):
        target = diffusion_agent.infer(images=tiled, prompts=prompts, negative_prompts=negative_prompts,
                                       num_inference_steps=eval_cfg["num_diffusion_steps"],
                                       guidance_scale=eval_cfg["guidance_scale"], generator=generator * len(tiled))
        untiled = untile_images(target[0], fw.CAMERAS, diffusion_agent.transform_to_half_resolution)
        obs = {f"{c}_rgb": untiled[c] for c in fw.CAMERAS}
        obs["low_dim_state"] = qpos
        obs["lang_tokens"] = lang
        obs = {k: torch.from_numpy(v).to("cuda").unsqueeze(0) for k, v in obs.items()}
        actions = controller.act(obs, step=0, eval_mode=True)[0]
        actions = actions.detach().cpu().numpy()
        actions2 = controller.act({k: v.clone() for k, v in obs.items()}, step=20, eval_mode=True)[0].cpu().numpy()
    assert actions.shape == (acfg.num_queries, acfg.action_dim)
    assert np.array_equal(actions, actions2)            # fresh tensors, same content: cached text, same graph, same bits
    assert len(controller._ident_cache) <= 4 and len(controller._emb_cache) == 1

    # ---- the oracle, reading the same files with its own (safetensors / torch.load) calls
    from safetensors.torch import load_file

    def rd(*p):
        d = os.path.join(*p)
        f = [n for n in sorted(os.listdir(d)) if n.endswith(".safetensors")][0]
        return {k: v.float() for k, v in load_file(os.path.join(d, f)).items()}

    unet_sd, vae_sd = rd(dirs["sd_ckpt"], "unet"), rd(dirs["sd_ckpt"], "vae")
    text_sd = rd(dirs["sd_ckpt"], "text_encoder")
    cn_sd = rd(dirs["diffusion_ckpt"], "checkpoint-1000", "controlnet")       # natsort: the LAST checkpoint-*
    act_sd = {k[len("actor."):]: v.float() for k, v in
              torch.load(os.path.join(dirs["controller_ckpt"], "latest.pt"), weights_only=False)["agent"].items()}
    with open(os.path.join(dirs["sd_ckpt"], "tokenizer", "vocab.json")) as f:
        vocab = json.load(f)
    ids = diffusion_agent.pipe.tokenizer(prompts)       # transformers.CLIPTokenizer on the snapshot's files
    assert ids.shape == (1, 77) and int(ids[0, 0]) == vocab["<|startoftext|>"]
    ctx, _ = clip_oracle.clip_text_forward(text_sd, tcfg, ids)
    lat = torch.randn((1, 4, S // 4, S // 4), generator=torch.Generator(device="cuda").manual_seed(2), device="cuda",
                      dtype=torch.float16).float().cpu()           # the draw diffusers' prepare_latents makes
    tile = tiling.tile_views(np.transpose(views, (0, 2, 3, 1)))[None]
    assert np.array_equal(tile[0], np.asarray(tiled[0]))
    ref = controlnet_pipeline(unet_sd, cn_sd, vae_sd, ucfg, vcfg, tile, ctx, lat, N_STEPS)
    got_tile = np.asarray(target[0][0])
    d = np.abs(got_tile.astype(np.int32) - ref["u8"][0].astype(np.int32))
    osd = torch.load(world["clip_path"], weights_only=False)
    from test_controller_plugin import _encode_clip_text_openai

    task_ref, _ = _encode_clip_text_openai({k: v.float() for k, v in osd.items()}, world["clip_cfg"],
                                           torch.from_numpy(lang)[None].long())
    gen = tiling.untile_views(ref["u8"][0])
    image = torch.from_numpy(gen).permute(0, 3, 1, 2)[None].float()
    a_ref, _ = act_oracle.act_forward(act_sd, acfg, torch.from_numpy(qpos), image, task_ref)
    err = float(np.abs(actions - a_ref[0].numpy()).max() / a_ref.abs().max())
    print(f"disk -> agents -> actions: tile max |diff| {d.max()} levels ({100 * (d <= 1).mean():.2f}% within 1); "
          f"a_hat normalised max err {err:.3e}")
    assert d.max() <= 3 and (d <= 1).mean() > 0.995
    assert err < 4e-3


def test_controller_reload_rebinds_the_device_policy(world):
    """load_controller_ckpt runs once per evaluated checkpoint (eval_genima.py:118-122): a second snapshot must replace the
    first one's weights (new DeviceACT, no stale CUDA graph)."""
    from genima_b200 import weights as W

    ucfg, vcfg, tcfg, acfg = world["cfgs"]
    S = acfg.image_size
    fw.install_reference_stub_modules()
    obs_space, act_space = fw.make_spaces(fw.CAMERAS, 1, S, acfg.state_dim, acfg.action_dim)
    node = dict(fw.GENIMA_ACT_YAML, _target_="genima_b200.agents.B200GenimaACT")
    ctl = fw.instantiate(node, device="cuda", observation_space=obs_space, action_space=act_space, act_cfg=acfg,
                         clip_ckpt=world["clip_path"], clip_cfg=world["clip_cfg"])
    g = torch.Generator().manual_seed(3)
    obs = {f"{c}_rgb": torch.randint(0, 256, (1, 1, 3, S, S), generator=g, dtype=torch.uint8).cuda() for c in fw.CAMERAS}
    obs["low_dim_state"] = torch.randn(1, 1, acfg.state_dim, generator=g).cuda()
    lang = torch.zeros(1, 1, 77, dtype=torch.int32)
    lang[0, 0, 0], lang[0, 0, 5] = 598, 599
    obs["lang_tokens"] = lang.cuda()
    outs = []
    for salt in (3, 9, 3):
        sd = W.synth_state_dict(W.act_shapes(acfg), salt=salt)
        ctl.load_state_dict({f"actor.{k}": v for k, v in sd.items()}, strict=False)
        outs.append(ctl.act(obs, step=0, eval_mode=True).float().cpu().clone())
    assert torch.equal(outs[0], outs[2]) and not torch.equal(outs[0], outs[1])
