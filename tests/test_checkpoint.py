"""Weight-format loaders on round-tripped synthetic checkpoints in the reference's on-disk layouts."""
import os

import pytest
import torch

from fake_workspace import write_tiny_clip_tokenizer

from genima_b200 import checkpoint as ckpt
from genima_b200 import weights as W
from genima_b200.configs import ACTConfig, CLIPTextConfig, UNetConfig, VAEConfig


@pytest.fixture(scope="module")
def dirs(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("ckpt"))
    return ckpt.save_synthetic_checkpoints(root, UNetConfig.tiny(), VAEConfig.tiny(), CLIPTextConfig.tiny(),
                                           ACTConfig.tiny())


def test_last_checkpoint_is_chosen_in_natural_order(dirs, tmp_path):
    assert ckpt.find_controlnet_dir(dirs["diffusion_ckpt"]).endswith(os.path.join("checkpoint-1000", "controlnet"))
    for n in ("checkpoint-9", "checkpoint-10", "checkpoint-100"):
        os.makedirs(tmp_path / n / "controlnet")
    assert ckpt.find_controlnet_dir(str(tmp_path)).endswith(os.path.join("checkpoint-100", "controlnet"))
    empty = tmp_path / "flat"
    os.makedirs(empty)
    assert ckpt.find_controlnet_dir(str(empty)) == str(empty)      # no checkpoint-* sub-directory: the dir itself


def test_load_sd_turbo_round_trip(dirs):
    out = ckpt.load_sd_turbo(dirs["sd_ckpt"], dirs["diffusion_ckpt"])
    assert out["unet_cfg"] == UNetConfig.tiny()
    assert out["vae_cfg"].block_out_channels == VAEConfig.tiny().block_out_channels
    assert out["scheduler_cfg"].timestep_spacing == "trailing"
    want = W.synth_state_dict(W.controlnet_shapes(UNetConfig.tiny()), salt=1)    # checkpoint-1000, not -500
    for k, v in want.items():
        assert torch.equal(out["controlnet"][k], v), k
    want = W.synth_state_dict(W.unet_shapes(UNetConfig.tiny()))
    assert all(torch.equal(out["unet"][k], v) for k, v in want.items())


def test_schema_mismatch_fails_loudly(dirs):
    sd = W.synth_state_dict(W.unet_shapes(UNetConfig.tiny()))
    sd.pop("conv_in.weight")
    with pytest.raises(ValueError):
        ckpt.check_schema(sd, W.unet_shapes(UNetConfig.tiny()), "U-Net")
    with pytest.raises(FileNotFoundError):
        ckpt.load_sd_turbo("stabilityai/sd-turbo", dirs["diffusion_ckpt"])   # hub ids cannot resolve offline
    with pytest.raises(NotImplementedError):
        ckpt.unet_config_from_json({"use_linear_projection": False})
    with pytest.raises(NotImplementedError):
        ckpt.scheduler_config_from_json({"_class_name": "DPMSolverMultistepScheduler"}) and None or \
            __import__("genima_b200.scheduler", fromlist=["x"]).EulerDiscreteSchedule(
                ckpt.scheduler_config_from_json({"_class_name": "DPMSolverMultistepScheduler"}))


def test_controller_snapshot_round_trip(dirs):
    sd = ckpt.load_controller_snapshot(os.path.join(dirs["controller_ckpt"], "latest.pt"), ACTConfig.tiny())
    want = W.synth_state_dict(W.act_shapes(ACTConfig.tiny()), salt=3)
    assert set(sd) == set(want) and all(torch.equal(sd[k], want[k]) for k in want)


def test_pix2pix_checkpoint_round_trip(dirs):
    """controller/agent/sd_pix2pix_agent.py:19-41: U-Net from <diffusion_ckpt>/checkpoint-*/unet, the rest from sd_ckpt."""
    import dataclasses

    assert ckpt.find_pix2pix_unet_dir(dirs["pix2pix_ckpt"]).endswith(os.path.join("checkpoint-200", "unet"))
    out = ckpt.load_sd_pix2pix(dirs["sd_ckpt"], dirs["pix2pix_ckpt"])
    ucfg = dataclasses.replace(UNetConfig.tiny(), in_channels=8)
    # (cond_embed_channels is a ControlNet property: a plain U-Net's config.json does not carry it)
    assert dataclasses.replace(out["unet_cfg"], cond_embed_channels=ucfg.cond_embed_channels) == ucfg
    assert out["controlnet"] is None
    want = W.synth_state_dict(W.unet_shapes(ucfg), salt=5)
    assert all(torch.equal(out["unet"][k], v) for k, v in want.items())
    assert out["unet"]["conv_in.weight"].shape[1] == 8
    enc = W.synth_state_dict(W.vae_encoder_shapes(VAEConfig.tiny()), salt=2)
    assert all(torch.equal(out["vae"][k], v) for k, v in enc.items())
    with pytest.raises(ValueError):               # a 4-channel U-Net is not an InstructPix2Pix U-Net
        ckpt.load_sd_pix2pix(dirs["sd_ckpt"], os.path.join(dirs["sd_ckpt"]))


def test_sdxl_snapshot_round_trip(tmp_path):
    """controller/agent/sdxl_controlnet_agent.py:19-42 layout: SDXL U-Net / ControlNet config fields, text_encoder_2 with
    projection, Euler-ancestral scheduler."""
    from genima_b200.scheduler import EulerDiscreteSchedule

    ucfg, t2 = UNetConfig.sdxl_tiny(), CLIPTextConfig.tiny(projection_dim=64)
    d = ckpt.save_synthetic_checkpoints(str(tmp_path), ucfg, VAEConfig.tiny(), CLIPTextConfig.tiny(), ACTConfig.tiny(),
                                        text2_cfg=t2)
    out = ckpt.load_sdxl(d["sd_ckpt"], d["diffusion_ckpt"])
    assert out["unet_cfg"] == ucfg and out["text2_cfg"] == t2 and out["text_cfg"].projection_dim == 0
    assert EulerDiscreteSchedule(out["scheduler_cfg"]).ancestral
    want = W.synth_state_dict(W.clip_text_shapes(t2), salt=4)
    assert all(torch.equal(out["text2"][k], v) for k, v in want.items())
    assert "add_embedding.linear_1.weight" in out["controlnet"]
    # stabilityai/sdxl-turbo's VAE config says force_upcast: true (diffusers decodes it in fp32): the flag must survive
    assert out["vae_cfg"].force_upcast is False
    import json
    vj = os.path.join(d["sd_ckpt"], "vae", "config.json")
    with open(vj) as f:
        j = json.load(f)
    with open(vj, "w") as f:
        json.dump(dict(j, force_upcast=True), f)
    assert ckpt.load_sdxl(d["sd_ckpt"], d["diffusion_ckpt"])["vae_cfg"].force_upcast is True
    with pytest.raises(ValueError):                # an SD-2.x snapshot is not an SDXL snapshot
        d2 = ckpt.save_synthetic_checkpoints(str(tmp_path / "sd2"), UNetConfig.tiny(), VAEConfig.tiny(),
                                             CLIPTextConfig.tiny(), ACTConfig.tiny())
        ckpt.load_sdxl(d2["sd_ckpt"], d2["diffusion_ckpt"])


def test_snapshot_tokenizer_is_picked_up(dirs):
    """String prompts (what controller/eval_genima.py:178 passes) tokenize through the snapshot's own tokenizer files the
    way diffusers' encode_prompt does: BOS, BPE ids, EOS, padded to 77."""
    assert ckpt.load_sd_turbo(dirs["sd_ckpt"], dirs["diffusion_ckpt"])["tokenizer"] is None      # no tokenizer/ directory
    vocab = write_tiny_clip_tokenizer(os.path.join(dirs["sd_ckpt"], "tokenizer"))
    tok = ckpt.load_sd_turbo(dirs["sd_ckpt"], dirs["diffusion_ckpt"])["tokenizer"]
    ids = tok(["open the box", "the box"])
    assert ids.shape == (2, 77) and ids.dtype == torch.int64
    bos, eos = vocab["<|startoftext|>"], vocab["<|endoftext|>"]
    assert ids[0, :5].tolist() == [bos, vocab["open</w>"], vocab["the</w>"], vocab["box</w>"], eos]
    assert ids[1, :4].tolist() == [bos, vocab["the</w>"], vocab["box</w>"], eos] and int(ids[1, 4:].min()) == eos
