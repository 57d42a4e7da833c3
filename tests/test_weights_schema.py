"""State-dict schemas reproduce the published parameter counts of the upstream models, which pins
the restated architectures structurally: a missing / mis-shaped layer changes the total."""
import torch

from genima_b200 import weights as W
from genima_b200.configs import ACTConfig, CLIPTextConfig, UNetConfig, VAEConfig


def test_parameter_counts_match_published_models():
    assert round(W.count_params(W.unet_shapes(UNetConfig())) / 1e6, 1) == 865.9       # SD-2.1 U-Net
    assert round(W.count_params(W.controlnet_shapes(UNetConfig())) / 1e6, 1) == 364.2  # ControlNetModel.from_unet
    assert round(W.count_params(W.vae_decoder_shapes(VAEConfig())) / 1e6, 1) == 49.5   # KL-VAE decoder + post_quant
    assert round(W.count_params(W.clip_text_shapes(CLIPTextConfig.sd_turbo())) / 1e6, 1) == 340.4  # OpenCLIP-H text


def test_clip_text_schema_matches_transformers_module():
    from transformers import CLIPTextConfig as HFConfig
    from transformers import CLIPTextModel

    cfg = CLIPTextConfig.tiny()
    hf = CLIPTextModel(HFConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size,
                                intermediate_size=cfg.intermediate_size, num_hidden_layers=cfg.num_layers,
                                num_attention_heads=cfg.num_heads, max_position_embeddings=cfg.max_positions))
    ours = W.clip_text_shapes(cfg)
    theirs = {k: tuple(v.shape) for k, v in hf.state_dict().items() if "position_ids" not in k}
    assert dict(ours) == theirs


def test_resnet18_trunk_schema_matches_torchvision():
    import torchvision

    tv = torchvision.models.resnet18(weights=None, norm_layer=torchvision.ops.FrozenBatchNorm2d)
    theirs = {f"encoder_model.backbone.{k}": tuple(v.shape) for k, v in tv.state_dict().items()
              if not k.startswith("fc.")}
    ours = {k: v for k, v in W.act_shapes(ACTConfig()).items()
            if k.startswith("encoder_model.backbone.") and ".film." not in k}
    assert ours == theirs


def test_unet_skip_channels_and_decoder_inputs():
    cfg = UNetConfig()
    assert W.unet_skip_channels(cfg) == [320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280]
    s = W.unet_shapes(cfg)
    # decoder ResBlock input widths after the skip concat
    cins = [s[f"up_blocks.{i}.resnets.{j}.conv1.weight"][1] for i in range(4) for j in range(3)]
    assert cins == [2560, 2560, 2560, 2560, 2560, 1920, 1920, 1280, 960, 960, 640, 640]


def test_synthetic_weights_are_deterministic_fp16_and_well_scaled():
    shapes = W.controlnet_shapes(UNetConfig.tiny())
    a, b = W.synth_state_dict(shapes, salt=1), W.synth_state_dict(shapes, salt=1)
    c = W.synth_state_dict(shapes, salt=2)
    for k in shapes:
        assert a[k].dtype == torch.float16 and torch.equal(a[k], b[k])
    assert not torch.equal(a["conv_in.weight"], c["conv_in.weight"])
    for k, v in a.items():
        if v.dim() > 1:      # conv / linear weights ~ N(0, 1/fan_in): activations stay O(1) through the stack
            fan_in = v[0].numel()
            assert 0.5 < float(v.float().std()) * fan_in ** 0.5 < 1.5, k


def test_taesd_decoder_schema_matches_published_size():
    """AutoencoderTiny (madebyollin/taesd) decoder: 1.22 M parameters, 35 convolutions, bias-free 64->64 convs after
    each of the three upsamples."""
    from genima_b200 import weights as W
    from genima_b200.configs import TAESDConfig

    cfg = TAESDConfig()
    shapes = W.taesd_decoder_shapes(cfg)
    n = sum(int(torch.tensor(s).prod()) for s in shapes.values())
    assert n == 1_222_531
    convs = [k for k in shapes if k.endswith(".weight")]
    assert len(convs) == 1 + 10 * 3 + 3 + 1
    no_bias = [k for k in convs if k.replace(".weight", ".bias") not in shapes]
    assert sorted(no_bias) == ["decoder.layers.11.weight", "decoder.layers.16.weight", "decoder.layers.6.weight"]
    assert shapes["decoder.layers.18.weight"] == (3, 64, 3, 3) and shapes["decoder.layers.0.weight"] == (64, 4, 3, 3)


def test_taesd_oracle_runs_and_is_bounded():
    from genima_b200 import weights as W
    from genima_b200.configs import TAESDConfig
    from oracle import sd_models

    cfg = TAESDConfig.tiny()
    sd = W.synth_state_dict(W.taesd_decoder_shapes(cfg), salt=2)
    z = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(0)) * 5.0
    img = sd_models.taesd_decode(sd, cfg, z)
    assert img.shape == (1, 3, 64, 64) and torch.isfinite(img).all()


def test_vae_encoder_schema_and_oracle():
    """AutoencoderKL encoder + quant_conv: 34,163,664 parameters; with the decoder + post_quant_conv the published
    83,653,863 of the SD VAE.  The oracle's Downsample2D pads (0, 1, 0, 1): output is exactly H/8 and depends on the last
    input row/column but the padding never shifts the image."""
    import math

    import torch

    from genima_b200.configs import VAEConfig
    from oracle import sd_models

    n_enc = sum(math.prod(v) for v in W.vae_encoder_shapes(VAEConfig()).values())
    n_dec = sum(math.prod(v) for v in W.vae_decoder_shapes(VAEConfig()).values())
    assert n_enc == 34_163_664 and n_enc + n_dec == 83_653_863
    cfg = VAEConfig.tiny()
    sd = W.synth_state_dict(W.vae_encoder_shapes(cfg), salt=2)
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(0)) * 2 - 1
    z = sd_models.vae_encode_mean(sd, cfg, x)
    assert z.shape == (1, 4, 8, 8) and torch.isfinite(z).all()
    x2 = x.clone()
    x2[..., -1, :] += 0.5                          # the last row is inside the receptive field (bottom padding only)
    assert not torch.equal(sd_models.vae_encode_mean(sd, cfg, x2), z)
    assert torch.equal(sd_models.vae_encode_mean(sd, cfg, x), z)


def test_sdxl_schemas_match_published_sizes():
    """SDXL U-Net 2,567,463,684 parameters (published), CLIP ViT-L text tower 123,060,480, OpenCLIP bigG text tower with
    projection 694,659,840 — the topology restated from memory reproduces all three exactly."""
    import math

    from genima_b200.configs import CLIPTextConfig, UNetConfig

    n = lambda shapes: sum(math.prod(v) for v in shapes.values())  # noqa: E731
    assert n(W.unet_shapes(UNetConfig.sdxl())) == 2_567_463_684
    assert n(W.clip_text_shapes(CLIPTextConfig.sdxl_clip_l())) == 123_060_480
    assert n(W.clip_text_shapes(CLIPTextConfig.sdxl_open_clip_bigg())) == 694_659_840
    cn = W.controlnet_shapes(UNetConfig.sdxl())
    assert "add_embedding.linear_1.weight" in cn and cn["add_embedding.linear_1.weight"] == (1280, 2816)
    assert "down_blocks.2.attentions.1.transformer_blocks.9.ff.net.2.weight" in cn
    assert "down_blocks.0.attentions.0.norm.weight" not in cn          # level 0 is a plain DownBlock2D
