"""Pins the CPU oracle (oracle/) against independent implementations that ARE available offline:
  * CLIP text towers      vs transformers.CLIPTextModel (the class the reference's diffusers pipeline instantiates) and
                          transformers.CLIPTextModelWithProjection (SDXL text_encoder_2: hidden_states[-2], text_embeds)
  * ResNet-18 trunk       vs torchvision.models.resnet18 with FrozenBatchNorm2d (FiLM projections zeroed -> identity)
  * multi-head attention  vs torch.nn.MultiheadAttention (what the DETR/ACT transformer layers wrap)
  * DETR encoder layer    vs torch.nn.TransformerEncoderLayer(norm_first=False) with zero positional input
  * tile / untile         vs the reference's own controller/utils/misc.py outputs (tests/golden/tiling.json)
The diffusers U-Net / ControlNet / VAE graphs have no offline second implementation: PARITY UNPINNED for those
(oracle/__init__.py); they are pinned structurally by the parameter counts in test_weights_schema.py."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from genima_b200 import weights as W
from genima_b200.configs import ACTConfig, CLIPTextConfig

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("act", ["gelu", "quick_gelu"])
def test_clip_text_oracle_matches_transformers(act):
    from transformers import CLIPTextConfig as HFConfig
    from transformers import CLIPTextModel

    from oracle.clip_text import clip_text_forward

    cfg = CLIPTextConfig(vocab_size=1000, hidden_size=128, intermediate_size=256, num_layers=3, num_heads=4, act=act)
    sd = {k: v.float() for k, v in W.synth_state_dict(W.clip_text_shapes(cfg)).items()}
    hf = CLIPTextModel(HFConfig(vocab_size=1000, hidden_size=128, intermediate_size=256, num_hidden_layers=3,
                                num_attention_heads=4, max_position_embeddings=77, hidden_act=act,
                                eos_token_id=999, bos_token_id=998, pad_token_id=0)).eval()
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in m for m in missing)
    ids = torch.zeros(2, 77, dtype=torch.int64)
    g = torch.Generator().manual_seed(5)
    for b in range(2):
        ids[b, 0] = 998
        ids[b, 1:10 + b] = torch.randint(1, 990, (9 + b,), generator=g)
        ids[b, 10 + b] = 999
    with torch.no_grad():
        ref = hf(input_ids=ids).last_hidden_state
    out, _ = clip_text_forward(sd, cfg, ids)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5), float((out - ref).abs().max())


def test_sdxl_text_conditioning_matches_transformers_with_projection():
    """What diffusers' SDXL encode_prompt reads off text_encoder_2 (transformers.CLIPTextModelWithProjection with
    output_hidden_states=True): hidden_states[-2] and text_embeds — the oracle's penultimate / pooled outputs."""
    from transformers import CLIPTextConfig as HFConfig
    from transformers import CLIPTextModelWithProjection

    from oracle.clip_text import clip_text_forward

    cfg = CLIPTextConfig(vocab_size=1000, hidden_size=128, intermediate_size=256, num_layers=4, num_heads=2, act="gelu",
                         projection_dim=64)
    sd = {k: v.float() for k, v in W.synth_state_dict(W.clip_text_shapes(cfg), salt=4).items()}
    hf = CLIPTextModelWithProjection(HFConfig(
        vocab_size=1000, hidden_size=128, intermediate_size=256, num_hidden_layers=4, num_attention_heads=2,
        max_position_embeddings=77, hidden_act="gelu", projection_dim=64, eos_token_id=999, bos_token_id=998,
        pad_token_id=0)).eval()
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in m for m in missing)
    ids = torch.zeros(2, 77, dtype=torch.int64)
    g = torch.Generator().manual_seed(6)
    for b in range(2):
        ids[b, 0] = 998
        ids[b, 1:12 + b] = torch.randint(1, 990, (11 + b,), generator=g)
        ids[b, 12 + b] = 999
    with torch.no_grad():
        ref = hf(input_ids=ids, output_hidden_states=True)
    pen, pooled = clip_text_forward(sd, cfg, ids, penultimate=True)
    assert torch.allclose(pen, ref.hidden_states[-2], rtol=1e-4, atol=1e-5)
    assert torch.allclose(pooled, ref.text_embeds, rtol=1e-4, atol=1e-5)
    assert not torch.allclose(pen, ref.last_hidden_state, rtol=1e-2, atol=1e-3)


def test_clip_pooled_projection_takes_eot_row():
    from oracle.clip_text import clip_text_forward

    cfg = CLIPTextConfig.tiny(projection_dim=64)
    sd = {k: v.float() for k, v in W.synth_state_dict(W.clip_text_shapes(cfg)).items()}
    ids = torch.zeros(1, 77, dtype=torch.int64)
    ids[0, :5] = torch.tensor([998, 3, 4, 5, 999])
    h, pooled = clip_text_forward(sd, cfg, ids)
    assert torch.allclose(pooled[0], sd["text_projection.weight"] @ h[0, 4], rtol=1e-5, atol=1e-6)


def test_resnet18_trunk_oracle_matches_torchvision():
    import torchvision

    from oracle.act import resnet18_film

    cfg = ACTConfig()
    sd = {k: v.float() for k, v in W.synth_state_dict(W.act_shapes(cfg), salt=3).items()}
    for k in sd:
        if ".film." in k:
            sd[k] = torch.zeros_like(sd[k])            # FiLM off: (1 + 0) * x + 0
    tv = torchvision.models.resnet18(weights=None, norm_layer=torchvision.ops.FrozenBatchNorm2d).eval()
    tv_sd = {k[len("encoder_model.backbone."):]: v for k, v in sd.items()
             if k.startswith("encoder_model.backbone.") and ".film." not in k}
    missing, unexpected = tv.load_state_dict(tv_sd, strict=False)
    assert not unexpected and set(missing) == {"fc.weight", "fc.bias"}
    x = torch.randn(2, 3, 96, 96, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        h = tv.maxpool(tv.relu(tv.bn1(tv.conv1(x))))
        ref = tv.layer4(tv.layer3(tv.layer2(tv.layer1(h))))
    out = resnet18_film(sd, cfg, x, torch.zeros(2, cfg.task_emb_dim))
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4), float((out - ref).abs().max())


def test_mha_oracle_matches_torch_multihead_attention():
    from oracle.act import mha

    d, nh = 64, 4
    m = torch.nn.MultiheadAttention(d, nh).eval()
    sd = {f"a.{k}": v.detach() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    q, k, v = torch.randn(5, 2, d, generator=g), torch.randn(9, 2, d, generator=g), torch.randn(9, 2, d, generator=g)
    with torch.no_grad():
        ref = m(q, k, v, need_weights=False)[0]
    assert torch.allclose(mha(sd, "a", q, k, v, nh), ref, rtol=1e-4, atol=1e-5)


def test_detr_encoder_layer_matches_torch_when_pos_is_zero():
    """oracle/act.py's encoder layer is DETR's post-norm layer; with pos = 0 it must equal nn.TransformerEncoderLayer."""
    from oracle import act as A

    d, nh, ff = 64, 4, 128
    layer = torch.nn.TransformerEncoderLayer(d, nh, ff, dropout=0.0, activation="relu", norm_first=False).eval()
    sd = {f"L.{k}": v.detach() for k, v in layer.state_dict().items()}
    src = torch.randn(7, 2, d, generator=torch.Generator().manual_seed(2))
    out = A._ln(sd, "L.norm1", src + A.mha(sd, "L.self_attn", src, src, src, nh), 1e-5)
    out = A._ln(sd, "L.norm2", out + A._ffn(sd, "L", out), 1e-5)
    with torch.no_grad():
        ref = layer(src)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5)


def _regen_inputs():
    rng = np.random.RandomState(0)
    views = rng.randint(0, 256, size=(4, 2, 256, 256, 3), dtype=np.uint8)
    gen = rng.randint(0, 256, size=(2, 512, 512, 3), dtype=np.uint8)
    return views, gen


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_tiling_oracle_and_host_glue_match_reference_golden():
    """tests/golden/tiling.json was produced by the reference's own tile_images / untile_images."""
    from PIL import Image

    from genima_b200.host_glue import tile_images, untile_images
    from oracle import tiling

    with open(os.path.join(HERE, "golden", "tiling.json")) as f:
        gold = json.load(f)
    views, gen = _regen_inputs()
    assert _sha(views) == gold["views"]["sha256"] and _sha(gen) == gold["gen"]["sha256"]
    # oracle (numpy) restatement
    tiles = np.stack([tiling.tile_views(views[:, t]) for t in range(2)])
    assert _sha(tiles) == gold["tiles"]["sha256"]
    for ci, cam in enumerate(gold["cameras"]):
        un = np.stack([np.transpose(tiling.untile_views(gen[t])[ci], (2, 0, 1)) for t in range(2)])
        assert list(un.shape) == gold["untiled"][cam]["shape"] and _sha(un) == gold["untiled"][cam]["sha256"]
    # host-side mirror of the reference API (PIL in / PIL out)
    rgbs = [Image.fromarray(views[c, t]) for c in range(4) for t in range(2)]
    pil_tiles = tile_images(rgbs, 2)
    assert _sha(np.stack([np.asarray(t) for t in pil_tiles])) == gold["tiles"]["sha256"]
    from genima_b200.agents import _CenterResize

    un = untile_images([Image.fromarray(g) for g in gen], gold["cameras"], _CenterResize(256))
    for cam in gold["cameras"]:
        assert un[cam].dtype == np.uint8 and _sha(un[cam]) == gold["untiled"][cam]["sha256"]


def test_tile_untile_round_trip_property():
    from oracle import tiling

    v = np.random.RandomState(3).randint(0, 256, size=(4, 64, 64, 3), dtype=np.uint8)
    assert np.array_equal(tiling.untile_views(tiling.tile_views(v)), v)


# ------------------------------------------------------------------------------------ the oracle's own architecture
class _Recording(dict):
    """State dict that records which keys the oracle graph reads."""

    def __init__(self, sd):
        super().__init__(sd)
        self.touched = set()

    def __getitem__(self, k):
        self.touched.add(k)
        return super().__getitem__(k)

    def __contains__(self, k):
        return super().__contains__(k)


def test_oracle_configs_are_independent_literals_that_reproduce_the_published_counts():
    """oracle/configs.py states the upstream config.json values itself and counts parameters in closed form (no key
    table): the published model sizes come out exactly, and the product's schemas agree with them."""
    from genima_b200.configs import CLIPTextConfig as PText
    from genima_b200.configs import UNetConfig, VAEConfig
    from oracle import configs as oc

    ucfg, vcfg, tcfg = oc.sd_turbo()
    assert oc.count_unet(ucfg) == 865_910_724                    # SD-2.1 UNet2DConditionModel
    assert oc.count_controlnet(ucfg) == 364_228_240              # ControlNetModel.from_unet(SD-2.1)
    assert oc.count_vae_decoder(vcfg) == 49_490_199              # AutoencoderKL decoder + post_quant_conv
    assert oc.count_clip_text(tcfg) == 340_387_840               # OpenCLIP ViT-H/14 text tower, 23 layers
    assert oc.count_clip_text(oc.text_from_json(oc.CLIP_VIT_B32_TEXT)) == 63_428_096      # CLIP ViT-B/32 text tower
    assert oc.count_unet(oc.unet_from_json(oc.SDXL_UNET_JSON)) == 2_567_463_684            # SDXL U-Net
    # the product's table-driven schemas (a different derivation) give the same totals ...
    assert W.count_params(W.unet_shapes(UNetConfig())) == oc.count_unet(ucfg)
    assert W.count_params(W.controlnet_shapes(UNetConfig())) == oc.count_controlnet(ucfg)
    assert W.count_params(W.vae_decoder_shapes(VAEConfig())) == oc.count_vae_decoder(vcfg)
    assert W.count_params(W.clip_text_shapes(PText.sd_turbo())) == oc.count_clip_text(tcfg)
    assert W.count_params(W.unet_shapes(UNetConfig.sdxl())) == oc.count_unet(oc.unet_from_json(oc.SDXL_UNET_JSON))
    # ... and field for field the same hyper-parameters
    for name in ("in_channels", "out_channels", "block_out_channels", "layers_per_block", "num_heads", "attn_levels",
                 "cross_attention_dim", "norm_num_groups", "norm_eps", "cond_embed_channels", "addition_embed"):
        assert getattr(ucfg, name) == getattr(UNetConfig(), name), name
    for name in ("latent_channels", "out_channels", "block_out_channels", "layers_per_block", "norm_num_groups",
                 "norm_eps", "scaling_factor"):
        assert getattr(vcfg, name) == getattr(VAEConfig(), name), name
    acfg = oc.act_from_dict(oc.GENIMA_ACT)
    for name in ("hidden_dim", "enc_layers", "dec_layers", "dim_feedforward", "nheads", "num_queries", "state_dim",
                 "action_dim", "latent_dim", "num_views", "image_size", "task_emb_dim", "resnet_widths"):
        assert getattr(acfg, name) == getattr(ACTConfig(), name), name


def test_oracle_graphs_touch_exactly_the_published_parameters():
    """Trace the oracle's U-Net / ControlNet / VAE / ACT graphs (built from oracle/configs.py, NOT the product's config
    classes) with a recording state dict at a reduced width: every parameter of the schema is read, none is invented —
    so the graph that defines parity uses exactly the tensors whose count equals the published model size."""
    from oracle import act as act_oracle
    from oracle import configs as oc
    from oracle import sd_models

    j = dict(oc.SD_TURBO_UNET_JSON, block_out_channels=[64, 128, 128, 128], attention_head_dim=[1, 2, 2, 2],
             cross_attention_dim=128)
    ucfg = oc.unet_from_json(j, cond_channels=[16, 32, 64, 64])
    vcfg = oc.vae_from_json(dict(oc.SD_TURBO_VAE_JSON, block_out_channels=[64, 64, 128, 128]))
    acfg = oc.act_from_dict(dict(oc.GENIMA_ACT, hidden_dim=64, enc_layers=1, dec_layers=2, dim_feedforward=128, nheads=2,
                                 num_queries=4, image_size=64, task_emb_dim=64, resnet_widths=[64, 64, 64, 64]))
    usd = _Recording({k: v.float() for k, v in W.synth_state_dict(W.unet_shapes(ucfg)).items()})
    csd = _Recording({k: v.float() for k, v in W.synth_state_dict(W.controlnet_shapes(ucfg), 1).items()})
    vsd = _Recording({k: v.float() for k, v in W.synth_state_dict(W.vae_decoder_shapes(vcfg), 2).items()})
    asd = _Recording({k: v.float() for k, v in W.synth_state_dict(W.act_shapes(acfg), 3).items()})
    assert sum(v.numel() for v in usd.values()) == oc.count_unet(ucfg)
    assert sum(v.numel() for v in csd.values()) == oc.count_controlnet(ucfg)
    assert sum(v.numel() for v in vsd.values()) == oc.count_vae_decoder(vcfg)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, 8, 8, generator=g)
    ctx = torch.randn(1, 77, 128, generator=g)
    cond = torch.rand(1, 3, 64, 64, generator=g)
    t = torch.tensor([999.0])
    with torch.no_grad():
        down, mid = sd_models.controlnet_forward(csd, ucfg, x, t, ctx, cond, 1.0)
        sd_models.unet_forward(usd, ucfg, x, t, ctx, down, mid)
        sd_models.vae_decode(vsd, vcfg, x)
        act_oracle.act_forward(asd, acfg, torch.randn(1, 8, generator=g),
                               torch.rand(1, 4, 3, 64, 64, generator=g) * 255, torch.randn(1, 64, generator=g))
    assert usd.touched == set(usd.keys())
    assert csd.touched == set(csd.keys())
    assert vsd.touched == set(vsd.keys())
    # ACT: the inference branch never reads the CVAE encoder side; everything in the schema is inference-side
    assert asd.touched == set(asd.keys())
