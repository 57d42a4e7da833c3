"""fp32 CPU restatement of the Genima ACT controller forward (TEST INFRASTRUCTURE — see oracle/__init__).

In-repo anchors: controller/method/genima_act.py:165-214 (GenimaACTPolicy.forward: /255, ImageNet normalise, encoder,
actor), :27-92 (GenimaMVTransformer.forward: proprio MLP, zero latent, transformer(...)[-1], heads), :221-241
(build_actor: 2-layer proprio projection), controller/cfgs/method/genima_act.yaml:13-39 (hyper-parameters).
RoboBase itself (ImageEncoderACT, MultiViewTransformerEncoderDecoderACT) is neither vendored nor pinned (README.md:40-46),
so its graph is restated from the ACT / DETR architecture it wraps: torchvision ResNet-18 trunk
with FrozenBatchNorm2d and FiLM on the task embedding, 1x1 input projection, DETR sine position embedding per view,
views concatenated along width, post-norm DETR encoder/decoder.  PARITY UNPINNED for the RoboBase wiring; the
ResNet-18 trunk and multi-head attention are pinned against torchvision / torch.nn in tests/test_oracle_pins.py.
Open items (cannot be verified offline): FiLM form ((1 + gamma) * x + beta, after bn2), no task token in the encoder
sequence, latent_dim = 32.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

ACTConfig = object   # duck-typed: oracle.configs.act_from_dict(...) or any object with the same attributes

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


_CONSTANTS = {}


def _imagenet_constants(device):
    """(mean, std) broadcastable over [B, V, 3, H, W], created once per device (a host -> device copy cannot be captured
    into a CUDA graph, which the stock-PyTorch GPU baseline does with this function)."""
    key = str(device)
    if key not in _CONSTANTS:
        _CONSTANTS[key] = (torch.tensor(IMAGENET_MEAN, device=device)[None, None, :, None, None],
                           torch.tensor(IMAGENET_STD, device=device)[None, None, :, None, None])
    return _CONSTANTS[key]


def position_embedding_sine(h: int, w: int, num_pos_feats: int, temperature: float = 10000.0) -> torch.Tensor:
    """DETR PositionEmbeddingSine(normalize=True, scale=2*pi) for an unmasked [h, w] map -> [1, 2*npf, h, w]."""
    eps, scale = 1e-6, 2 * math.pi
    y_embed = torch.arange(1, h + 1, dtype=torch.float32)[:, None].expand(h, w)
    x_embed = torch.arange(1, w + 1, dtype=torch.float32)[None, :].expand(h, w)
    y_embed = y_embed / (h + eps) * scale
    x_embed = x_embed / (w + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    pos_x = x_embed[:, :, None] / dim_t
    pos_y = y_embed[:, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, 0::2].sin(), pos_x[:, :, 1::2].cos()), dim=3).flatten(2)
    pos_y = torch.stack((pos_y[:, :, 0::2].sin(), pos_y[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((pos_y, pos_x), dim=2).permute(2, 0, 1)[None]


def frozen_bn(sd, p, x, eps):
    w = sd[f"{p}.weight"].float()
    b = sd[f"{p}.bias"].float()
    rm = sd[f"{p}.running_mean"].float()
    rv = sd[f"{p}.running_var"].float()
    scale = w * (rv + eps).rsqrt()
    shift = b - rm * scale
    return x * scale[None, :, None, None] + shift[None, :, None, None]


def resnet18_film(sd: Dict[str, torch.Tensor], cfg: ACTConfig, x: torch.Tensor, task_emb: torch.Tensor):
    """x: [N, 3, H, W] normalised; task_emb: [N, task_emb_dim] -> layer4 features [N, widths[-1], H/32, W/32]."""
    b = "encoder_model.backbone"
    h = F.conv2d(x, sd[f"{b}.conv1.weight"].float(), stride=2, padding=3)
    h = F.relu(frozen_bn(sd, f"{b}.bn1", h, cfg.bn_eps))
    h = F.max_pool2d(h, 3, 2, 1)
    for li, cout in enumerate(cfg.resnet_widths):
        for bi in range(2):
            p = f"{b}.layer{li + 1}.{bi}"
            stride = 2 if (li > 0 and bi == 0) else 1
            idt = h
            o = F.conv2d(h, sd[f"{p}.conv1.weight"].float(), stride=stride, padding=1)
            o = F.relu(frozen_bn(sd, f"{p}.bn1", o, cfg.bn_eps))
            o = F.conv2d(o, sd[f"{p}.conv2.weight"].float(), padding=1)
            o = frozen_bn(sd, f"{p}.bn2", o, cfg.bn_eps)
            film = F.linear(task_emb, sd[f"{p}.film.weight"].float(), sd[f"{p}.film.bias"].float())
            gamma, beta = film[:, :cout], film[:, cout:]
            o = (1 + gamma)[:, :, None, None] * o + beta[:, :, None, None]
            if f"{p}.downsample.0.weight" in sd:
                idt = F.conv2d(h, sd[f"{p}.downsample.0.weight"].float(), stride=stride)
                idt = frozen_bn(sd, f"{p}.downsample.1", idt, cfg.bn_eps)
            h = F.relu(o + idt)
    return h


def mha(sd, p, q, k, v, nheads):
    """torch.nn.MultiheadAttention forward (batch_first=False layout [T, B, C]) from its packed parameters."""
    d = q.shape[-1]
    wi = sd[f"{p}.in_proj_weight"].float()
    bi = sd[f"{p}.in_proj_bias"].float()
    qp = F.linear(q, wi[:d], bi[:d])
    kp = F.linear(k, wi[d:2 * d], bi[d:2 * d])
    vp = F.linear(v, wi[2 * d:], bi[2 * d:])
    tq, bsz, _ = qp.shape
    tk = kp.shape[0]
    hd = d // nheads
    qh = qp.reshape(tq, bsz, nheads, hd).permute(1, 2, 0, 3)
    kh = kp.reshape(tk, bsz, nheads, hd).permute(1, 2, 0, 3)
    vh = vp.reshape(tk, bsz, nheads, hd).permute(1, 2, 0, 3)
    s = (qh @ kh.transpose(-1, -2)) / math.sqrt(hd)
    o = (torch.softmax(s, dim=-1) @ vh).permute(2, 0, 1, 3).reshape(tq, bsz, d)
    return F.linear(o, sd[f"{p}.out_proj.weight"].float(), sd[f"{p}.out_proj.bias"].float())


def _ln(sd, p, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[f"{p}.weight"].float(), sd[f"{p}.bias"].float(), eps)


def _ffn(sd, p, x):
    h = F.relu(F.linear(x, sd[f"{p}.linear1.weight"].float(), sd[f"{p}.linear1.bias"].float()))
    return F.linear(h, sd[f"{p}.linear2.weight"].float(), sd[f"{p}.linear2.bias"].float())


def act_forward(sd: Dict[str, torch.Tensor], cfg: ACTConfig, qpos: torch.Tensor, image: torch.Tensor,
                task_emb: torch.Tensor):
    """qpos [B, state_dim]; image [B, V, 3, H, W] in 0..255 (float); task_emb [B, task_emb_dim].
    Returns (a_hat [B, num_queries, action_dim], is_pad_hat [B, num_queries, 1])."""
    bsz, nv = image.shape[:2]
    d = cfg.hidden_dim
    mean, std = _imagenet_constants(image.device)
    img = (image.float() / 255.0 - mean) / std                     # genima_act.py:188
    feats, poss = [], []
    for v in range(nv):
        f = resnet18_film(sd, cfg, img[:, v], task_emb.float())
        f = F.conv2d(f, sd["encoder_model.input_proj.weight"].float(), sd["encoder_model.input_proj.bias"].float())
        feats.append(f)
        poss.append(position_embedding_sine(f.shape[2], f.shape[3], d // 2))
    feat = torch.cat(feats, dim=3)                                  # views along width
    pos = torch.cat(poss, dim=3)
    a = "actor_model"
    h0 = F.linear(qpos.float(), sd[f"{a}.input_proj_robot_state.0.weight"].float(),
                  sd[f"{a}.input_proj_robot_state.0.bias"].float())
    proprio = F.linear(h0, sd[f"{a}.input_proj_robot_state.2.weight"].float(),
                       sd[f"{a}.input_proj_robot_state.2.bias"].float())       # Dropout(0.3) is identity in eval
    latent = F.linear(torch.zeros(bsz, cfg.latent_dim), sd[f"{a}.latent_out_proj.weight"].float(),
                      sd[f"{a}.latent_out_proj.bias"].float())                 # genima_act.py:71-75
    src = feat.flatten(2).permute(2, 0, 1)
    pos_seq = pos.flatten(2).permute(2, 0, 1).repeat(1, bsz, 1)
    add_pos = sd[f"{a}.additional_pos_embed.weight"].float()[:, None, :].repeat(1, bsz, 1)
    pos_seq = torch.cat([add_pos, pos_seq], dim=0)
    src = torch.cat([torch.stack([latent, proprio], dim=0), src], dim=0)
    for i in range(cfg.enc_layers):
        p = f"{a}.transformer.encoder.layers.{i}"
        qk = src + pos_seq
        src = _ln(sd, f"{p}.norm1", src + mha(sd, f"{p}.self_attn", qk, qk, src, cfg.nheads), cfg.ln_eps)
        src = _ln(sd, f"{p}.norm2", src + _ffn(sd, p, src), cfg.ln_eps)
    memory = src
    query_pos = sd[f"{a}.query_embed.weight"].float()[:, None, :].repeat(1, bsz, 1)
    tgt = torch.zeros_like(query_pos)
    for i in range(cfg.dec_layers):
        p = f"{a}.transformer.decoder.layers.{i}"
        qk = tgt + query_pos
        tgt = _ln(sd, f"{p}.norm1", tgt + mha(sd, f"{p}.self_attn", qk, qk, tgt, cfg.nheads), cfg.ln_eps)
        tgt = _ln(sd, f"{p}.norm2",
                  tgt + mha(sd, f"{p}.multihead_attn", tgt + query_pos, memory + pos_seq, memory, cfg.nheads),
                  cfg.ln_eps)
        tgt = _ln(sd, f"{p}.norm3", tgt + _ffn(sd, p, tgt), cfg.ln_eps)
    hs = _ln(sd, f"{a}.transformer.decoder.norm", tgt, cfg.ln_eps).transpose(0, 1)   # [-1] of the intermediate stack
    a_hat = F.linear(hs, sd[f"{a}.action_head.weight"].float(), sd[f"{a}.action_head.bias"].float())
    is_pad = F.linear(hs, sd[f"{a}.is_pad_head.weight"].float(), sd[f"{a}.is_pad_head.bias"].float())
    return a_hat, is_pad
