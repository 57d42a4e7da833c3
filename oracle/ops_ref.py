"""fp32 CPU references for the individual fused kernels (the per-op parity oracle).

Each function mirrors one C-ABI entry point of include/genima_b200.h and is written with torch.nn.functional
primitives only, in the tensor layout the upstream libraries use (NCHW for images), so it reads like the upstream
module it stands for:
  linear_ref      torch.nn.Linear (+ the bias / activation / residual that follows it upstream)
  conv2d_ref      torch.nn.Conv2d as used by diffusers ResnetBlock2D / Downsample2D / Upsample2D, torchvision ResNet
  group_norm_ref  torch.nn.GroupNorm (+ SiLU)             diffusers resnet.py ResnetBlock2D.norm1/norm2
  layer_norm_ref  torch.nn.LayerNorm                      diffusers attention.py BasicTransformerBlock.norm1-3
  attention_ref   F.scaled_dot_product_attention math     diffusers attention_processor.py AttnProcessor2_0
  geglu_ref       diffusers activations.py GEGLU
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

ACTS = {
    None: lambda x: x,
    "none": lambda x: x,
    "silu": F.silu,
    "gelu": lambda x: F.gelu(x),  # exact erf GELU
    "relu": F.relu,
    "quick_gelu": lambda x: x * torch.sigmoid(1.702 * x),
}


def f32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.detach().to("cpu", torch.float32)


def epilogue_ref(acc, bias=None, scale=None, rowvec=None, rows_per_batch=0, residual=None, act_pre=None,
                 act_post=None, alpha=1.0, beta=1.0):
    """acc: [M, N] fp32.  Mirrors struct gn_epilogue."""
    v = acc
    if scale is not None:
        v = v * f32(scale)[None, :]
    if bias is not None:
        v = v + f32(bias)[None, :]
    if rowvec is not None:
        rv = f32(rowvec)
        v = v + rv.repeat_interleave(rows_per_batch, dim=0)[: v.shape[0]]
    v = ACTS[act_pre](v)
    v = alpha * v
    if residual is not None:
        v = v + beta * f32(residual).reshape(v.shape[0], -1)[:, : v.shape[1]]
    return ACTS[act_post](v)


def linear_ref(a, w, geglu=False, **epi):
    a2 = f32(a).reshape(-1, a.shape[-1])
    acc = a2 @ f32(w).t()
    if geglu:
        # un-interleave 64-wide (value, gate) blocks -> value * gelu(gate)
        bias = epi.pop("bias", None)
        if bias is not None:
            acc = acc + f32(bias)[None, :]
        m, n = acc.shape
        blk = acc.reshape(m, n // 128, 2, 64)
        out = blk[:, :, 0, :] * F.gelu(blk[:, :, 1, :])
        return out.reshape(m, n // 2)
    return epilogue_ref(acc, **epi)


def conv2d_ref(x_nhwc, w_oihw, stride=1, pad=1, extras=(), extra_weights=(), **epi):
    """x_nhwc: [B, H, W, C]; w: [Cout, Cin, KH, KW] (unpacked, Cin may be < C when x carries zero pad channels)."""
    x = f32(x_nhwc).permute(0, 3, 1, 2)
    w = f32(w_oihw)
    x = x[:, : w.shape[1]]
    y = F.conv2d(x, w, stride=stride, padding=pad)
    for ex, ew in zip(extras, extra_weights):
        e = f32(ex).permute(0, 3, 1, 2)
        ew = f32(ew)
        ew = ew.reshape(ew.shape[0], -1, 1, 1)
        y = y + F.conv2d(e[:, : ew.shape[1]], ew)
    b, c, ho, wo = y.shape
    acc = y.permute(0, 2, 3, 1).reshape(b * ho * wo, c)
    epi.setdefault("rows_per_batch", ho * wo)
    out = epilogue_ref(acc, **epi)
    return out.reshape(b, ho, wo, c)


def group_norm_ref(x0, gamma, beta, groups=32, eps=1e-5, silu=False, x1=None):
    x = f32(x0)
    if x1 is not None:
        x = torch.cat([x, f32(x1)], dim=-1)
    shp = x.shape
    xn = x.reshape(shp[0], -1, shp[-1]).permute(0, 2, 1)  # [B, C, HW]
    y = F.group_norm(xn, groups, f32(gamma), f32(beta), eps)
    if silu:
        y = F.silu(y)
    return y.permute(0, 2, 1).reshape(shp)


def layer_norm_ref(x, gamma, beta, eps=1e-5):
    return F.layer_norm(f32(x), (x.shape[-1],), f32(gamma), f32(beta), eps)


def attention_ref(q, k, v, B, heads, head_dim, Tq, Tk, scale, causal=False):
    """q: [B*Tq, heads*head_dim] etc.  Plain softmax(q k^T * scale) v in fp32."""
    qf = f32(q).reshape(B, Tq, heads, head_dim).permute(0, 2, 1, 3)
    kf = f32(k).reshape(B, Tk, heads, head_dim).permute(0, 2, 1, 3)
    vf = f32(v).reshape(B, Tk, heads, head_dim).permute(0, 2, 1, 3)
    s = (qf @ kf.transpose(-1, -2)) * scale
    if causal:
        mask = torch.ones(Tq, Tk, dtype=torch.bool).tril()
        s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = p @ vf
    return o.permute(0, 2, 1, 3).reshape(B * Tq, heads * head_dim)


def timestep_embedding_ref(t: float, dim: int):
    """diffusers embeddings.get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half
    emb = torch.tensor([float(t)], dtype=torch.float32)[:, None] * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)
