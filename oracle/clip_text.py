"""fp32 CPU restatement of the CLIP text towers on the hot path (TEST INFRASTRUCTURE — see oracle/__init__).

  * SD-Turbo `text_encoder` (transformers CLIPTextModel, OpenCLIP-H text: 23 layers, d 1024, GELU) whose
    last_hidden_state conditions the U-Net / ControlNet cross-attention (diffusers encode_prompt);
  * OpenAI CLIP ViT-B/32 text tower used by GenimaACT.encode_clip_text (controller/method/genima_act.py:314-346):
    token + positional embedding -> 12 pre-LN layers (QuickGELU, causal mask) -> ln_final -> EOT row @ text_projection.
Operates on transformers' CLIPTextModel key names; PINNED against transformers.CLIPTextModel in tests/test_oracle_pins.py.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

CLIPTextConfig = object   # duck-typed: oracle.configs.text_from_json(...) or any object with the same attributes


def clip_text_forward(sd: Dict[str, torch.Tensor], cfg: CLIPTextConfig, ids: torch.Tensor, penultimate: bool = False):
    """ids: [B, T] int64.  Returns (last_hidden_state [B, T, d], pooled-projected [B, proj] or None); with
    penultimate=True the first item is transformers' `hidden_states[-2]` instead (the input of the last layer, no final
    LayerNorm), which is what diffusers' SDXL encode_prompt feeds to the U-Net."""
    w = lambda k: sd[k].to(torch.float32)  # noqa: E731
    b, t = ids.shape
    d = cfg.hidden_size
    h = w("text_model.embeddings.token_embedding.weight")[ids] + w("text_model.embeddings.position_embedding.weight")[:t]
    mask = torch.full((t, t), float("-inf")).triu(1)
    hd = d // cfg.num_heads
    pen = None
    for i in range(cfg.num_layers):
        if i == cfg.num_layers - 1:
            pen = h
        p = f"text_model.encoder.layers.{i}"
        n = F.layer_norm(h, (d,), w(f"{p}.layer_norm1.weight"), w(f"{p}.layer_norm1.bias"), cfg.eps)
        q = F.linear(n, w(f"{p}.self_attn.q_proj.weight"), w(f"{p}.self_attn.q_proj.bias"))
        k = F.linear(n, w(f"{p}.self_attn.k_proj.weight"), w(f"{p}.self_attn.k_proj.bias"))
        v = F.linear(n, w(f"{p}.self_attn.v_proj.weight"), w(f"{p}.self_attn.v_proj.bias"))
        q = q.reshape(b, t, cfg.num_heads, hd).transpose(1, 2)
        k = k.reshape(b, t, cfg.num_heads, hd).transpose(1, 2)
        v = v.reshape(b, t, cfg.num_heads, hd).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) / math.sqrt(hd) + mask
        o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(b, t, d)
        h = h + F.linear(o, w(f"{p}.self_attn.out_proj.weight"), w(f"{p}.self_attn.out_proj.bias"))
        n = F.layer_norm(h, (d,), w(f"{p}.layer_norm2.weight"), w(f"{p}.layer_norm2.bias"), cfg.eps)
        m = F.linear(n, w(f"{p}.mlp.fc1.weight"), w(f"{p}.mlp.fc1.bias"))
        m = m * torch.sigmoid(1.702 * m) if cfg.act == "quick_gelu" else F.gelu(m)
        h = h + F.linear(m, w(f"{p}.mlp.fc2.weight"), w(f"{p}.mlp.fc2.bias"))
    h = F.layer_norm(h, (d,), w("text_model.final_layer_norm.weight"), w("text_model.final_layer_norm.bias"), cfg.eps)
    pooled = None
    if cfg.projection_dim:
        eot = ids.argmax(dim=-1)  # OpenAI CLIP: the EOT token has the largest id
        pooled = F.linear(h[torch.arange(b), eot], w("text_projection.weight"))
    return (pen if penultimate else h), pooled
