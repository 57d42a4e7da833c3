"""Device-side CLIP text towers (transformers CLIPTextModel graph) on the C-ABI kernels.

  * SD-Turbo text encoder (OpenCLIP-H text, 23 layers): diffusers `encode_prompt` inside pipe(...) —
    controller/agent/sd_controlnet_agent.py:67-76 — run once per distinct prompt and cached (the prompt is constant
    for a whole episode, controller/eval_genima.py:139,178);
  * the two SDXL text encoders (CLIP ViT-L and OpenCLIP bigG with projection; penultimate hidden states + pooled
    projection) for the SDXL-ControlNet sibling (controller/agent/sdxl_controlnet_agent.py);
  * OpenAI CLIP ViT-B/32 text tower of GenimaACT.encode_clip_text (controller/method/genima_act.py:314-346), likewise
    constant per episode.
Token ids are the input: no CLIP BPE vocabulary exists offline, so string -> ids stays upstream.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .configs import CLIPTextConfig
from .ops import Ops
from .unet import _Params


class DeviceCLIPText:
    def __init__(self, ops: Ops, sd: Dict[str, torch.Tensor], cfg: CLIPTextConfig):
        self.ops, self.cfg = ops, cfg
        P = _Params(sd, ops.device)
        self.tok = P.f16("text_model.embeddings.token_embedding.weight")
        self.pos = P.f16("text_model.embeddings.position_embedding.weight")
        self.layers = []
        for i in range(cfg.num_layers):
            p = f"text_model.encoder.layers.{i}"
            wqkv = torch.cat([P.host16(f"{p}.self_attn.{n}_proj.weight") for n in "qkv"], 0).contiguous().to(ops.device)
            bqkv = torch.cat([sd[f"{p}.self_attn.{n}_proj.bias"].float() for n in "qkv"], 0).to(ops.device)
            self.layers.append(dict(
                ln1=(P.f32(f"{p}.layer_norm1.weight"), P.f32(f"{p}.layer_norm1.bias")),
                ln2=(P.f32(f"{p}.layer_norm2.weight"), P.f32(f"{p}.layer_norm2.bias")),
                wqkv=wqkv, bqkv=bqkv,
                wo=P.f16(f"{p}.self_attn.out_proj.weight"), bo=P.f32(f"{p}.self_attn.out_proj.bias"),
                w1=P.f16(f"{p}.mlp.fc1.weight"), b1=P.f32(f"{p}.mlp.fc1.bias"),
                w2=P.f16(f"{p}.mlp.fc2.weight"), b2=P.f32(f"{p}.mlp.fc2.bias")))
        self.lnf = (P.f32("text_model.final_layer_norm.weight"), P.f32("text_model.final_layer_norm.bias"))
        self.proj = P.f16("text_projection.weight") if cfg.projection_dim else None
        self.head_dim = cfg.hidden_size // cfg.num_heads

    def __call__(self, ids: torch.Tensor, penultimate: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """ids [B, T] int64 (device) -> (last_hidden_state [B, T, d] fp16, pooled projection [B, proj] fp32 or None).
        penultimate=True: the first item is `hidden_states[-2]` (input of the last layer, no final LayerNorm), the
        conditioning diffusers' SDXL encode_prompt uses."""
        ops, cfg = self.ops, self.cfg
        B, T = ids.shape
        d = cfg.hidden_size
        h = ops.embed_tokens(ids, self.tok, self.pos).reshape(B * T, d)
        act = "quick_gelu" if cfg.act == "quick_gelu" else "gelu"
        pen = None
        for li, L in enumerate(self.layers):
            if li == len(self.layers) - 1:
                pen = h                                  # (every op below allocates its output: h is never overwritten)
            n = ops.layer_norm(h, *L["ln1"], eps=cfg.eps)
            qkv = ops.linear(n, L["wqkv"], bias=L["bqkv"])
            a = ops.attention_small(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], B, cfg.num_heads, self.head_dim, T, T,
                                    self.head_dim ** -0.5, causal=True)
            h = ops.linear(a, L["wo"], bias=L["bo"], residual=h)
            n = ops.layer_norm(h, *L["ln2"], eps=cfg.eps)
            m = ops.linear(n, L["w1"], bias=L["b1"], act_pre=act)
            h = ops.linear(m, L["w2"], bias=L["b2"], residual=h)
        h = ops.layer_norm(h, *self.lnf, eps=cfg.eps).reshape(B, T, d)
        pooled = None
        if self.proj is not None:
            eot = ids.argmax(dim=-1)  # index arithmetic on token ids (plumbing): the EOT token has the largest id
            rows = h[torch.arange(B, device=ids.device), eot].contiguous()
            pooled = ops.linear(rows, self.proj, out_fp32=True)
        return (pen.reshape(B, T, d) if penultimate else h), pooled
