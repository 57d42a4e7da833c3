"""Device-side KL-VAE decoder (diffusers AutoencoderKL.decode: post_quant_conv + Decoder), the encoder
(AutoencoderKL.encode, for the InstructPix2Pix sibling) and the TAESD decoder on the C-ABI kernels.

Replaces `pipe.vae.decode(latents / scaling_factor)` at the end of diffusers' StableDiffusionControlNetPipeline.__call__
(reached from controller/agent/sd_controlnet_agent.py:67-76).  Same NHWC / fused-epilogue execution as unet.py; the
mid-block attention (1 head, d = 512, N = h*w) runs as two tcgen05 GEMMs around a row softmax:
S = Q K^T (fp32), P = softmax(S / sqrt(C)) (fp16), O = P V with V^T produced directly by the value projection.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .configs import TAESDConfig, VAEConfig
from .ops import Ops, gn_bucket_for
from .packing import pack_conv_weight
from .unet import LATENT_CPAD, _Conv, _Params, _ResBlock
from .weights import taesd_layer_plan

RGB_CPAD = 8  # decoded image travels as [B, H, W, 8] fp16 (3 real channels)


class _MidAttention:
    """The VAE mid-block attention (1 head over all channels, projections with bias, residual) as tcgen05 GEMMs around a
    row softmax; shared by the decoder and the encoder."""

    def __init__(self, P: _Params, prefix: str, cfg: VAEConfig):
        self.P, self.cfg = P, cfg
        a = prefix
        self.at_g, self.at_b = P.f32(f"{a}.group_norm.weight"), P.f32(f"{a}.group_norm.bias")
        self.wq, self.bq = P.f16(f"{a}.to_q.weight"), P.f32(f"{a}.to_q.bias")
        self.wk, self.bk = P.f16(f"{a}.to_k.weight"), P.f32(f"{a}.to_k.bias")
        self.wv, self.bv = P.f16(f"{a}.to_v.weight"), P.f32(f"{a}.to_v.bias")
        self.wo, self.bo = P.f16(f"{a}.to_out.0.weight"), P.f32(f"{a}.to_out.0.bias")

    def __call__(self, ops: Ops, h: torch.Tensor) -> torch.Tensor:
        B, H, W, C = h.shape
        T = H * W
        n = ops.group_norm(h, self.at_g, self.at_b, self.cfg.norm_num_groups, self.cfg.norm_eps, silu=False)
        outs = []
        for b in range(B):  # one [T, T] score matrix at a time
            nb = n[b].reshape(T, C)
            q = ops.linear(nb, self.wq, bias=self.bq)
            k = ops.linear(nb, self.wk, bias=self.bk)
            # V^T [C, T] = Wv @ n^T: the value projection with operands swapped, so P V needs no transpose kernel.
            # Its bias is added after P V instead (softmax rows sum to 1, so P (V + 1 b^T) = P V + b^T).
            # (w_dynamic: the "weight" operand of these three GEMMs is an activation produced just before)
            vt = ops.linear(self.wv, nb, w_dynamic=True)
            s = ops.linear(q, k, out_fp32=True, w_dynamic=True)   # fp32 scores: |q.k| over 512 dims is too coarse in fp16
            pr = ops.softmax_rows(s, scale=C ** -0.5)
            o = ops.linear(pr, vt, bias=self.bv, w_dynamic=True)
            outs.append(ops.linear(o, self.wo, bias=self.bo, residual=h[b].reshape(T, C),
                                   gn_stats=self.P.gn_bucket if B == 1 else 0))
        out = outs[0] if B == 1 else torch.cat(outs, dim=0)
        return ops.carry_stats(out.reshape(B, H, W, C), out)


class DeviceVAEDecoder:
    def __init__(self, ops: Ops, sd: Dict[str, torch.Tensor], cfg: VAEConfig):
        self.ops, self.cfg = ops, cfg
        P = self.P = _Params(sd, ops.device)
        g, eps = cfg.norm_num_groups, cfg.norm_eps
        ch = cfg.block_out_channels
        top = ch[-1]
        lc = cfg.latent_channels
        P.gn_bucket = gn_bucket_for(ch, g)
        # post_quant_conv (1x1, 4 -> 4) as a linear over 8-channel padded pixels
        wq = torch.zeros(lc, LATENT_CPAD, dtype=torch.float16)
        wq[:, :lc] = P.host16("post_quant_conv.weight").reshape(lc, lc)
        self.pq_w, self.pq_b = wq.to(ops.device), P.f32("post_quant_conv.bias")
        self.conv_in = _Conv(P, "decoder.conv_in", cin_layout=(lc, LATENT_CPAD), gn=True)
        self.mid0 = _ResBlock(P, "decoder.mid_block.resnets.0", (top,), g, eps)
        self.mid1 = _ResBlock(P, "decoder.mid_block.resnets.1", (top,), g, eps)
        self.mid_attn = _MidAttention(P, "decoder.mid_block.attentions.0", cfg)
        self.up: List = []
        prev = top
        for i, cout in enumerate(reversed(ch)):
            res = []
            for j in range(cfg.layers_per_block + 1):
                res.append(_ResBlock(P, f"decoder.up_blocks.{i}.resnets.{j}", (prev,), g, eps))
                prev = cout
            us = _Conv(P, f"decoder.up_blocks.{i}.upsamplers.0.conv", gn=True, upsample=True) if i < len(ch) - 1 else None
            self.up.append((res, us))
        self.out_g, self.out_b = P.f32("decoder.conv_norm_out.weight"), P.f32("decoder.conv_norm_out.bias")
        self.conv_out = _Conv(P, "decoder.conv_out")

    def _mid_attention(self, h: torch.Tensor) -> torch.Tensor:
        return self.mid_attn(self.ops, h)

    def decode(self, z: torch.Tensor, out: Optional[torch.Tensor] = None, in_scale: float = 1.0) -> torch.Tensor:
        """z: [B, h, w, 8] fp16 latents -> [B, 8h, 8w, 8] fp16 (RGB in 0..2).  in_scale: `latents / scaling_factor` of
        the pipeline folded into post_quant_conv's epilogue (W (s z) + b == s (W z) + b); 1.0 = z is already divided."""
        ops = self.ops
        B, hh, ww, _ = z.shape
        zq = torch.zeros_like(z)
        sc = None
        if in_scale != 1.0:
            if getattr(self, "_pq_scale_val", None) != in_scale:
                self._pq_scale = torch.full((self.pq_w.shape[0],), float(in_scale), dtype=torch.float32, device=z.device)
                self._pq_scale_val = in_scale
            sc = self._pq_scale
        ops.linear(z.reshape(-1, LATENT_CPAD), self.pq_w, bias=self.pq_b, scale=sc, out=zq.reshape(-1, LATENT_CPAD))
        h = self.conv_in(ops, zq)
        h = self.mid0(ops, h, None, None)
        h = self._mid_attention(h)
        h = self.mid1(ops, h, None, None)
        for res, us in self.up:
            for rb in res:
                h = rb(ops, h, None, None)
            if us is not None:
                h = us.upsampled(ops, h)
        n = ops.group_norm(h, self.out_g, self.out_b, self.cfg.norm_num_groups, self.cfg.norm_eps, silu=True)
        if out is None:
            out = torch.zeros(B, hh * 8, ww * 8, RGB_CPAD, dtype=torch.float16, device=z.device)
        return self.conv_out(ops, n, out=out)


class DeviceVAEEncoder:
    """diffusers AutoencoderKL.encode(x).latent_dist.mode() (Encoder + quant_conv, mean half of the moments) on the same
    kernels: what StableDiffusionInstructPix2PixPipeline.prepare_image_latents computes once per call
    (controller/agent/sd_pix2pix_agent.py:52-60).  Downsample2D's F.pad(0, 1, 0, 1) + stride-2 convolution is one
    gn_conv2d_asym launch (the padding is TMA zero fill); only the 4 mean rows of conv_out∘quant_conv are evaluated."""

    def __init__(self, ops: Ops, sd: Dict[str, torch.Tensor], cfg: VAEConfig):
        self.ops, self.cfg = ops, cfg
        P = self.P = _Params(sd, ops.device)
        g, eps = cfg.norm_num_groups, cfg.norm_eps
        ch = cfg.block_out_channels
        top = ch[-1]
        lc = cfg.latent_channels
        P.gn_bucket = gn_bucket_for(ch, g)
        self.conv_in = _Conv(P, "encoder.conv_in", cin_layout=(cfg.out_channels, RGB_CPAD), gn=True)
        self.down: List = []
        prev = ch[0]
        for i, cout in enumerate(ch):
            res = []
            for j in range(cfg.layers_per_block):
                res.append(_ResBlock(P, f"encoder.down_blocks.{i}.resnets.{j}", (prev,), g, eps))
                prev = cout
            ds = _Conv(P, f"encoder.down_blocks.{i}.downsamplers.0.conv", stride=2, gn=True) if i < len(ch) - 1 else None
            self.down.append((res, ds))
        self.mid0 = _ResBlock(P, "encoder.mid_block.resnets.0", (top,), g, eps)
        self.mid_attn = _MidAttention(P, "encoder.mid_block.attentions.0", cfg)
        self.mid1 = _ResBlock(P, "encoder.mid_block.resnets.1", (top,), g, eps)
        self.out_g, self.out_b = P.f32("encoder.conv_norm_out.weight"), P.f32("encoder.conv_norm_out.bias")
        self.conv_out = _Conv(P, "encoder.conv_out")
        # quant_conv (1x1, 8 -> 8): only the mean rows, written into an 8-channel padded latent pixel
        wq = torch.zeros(lc, 2 * lc, dtype=torch.float16)
        wq[:, :] = P.host16("quant_conv.weight").reshape(2 * lc, 2 * lc)[:lc]
        self.q_w, self.q_b = wq.to(ops.device), P.f32("quant_conv.bias")[:lc].contiguous()

    def encode(self, image_u8: torch.Tensor) -> torch.Tensor:
        """image_u8: [B, H, W, 3] uint8 on the device -> [B, H/8, W/8, 8] fp16, channels 0..3 = posterior mean (not
        multiplied by scaling_factor), channels 4..7 zero.  The [-1, 1] normalisation of VaeImageProcessor.preprocess
        happens inside the layout kernel."""
        ops = self.ops
        x = ops.u8_to_nhwc(image_u8, cpad=RGB_CPAD, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5))
        h = self.conv_in(ops, x)
        for res, ds in self.down:
            for rb in res:
                h = rb(ops, h, None, None)
            if ds is not None:
                h = ops.conv2d_asym(h, ds.w, ds.cout, ksize=3, stride=2, pads=(0, 0, 1, 1), bias=ds.b,
                                    gn_stats=ds.bucket)
        h = self.mid0(ops, h, None, None)
        h = self.mid_attn(ops, h)
        h = self.mid1(ops, h, None, None)
        n = ops.group_norm(h, self.out_g, self.out_b, self.cfg.norm_num_groups, self.cfg.norm_eps, silu=True)
        m = self.conv_out(ops, n)                                              # [B, h, w, 8] moments before quant_conv
        B, hh, ww, c2 = m.shape
        z = torch.zeros(B, hh, ww, LATENT_CPAD, dtype=torch.float16, device=m.device)
        ops.linear(m.reshape(-1, c2), self.q_w, bias=self.q_b, out=z.reshape(-1, LATENT_CPAD))
        return z


class DeviceTAESDDecoder:
    """diffusers AutoencoderTiny.decode (DecoderTiny) on the same implicit-GEMM convolution kernel: every ReLU, the block
    residual add + ReLU and the final `x * 2 - 1` live in GEMM epilogues (the last one folded into conv_out's weights and
    bias), so the decoder is 35 convolutions, 3 nearest-upsample copies and one tanh clamp.  Same interface as
    DeviceVAEDecoder.decode; selected when eval_cfg.autoencoder names a TAESD checkpoint
    (controller/agent/sd_controlnet_agent.py:45-49)."""

    def __init__(self, ops: Ops, sd: Dict[str, torch.Tensor], cfg: TAESDConfig):
        self.ops, self.cfg = ops, cfg
        P = self.P = _Params(sd, ops.device)
        self.plan = []
        for kind, i in taesd_layer_plan(cfg):
            p = f"decoder.layers.{i}"
            if kind == "conv_in":
                self.plan.append(("conv", _Conv(P, p, cin_layout=(cfg.latent_channels, LATENT_CPAD)), "relu"))
            elif kind == "relu":
                continue                                            # fused into conv_in's epilogue
            elif kind == "block":
                self.plan.append(("block", [_Conv(P, f"{p}.conv.{j}") for j in (0, 2, 4)], None))
            elif kind == "up":
                self.plan.append(("up", None, None))
            elif kind == "conv":                                    # bias-free 64 -> 64 convolution after an upsample
                assert self.plan[-1][0] == "up"
                self.plan[-1] = ("upconv", _Conv(P, p, upsample=True), None)   # folded: four 2x2 phase convolutions
            elif kind == "conv_out":                                # (x * 2 - 1) folded: W' = 2 W, b' = 2 b - 1
                c = _Conv.__new__(_Conv)
                w = P.host16(f"{p}.weight")
                c.cout, c.k, c.stride, c.pad, c.bucket = w.shape[0], 3, 1, 1, 0
                c.w = pack_conv_weight((w.float() * 2.0).to(torch.float16)).to(ops.device)
                c.b = (P.sd[f"{p}.bias"].float() * 2.0 - 1.0).to(ops.device)
                self.plan.append(("conv_out", c, None))

    def decode(self, z: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """z: [B, h, w, 8] fp16 latents (already divided by scaling_factor = 1) -> [B, 8h, 8w, 8] fp16 (RGB in 0..2)."""
        ops = self.ops
        B, hh, ww, _ = z.shape
        up = 2 ** (len(self.cfg.num_blocks) - 1)
        h = ops.tanh_clamp(z, self.cfg.latent_magnitude)
        for kind, mod, act in self.plan:
            if kind == "conv":
                h = mod(ops, h, act_pre=act) if act else mod(ops, h)
            elif kind == "block":
                y = mod[0](ops, h, act_pre="relu")
                y = mod[1](ops, y, act_pre="relu")
                h = mod[2](ops, y, residual=h, act_post="relu")
            elif kind == "upconv":
                h = mod.upsampled(ops, h)
            elif kind == "up":
                h = ops.upsample_nearest2x(h)
            else:
                if out is None:
                    out = torch.zeros(B, hh * up, ww * up, RGB_CPAD, dtype=torch.float16, device=z.device)
                return mod(ops, h, out=out)
        raise RuntimeError("TAESD layer plan has no conv_out")
