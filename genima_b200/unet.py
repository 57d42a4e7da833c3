"""Device-side U-Net and ControlNet (SD-2.1-base topology, SD-Turbo weights) driven through the C-ABI kernels.

Replaces, on the reference's eval path, `pipe.unet(...)` / `pipe.controlnet(...)` of diffusers 0.29.0
(`UNet2DConditionModel.forward`, `ControlNetModel.forward`), built at controller/agent/sd_controlnet_agent.py:31-42
and invoked inside the `pipe(...)` call at :67-76.  Same graph as oracle/sd_models.py, different execution:

  * activations are NHWC fp16 ([B, H, W, C] == row-major [B*H*W, C]), so 1x1 convs, linear layers and transformer
    blocks all see plain matrices and no permute/reshape kernel exists anywhere in the network;
  * ResnetBlock2D = gn_group_norm(+SiLU) -> gn_conv2d(+bias +time-embedding row) -> gn_group_norm(+SiLU) ->
    gn_conv2d(+bias +residual); when Cin != Cout the 1x1 conv_shortcut is folded into the second convolution's
    accumulation as extra K segments, and the decoder's skip concat is never materialised (two-source GroupNorm /
    two extra conv sources);
  * BasicTransformerBlock = LN -> fused QKV GEMM -> tcgen05 flash attention -> out-proj(+bias +residual) ->
    LN -> Q GEMM -> cross attention against per-prompt cached K/V -> out-proj -> LN -> GEGLU GEMM -> GEMM(+residual);
  * work that does not depend on the latents is hoisted out of the denoise loop (exactly, not approximately):
    time-embedding MLP + the 32 per-ResBlock projections per timestep, the cross-attention K/V of the 23 transformer
    layers per prompt, ControlNet's conditioning embedding per control image.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .configs import UNetConfig
from .ops import Ops, gn_bucket_for
from .packing import fold_layer_norm, pack_conv_weight, pack_geglu_weight, pack_upsample_conv_weight
from .weights import unet_skip_channels

LATENT_CPAD = 8  # latents travel as [B, h, w, 8] fp16 (4 real channels): 16-byte pixels for TMA


def tensor_key(t: torch.Tensor) -> tuple:
    """Identity of a tensor's current contents for the per-prompt / per-task caches (inference-mode tensors carry no
    version counter; they are immutable in practice)."""
    try:
        ver = t._version
    except RuntimeError:
        ver = -1
    return (t.data_ptr(), ver, tuple(t.shape), t.dtype)


class _Params:
    """Moves a state dict to the device once, in the layouts the kernels want."""

    def __init__(self, sd: Dict[str, torch.Tensor], device: torch.device):
        self.sd = sd
        self.device = device
        # channels per GroupNorm-statistics bucket of this network (ops.gn_bucket_for); 0 = statistics are not fused
        self.gn_bucket = 0
        self._f16: Dict[str, torch.Tensor] = {}
        self._f32: Dict[str, torch.Tensor] = {}

    def has(self, key: str) -> bool:
        return key in self.sd

    def f16(self, key: str) -> torch.Tensor:
        if key not in self._f16:
            self._f16[key] = self.sd[key].to(self.device, torch.float16).contiguous()
        return self._f16[key]

    def f32(self, key: str) -> torch.Tensor:
        if key not in self._f32:
            self._f32[key] = self.sd[key].to(self.device, torch.float32).contiguous()
        return self._f32[key]

    def host16(self, key: str) -> torch.Tensor:
        return self.sd[key].to(torch.float16)


class _ResBlock:
    def __init__(self, P: _Params, prefix: str, cin_parts: Sequence[int], groups: int, eps: float):
        self.groups, self.eps = groups, eps
        self.prefix = prefix
        self.bucket = P.gn_bucket
        w1 = P.host16(f"{prefix}.conv1.weight")
        self.cout = w1.shape[0]
        self.cin_parts = tuple(cin_parts)
        assert sum(cin_parts) == w1.shape[1], (prefix, cin_parts, w1.shape)
        dev = P.device
        self.g1, self.b1 = P.f32(f"{prefix}.norm1.weight"), P.f32(f"{prefix}.norm1.bias")
        self.g2, self.b2 = P.f32(f"{prefix}.norm2.weight"), P.f32(f"{prefix}.norm2.bias")
        self.w1 = pack_conv_weight(w1).to(dev)
        self.cb1 = P.f32(f"{prefix}.conv1.bias")
        w2 = P.host16(f"{prefix}.conv2.weight")
        self.has_shortcut = P.has(f"{prefix}.conv_shortcut.weight")
        if self.has_shortcut:
            wsc = P.host16(f"{prefix}.conv_shortcut.weight").reshape(self.cout, -1)
            extras, off = [], 0
            for c in cin_parts:
                extras.append(wsc[:, off:off + c])
                off += c
            self.w2 = pack_conv_weight(w2, extras=extras).to(dev)
            self.cb2 = (P.sd[f"{prefix}.conv2.bias"].float() + P.sd[f"{prefix}.conv_shortcut.bias"].float()).to(dev)
        else:
            assert len(cin_parts) == 1
            self.w2 = pack_conv_weight(w2).to(dev)
            self.cb2 = P.f32(f"{prefix}.conv2.bias")
        self.has_temb = P.has(f"{prefix}.time_emb_proj.weight")
        if self.has_temb:
            self.wt = P.f16(f"{prefix}.time_emb_proj.weight")
            self.bt = P.f32(f"{prefix}.time_emb_proj.bias")

    def __call__(self, ops: Ops, x0: torch.Tensor, x1: Optional[torch.Tensor], temb_row: Optional[torch.Tensor]):
        n1 = ops.group_norm(x0, self.g1, self.b1, self.groups, self.eps, silu=True, x1=x1)
        hw = x0.shape[1] * x0.shape[2]
        # both convolutions also accumulate the GroupNorm statistics of their output (for norm2 / the next block's norm)
        if temb_row is not None:
            h = ops.conv2d(n1, self.w1, self.cout, bias=self.cb1, rowvec=temb_row, rows_per_batch=hw,
                           gn_stats=self.bucket)
        else:
            h = ops.conv2d(n1, self.w1, self.cout, bias=self.cb1, gn_stats=self.bucket)
        n2 = ops.group_norm(h, self.g2, self.b2, self.groups, self.eps, silu=True)
        if self.has_shortcut:
            extras = [x0] if x1 is None else [x0, x1]
            return ops.conv2d(n2, self.w2, self.cout, extras=extras, bias=self.cb2, gn_stats=self.bucket)
        return ops.conv2d(n2, self.w2, self.cout, bias=self.cb2, residual=x0, gn_stats=self.bucket)


class _TransformerBlock:
    """One BasicTransformerBlock (self attention, cross attention, GEGLU feed-forward).  The three LayerNorms are folded
    into the GEMMs that consume them (gn_epilogue.ln_*): gamma goes into the weights, beta into the bias, and
    (mean, rstd) come from row statistics written by the producing GEMM's epilogue."""

    def __init__(self, P: _Params, t: str):
        dev = P.device
        ln = [(P.sd[f"{t}.norm{i}.weight"].to(dev).float(), P.sd[f"{t}.norm{i}.bias"].to(dev).float()) for i in (1, 2, 3)]
        w_qkv = torch.cat([P.host16(f"{t}.attn1.to_{n}.weight") for n in "qkv"], dim=0).to(dev)
        self.w_qkv, self.cs_qkv, self.b_qkv = fold_layer_norm(w_qkv, *ln[0])
        self.w_o1, self.b_o1 = P.f16(f"{t}.attn1.to_out.0.weight"), P.f32(f"{t}.attn1.to_out.0.bias")
        self.w_q2, self.cs_q2, self.b_q2 = fold_layer_norm(P.f16(f"{t}.attn2.to_q.weight"), *ln[1])
        self.w_kv2 = torch.cat([P.host16(f"{t}.attn2.to_k.weight"), P.host16(f"{t}.attn2.to_v.weight")], dim=0)
        self.w_o2, self.b_o2 = P.f16(f"{t}.attn2.to_out.0.weight"), P.f32(f"{t}.attn2.to_out.0.bias")
        wg, bg = pack_geglu_weight(P.host16(f"{t}.ff.net.0.proj.weight"), P.sd[f"{t}.ff.net.0.proj.bias"].float())
        self.w_ff1, self.cs_ff1, self.b_ff1 = fold_layer_norm(wg.to(dev), *ln[2], bias=bg.to(dev))
        self.w_ff2, self.b_ff2 = P.f16(f"{t}.ff.net.2.weight"), P.f32(f"{t}.ff.net.2.bias")


class _Transformer2D:
    def __init__(self, P: _Params, prefix: str, c: int, heads: int, groups: int, depth: int = 1):
        self.c, self.heads, self.groups = c, heads, groups
        self.prefix = prefix
        self.bucket = P.gn_bucket
        self.gn_g, self.gn_b = P.f32(f"{prefix}.norm.weight"), P.f32(f"{prefix}.norm.bias")
        self.w_in, self.b_in = P.f16(f"{prefix}.proj_in.weight"), P.f32(f"{prefix}.proj_in.bias")
        self.w_out, self.b_out = P.f16(f"{prefix}.proj_out.weight"), P.f32(f"{prefix}.proj_out.bias")
        self.blocks = [_TransformerBlock(P, f"{prefix}.transformer_blocks.{k}") for k in range(depth)]
        # the context K/V projections of all blocks as one GEMM: block k reads columns [2 C k, 2 C (k + 1))
        self.w_kv2 = torch.cat([b.w_kv2 for b in self.blocks], dim=0).contiguous().to(P.device)
        for b in self.blocks:
            del b.w_kv2
        self.ln_eps = 1e-5
        self.scale = 64 ** -0.5
        assert c // heads == 64, "the tcgen05 attention kernel is specialised for head_dim 64"

    def project_context(self, ops: Ops, ctx: torch.Tensor) -> torch.Tensor:
        """ctx [B, Tk, D] -> [B*Tk, depth * 2C] = (K | V) per block; constant per prompt, so computed once and cached by
        the caller."""
        return ops.linear(ctx.reshape(-1, ctx.shape[-1]), self.w_kv2)

    def __call__(self, ops: Ops, x: torch.Tensor, kv: torch.Tensor, tk: int) -> torch.Tensor:
        B, H, W, C = x.shape
        T = H * W
        n = ops.group_norm(x, self.gn_g, self.gn_b, self.groups, 1e-6, silu=False)
        st = ops.new_row_stats(B * T, C)
        h = ops.linear(n.reshape(B * T, C), self.w_in, bias=self.b_in, row_stats=st)
        for i, blk in enumerate(self.blocks):
            k2 = kv[:, 2 * C * i:2 * C * (i + 1)]
            # self attention (norm1 folded into the QKV projection)
            qkv = ops.linear(h, blk.w_qkv, bias=blk.b_qkv, ln=(st, blk.cs_qkv, self.ln_eps))
            a = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, self.heads, T, T, self.scale)
            st = ops.new_row_stats(B * T, C)
            h = ops.linear(a, blk.w_o1, bias=blk.b_o1, residual=h, row_stats=st)
            # cross attention against the cached text K/V (norm2 folded into the Q projection)
            if ops.fuse_qproj:   # ... inside the attention kernel: one launch, no trip of Q through memory
                a = ops.attention_qproj(h, blk.w_q2, k2[:, :C], k2[:, C:], B, self.heads, T, tk, self.scale,
                                        bias=blk.b_q2, ln=(st, blk.cs_q2, self.ln_eps))
            else:
                q = ops.linear(h, blk.w_q2, bias=blk.b_q2, ln=(st, blk.cs_q2, self.ln_eps))
                a = ops.attention(q, k2[:, :C], k2[:, C:], B, self.heads, T, tk, self.scale)
            st = ops.new_row_stats(B * T, C)
            h = ops.linear(a, blk.w_o2, bias=blk.b_o2, residual=h, row_stats=st)
            # GEGLU feed-forward (norm3 folded into the first projection)
            g = ops.linear(h, blk.w_ff1, bias=blk.b_ff1, ln=(st, blk.cs_ff1, self.ln_eps), geglu=True)
            if i + 1 < len(self.blocks):     # the next block's norm1 needs the row statistics of this output
                st = ops.new_row_stats(B * T, C)
                h = ops.linear(g, blk.w_ff2, bias=blk.b_ff2, residual=h, row_stats=st)
            else:
                h = ops.linear(g, blk.w_ff2, bias=blk.b_ff2, residual=h)
        out = ops.linear(h, self.w_out, bias=self.b_out, residual=x.reshape(B * T, C), rows_per_batch=T,
                         gn_stats=self.bucket)
        return ops.carry_stats(out.reshape(B, H, W, C), out)


class _Conv:
    def __init__(self, P: _Params, prefix: str, stride: int = 1, cin_layout: Sequence[int] = (), gn: bool = False,
                 upsample: bool = False):
        self.bucket = P.gn_bucket if gn else 0   # gn: the output feeds a GroupNorm -> accumulate its statistics
        w = P.host16(f"{prefix}.weight")
        self.cout, self.k, self.stride = w.shape[0], w.shape[2], stride
        self.pad = self.k // 2
        self.w = pack_conv_weight(w, cin_layout=cin_layout).to(P.device)
        self.b = P.f32(f"{prefix}.bias") if P.has(f"{prefix}.bias") else torch.zeros(self.cout, device=P.device)
        # upsample: this convolution follows a nearest x2 upsample (Upsample2D): keep the four 2x2 phase kernels too
        self.w4 = pack_upsample_conv_weight(w).to(P.device) if upsample and self.k == 3 and w.shape[1] % 8 == 0 else None

    def __call__(self, ops: Ops, x: torch.Tensor, **epi) -> torch.Tensor:
        return ops.conv2d(x, self.w, self.cout, ksize=self.k, stride=self.stride, pad=self.pad, bias=self.b,
                          gn_stats=self.bucket, **epi)

    def upsampled(self, ops: Ops, x: torch.Tensor, **epi) -> torch.Tensor:
        """conv(nearest_upsample_x2(x)): folded into four 2x2 phase convolutions over x (4/9 of the multiply-adds, no
        upsampled tensor) when the phase kernels exist, else upsample kernel + convolution."""
        if self.w4 is not None and ops.fold_upsample and x.shape[2] % 8 == 0:
            return ops.conv2d_up2x(x, self.w4, self.cout, bias=self.b, gn_stats=self.bucket, **epi)
        return self(ops, ops.upsample_nearest2x(x), **epi)


class _Encoder:
    """conv_in + time embedding + down blocks + mid block: shared by the U-Net and the ControlNet."""

    def __init__(self, ops: Ops, sd: Dict[str, torch.Tensor], cfg: UNetConfig):
        self.ops, self.cfg = ops, cfg
        P = self.P = _Params(sd, ops.device)
        g, eps = cfg.norm_num_groups, cfg.norm_eps
        ch = cfg.block_out_channels
        P.gn_bucket = gn_bucket_for(ch, g)
        self.conv_in = _Conv(P, "conv_in", cin_layout=(cfg.in_channels, LATENT_CPAD), gn=True)
        self.te_w1, self.te_b1 = P.f16("time_embedding.linear_1.weight"), P.f32("time_embedding.linear_1.bias")
        self.te_w2, self.te_b2 = P.f16("time_embedding.linear_2.weight"), P.f32("time_embedding.linear_2.bias")
        if cfg.addition_embed:   # SDXL text_time conditioning
            self.ae_w1, self.ae_b1 = P.f16("add_embedding.linear_1.weight"), P.f32("add_embedding.linear_1.bias")
            self.ae_w2, self.ae_b2 = P.f16("add_embedding.linear_2.weight"), P.f32("add_embedding.linear_2.bias")
        self.down: List[Tuple[List[_ResBlock], List[Optional[_Transformer2D]], Optional[_Conv]]] = []
        cin = ch[0]
        for i, cout in enumerate(ch):
            res, att = [], []
            for j in range(cfg.layers_per_block):
                res.append(_ResBlock(P, f"down_blocks.{i}.resnets.{j}", (cin,), g, eps))
                att.append(_Transformer2D(P, f"down_blocks.{i}.attentions.{j}", cout, cfg.num_heads[i], g,
                                          cfg.tf_layers(i)) if cfg.attn_levels[i] else None)
                cin = cout
            ds = _Conv(P, f"down_blocks.{i}.downsamplers.0.conv", stride=2, gn=True) if i < len(ch) - 1 else None
            self.down.append((res, att, ds))
        self.mid_res0 = _ResBlock(P, "mid_block.resnets.0", (ch[-1],), g, eps)
        self.mid_attn = _Transformer2D(P, "mid_block.attentions.0", ch[-1], cfg.num_heads[-1], g,
                                       cfg.tf_layers(len(ch) - 1))
        self.mid_res1 = _ResBlock(P, "mid_block.resnets.1", (ch[-1],), g, eps)

    # ---- hoisted, latent-independent work --------------------------------------------------------------------
    def resblocks(self) -> List[_ResBlock]:
        out = []
        for res, _, _ in self.down:
            out += res
        return out + [self.mid_res0, self.mid_res1]

    def transformers(self) -> List[_Transformer2D]:
        out = []
        for _, att, _ in self.down:
            out += [a for a in att if a is not None]
        return out + [self.mid_attn]

    def time_embedding(self, t: float, added: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        """silu(emb) [1, 1280] fp16, emb = Linear(SiLU(Linear(sinusoid(t)))) [+ add_embedding(cat(text_embeds,
        sinusoid(time_ids))) for SDXL; added = dict(text_embeds [1, P] , time_ids: 6 floats)]."""
        ops = self.ops
        e = ops.timestep_embedding(t, self.cfg.block_out_channels[0])
        e = ops.linear(e, self.te_w1, bias=self.te_b1, act_pre="silu")
        # every consumer applies SiLU to temb first (ResnetBlock2D.nonlinearity), so fuse it here
        if not self.cfg.addition_embed:
            return ops.linear(e, self.te_w2, bias=self.te_b2, act_pre="silu")
        if added is None:
            raise ValueError("this model needs added conditioning (text_embeds, time_ids)")
        temb = ops.linear(e, self.te_w2, bias=self.te_b2)
        d = self.cfg.addition_time_embed_dim
        parts = [added["text_embeds"].reshape(1, -1).to(self.ops.device, torch.float16)]
        parts += [ops.timestep_embedding(float(v), d) for v in added["time_ids"]]
        a = torch.cat(parts, dim=1).contiguous()                       # (concatenation: data movement only)
        if a.shape[1] != self.cfg.projection_input_dim:
            raise ValueError(f"added conditioning has {a.shape[1]} features, the model expects "
                             f"{self.cfg.projection_input_dim}")
        a = ops.linear(a, self.ae_w1, bias=self.ae_b1, act_pre="silu")
        return ops.linear(a, self.ae_w2, bias=self.ae_b2, residual=temb, act_post="silu")   # silu(temb + aug_emb)

    def temb_rows(self, blocks: Sequence[_ResBlock], silu_temb: torch.Tensor, batch: int) -> Dict[str, torch.Tensor]:
        rows = {}
        for rb in blocks:
            if rb.has_temb:
                r = self.ops.linear(silu_temb, rb.wt, bias=rb.bt, out_fp32=True)
                rows[rb.prefix] = r.expand(batch, -1).contiguous()
        return rows

    def run(self, h: torch.Tensor, temb: Dict[str, torch.Tensor], kv: Dict[str, torch.Tensor], tk: int, on_skip=None):
        """h: conv_in output [B, H, W, C0] -> (mid output, [S0..S11]).  on_skip(i) is called right after skip i (and
        finally i = len(skips) for the mid-block output) has been enqueued, so that a caller can record a stream event
        per tensor and start consuming them while the rest of the encoder is still running."""
        ops = self.ops
        skips = [h]
        note = on_skip if on_skip is not None else (lambda i: None)
        note(0)
        for res, att, ds in self.down:
            for rb, tr in zip(res, att):
                h = rb(ops, h, None, temb[rb.prefix])
                if tr is not None:
                    h = tr(ops, h, kv[tr.prefix], tk)
                skips.append(h)
                note(len(skips) - 1)
            if ds is not None:
                h = ds(ops, h)
                skips.append(h)
                note(len(skips) - 1)
        h = self.mid_res0(ops, h, None, temb[self.mid_res0.prefix])
        h = self.mid_attn(ops, h, kv[self.mid_attn.prefix], tk)
        h = self.mid_res1(ops, h, None, temb[self.mid_res1.prefix])
        note(len(skips))
        return h, skips


class DeviceUNet(_Encoder):
    """UNet2DConditionModel.forward(sample, t, encoder_hidden_states, down_block_additional_residuals, mid_...)."""

    def __init__(self, ops: Ops, sd: Dict[str, torch.Tensor], cfg: UNetConfig):
        super().__init__(ops, sd, cfg)
        P = self.P
        g, eps = cfg.norm_num_groups, cfg.norm_eps
        ch = cfg.block_out_channels
        skips = unet_skip_channels(cfg)
        self.up: List[Tuple[List[_ResBlock], List[Optional[_Transformer2D]], Optional[_Conv]]] = []
        prev = ch[-1]
        for i, cout in enumerate(reversed(ch)):
            level = len(ch) - 1 - i
            res, att = [], []
            for j in range(cfg.layers_per_block + 1):
                res.append(_ResBlock(P, f"up_blocks.{i}.resnets.{j}", (prev, skips.pop()), g, eps))
                att.append(_Transformer2D(P, f"up_blocks.{i}.attentions.{j}", cout, cfg.num_heads[level], g,
                                          cfg.tf_layers(level)) if cfg.attn_levels[level] else None)
                prev = cout
            us = _Conv(P, f"up_blocks.{i}.upsamplers.0.conv", gn=True, upsample=True) if i < len(ch) - 1 else None
            self.up.append((res, att, us))
        self.out_g, self.out_b = P.f32("conv_norm_out.weight"), P.f32("conv_norm_out.bias")
        self.conv_out = _Conv(P, "conv_out")
        # conv_out with its rows zero-padded to the 8-channel latent pixel: the scheduler step is fused into its epilogue
        # (x' = a * x + b * (conv + bias) written straight into the next latent tensor, padding channels stay 0)
        co = self.conv_out
        if co.cout <= LATENT_CPAD:
            self.w_out8 = torch.zeros(LATENT_CPAD, co.w.shape[1], dtype=torch.float16, device=P.device)
            self.w_out8[:co.cout] = co.w
            self.b_out8 = torch.zeros(LATENT_CPAD, dtype=torch.float32, device=P.device)
            self.b_out8[:co.cout] = co.b
        else:
            self.w_out8 = None

    def resblocks(self) -> List[_ResBlock]:
        out = super().resblocks()
        for res, _, _ in self.up:
            out += res
        return out

    def transformers(self) -> List[_Transformer2D]:
        out = super().transformers()
        for _, att, _ in self.up:
            out += [a for a in att if a is not None]
        return out

    def encode(self, x: torch.Tensor, temb, kv, tk, on_skip=None, in_scale: Optional[torch.Tensor] = None):
        """in_scale: fp32 [C0] vector s (all entries equal): the scheduler's scale_model_input folded into conv_in's
        epilogue, conv_in(s * x) + b == s * conv(x) + b, so `x` is the UNSCALED sample."""
        h = self.conv_in(self.ops, x, scale=in_scale) if in_scale is not None else self.conv_in(self.ops, x)
        return self.run(h, temb, kv, tk, on_skip)

    def decode(self, h: torch.Tensor, skips: List[torch.Tensor], temb, kv, tk, eps_out: torch.Tensor,
               step: Optional[Tuple[torch.Tensor, float, float]] = None) -> torch.Tensor:
        """`skips` / `h` already include the ControlNet residuals.  Writes eps into eps_out [B, H, W, 8] (4 valid).
        step = (x, a, b): the scheduler update is fused into conv_out's epilogue instead -- eps_out receives
        x' = a * x + b * eps (Euler: a = 1, b = sigma_next - sigma; DDIM: its (a, b)), eps itself is never stored."""
        ops = self.ops
        skips = list(skips)
        for res, att, us in self.up:
            for rb, tr in zip(res, att):
                h = rb(ops, h, skips.pop(), temb[rb.prefix])
                if tr is not None:
                    h = tr(ops, h, kv[tr.prefix], tk)
            if us is not None:
                h = us.upsampled(ops, h)
        n = ops.group_norm(h, self.out_g, self.out_b, self.cfg.norm_num_groups, self.cfg.norm_eps, silu=True)
        if step is not None:
            x, a, b = step
            if self.w_out8 is None:
                raise ValueError("fused scheduler step needs out_channels <= 8")
            co = self.conv_out
            return ops.conv2d(n, self.w_out8, LATENT_CPAD, ksize=co.k, stride=1, pad=co.pad, bias=self.b_out8,
                              residual=x, alpha=float(b), beta=float(a), out=eps_out)
        return self.conv_out(ops, n, out=eps_out)


class DeviceControlNet(_Encoder):
    """ControlNetModel.forward(guess_mode=False, conditioning_scale=1.0)."""

    def __init__(self, ops: Ops, sd: Dict[str, torch.Tensor], cfg: UNetConfig):
        super().__init__(ops, sd, cfg)
        P = self.P
        ce = cfg.cond_embed_channels
        p = "controlnet_cond_embedding"
        self.ce_convs: List[_Conv] = [_Conv(P, f"{p}.conv_in", cin_layout=(3, 64))]
        for k in range(2 * (len(ce) - 1)):
            self.ce_convs.append(_Conv(P, f"{p}.blocks.{k}", stride=2 if k % 2 == 1 else 1))
        self.ce_out = _Conv(P, f"{p}.conv_out")
        self.zero_w = [P.f16(f"controlnet_down_blocks.{i}.weight").reshape(c, c).contiguous()
                       for i, c in enumerate(unet_skip_channels(cfg))]
        self.zero_b = [P.f32(f"controlnet_down_blocks.{i}.bias") for i in range(len(self.zero_w))]
        cm = cfg.block_out_channels[-1]
        self.mid_w = P.f16("controlnet_mid_block.weight").reshape(cm, cm).contiguous()
        self.mid_b = P.f32("controlnet_mid_block.bias")

    def cond_embedding(self, cond_nhwc: torch.Tensor) -> torch.Tensor:
        """cond [B, 8h, 8w, 64] fp16 (3 real channels, values in [0, 1]) -> [B, h, w, C0].  Constant per control image."""
        ops = self.ops
        e = cond_nhwc
        for conv in self.ce_convs:
            e = conv(ops, e, act_pre="silu")
        return self.ce_out(ops, e)

    def encode(self, x: torch.Tensor, cond_emb: torch.Tensor, temb, kv, tk, on_skip=None,
               in_scale: Optional[torch.Tensor] = None):
        """conv_in(x) + cond_emb -> encoder copy -> (mid, [S0..S11]) BEFORE the zero-convs.  Independent of the U-Net's own
        encoder, so the pipeline runs it on a second stream concurrently with DeviceUNet.encode.  in_scale: see
        DeviceUNet.encode."""
        epi = dict(scale=in_scale) if in_scale is not None else {}
        h = self.conv_in(self.ops, x, residual=cond_emb, **epi)
        return self.run(h, temb, kv, tk, on_skip)

    def zero_conv(self, i: int, s: torch.Tensor, us: torch.Tensor, conditioning_scale: float = 1.0,
                  ops: Optional[Ops] = None) -> torch.Tensor:
        """Zero-conv of ControlNet skip i (i == len(zero_w): the mid-block one) added to the U-Net tensor `us`."""
        ops = ops or self.ops
        w, b = (self.zero_w[i], self.zero_b[i]) if i < len(self.zero_w) else (self.mid_w, self.mid_b)
        B, H, W, C = s.shape
        o = ops.linear(s.reshape(B * H * W, C), w, bias=b, residual=us.reshape(B * H * W, C), alpha=conditioning_scale,
                       rows_per_batch=H * W, gn_stats=self.P.gn_bucket)
        return ops.carry_stats(o.reshape(B, H, W, C), o)

    def zero_convs(self, mid: torch.Tensor, skips: List[torch.Tensor], unet_skips: List[torch.Tensor],
                   unet_mid: torch.Tensor, conditioning_scale: float = 1.0, ops: Optional[Ops] = None):
        """Returns (U-Net skips + down residuals, U-Net mid + mid residual): the zero-conv epilogues add the U-Net tensors,
        so the `sample + residual` adds of UNet2DConditionModel.forward cost no extra pass."""
        out = [self.zero_conv(i, s, us, conditioning_scale, ops) for i, (s, us) in enumerate(zip(skips, unet_skips))]
        return out, self.zero_conv(len(self.zero_w), mid, unet_mid, conditioning_scale, ops)

    def residuals(self, x: torch.Tensor, cond_emb: torch.Tensor, temb, kv, tk, unet_skips: List[torch.Tensor],
                  unet_mid: torch.Tensor, conditioning_scale: float = 1.0):
        """encode + zero_convs on the current stream."""
        mid, skips = self.encode(x, cond_emb, temb, kv, tk)
        return self.zero_convs(mid, skips, unet_skips, unet_mid, conditioning_scale)
