"""GPU debugging aid: runs every block of the (tiny or full) ControlNet / U-Net in isolation, feeding each device block
the oracle's input for that block, and prints the normalised error per block.  Usage: python tests/debug_blocks.py [full]  (lives under tests/: it executes the oracle)"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from genima_b200 import weights as W  # noqa: E402
from genima_b200.configs import UNetConfig  # noqa: E402
from genima_b200.ops import Ops  # noqa: E402
from genima_b200.unet import DeviceControlNet, DeviceUNet  # noqa: E402
from oracle import sd_models as O  # noqa: E402


def nhwc16(t):
    return t.permute(0, 2, 3, 1).contiguous().to("cuda", torch.float16)


def err(name, out, ref_nchw):
    o = out.float().cpu()
    r = ref_nchw.permute(0, 2, 3, 1).float() if ref_nchw.dim() == 4 else ref_nchw.float()
    o = o[..., : r.shape[-1]]
    fin = bool(torch.isfinite(o).all())
    e = float((o - r).abs().max() / r.abs().max().clamp_min(1e-12)) if fin else float("nan")
    print(f"{'ok  ' if fin and e < 4e-3 else 'BAD '} {name}: {e:.3e} shape {tuple(o.shape)}", flush=True)


def main():
    full = len(sys.argv) > 1 and sys.argv[1] == "full"
    cfg = UNetConfig() if full else UNetConfig.tiny()
    ops = Ops(0)
    usd = W.synth_state_dict(W.unet_shapes(cfg))
    csd = W.synth_state_dict(W.controlnet_shapes(cfg), salt=1)
    g = torch.Generator().manual_seed(0)
    s = cfg.sample_size
    B = 1
    x = torch.randn(B, 4, s, s, generator=g).half().float()
    ctx = torch.randn(B, 77, cfg.cross_attention_dim, generator=g).half().float()
    cond_u8 = torch.randint(0, 256, (B, 8 * s, 8 * s, 3), generator=g, dtype=torch.uint8)
    cond = (cond_u8.float() / 255).permute(0, 3, 1, 2)
    t = 599.0
    unet = DeviceUNet(ops, usd, cfg)
    cn = DeviceControlNet(ops, csd, cfg)
    ctx16 = ctx.to("cuda", torch.float16)

    # ---- cond embedding, conv by conv
    p = "controlnet_cond_embedding"
    e_ref = cond
    e_dev = ops.u8_to_nhwc(cond_u8.cuda(), cpad=64)
    err("u8_to_nhwc", e_dev, cond)
    names = ["conv_in"] + [f"blocks.{k}" for k in range(2 * (len(cfg.cond_embed_channels) - 1))]
    for name, conv in zip(names, cn.ce_convs):
        e_ref = F.silu(F.conv2d(e_ref, csd[f"{p}.{name}.weight"].float(), csd[f"{p}.{name}.bias"].float(),
                                stride=conv.stride, padding=1))
        out = conv(ops, e_dev, act_pre="silu")
        err(f"cond_embedding.{name} (stride {conv.stride}, C {e_dev.shape[-1]} -> {conv.cout})", out, e_ref)
        e_dev = nhwc16(e_ref)
    # ---- time embedding
    temb_ref = O.time_embed(usd, cfg, torch.tensor([t]))
    st = unet.time_embedding(t)
    err("silu(time embedding)", st, F.silu(temb_ref))
    rows = unet.temb_rows(unet.resblocks(), st, B)
    rb0 = unet.resblocks()[0]
    ref_row = F.linear(F.silu(temb_ref), usd[f"{rb0.prefix}.time_emb_proj.weight"].float(),
                       usd[f"{rb0.prefix}.time_emb_proj.bias"].float())
    err("temb row of first resblock", rows[rb0.prefix], ref_row)
    # ---- conv_in
    xs = torch.zeros(B, s, s, 8, dtype=torch.float16)
    xs[..., :4] = x.permute(0, 2, 3, 1).half()
    xs = xs.cuda()
    h_ref = F.conv2d(x, usd["conv_in.weight"].float(), usd["conv_in.bias"].float(), padding=1)
    err("conv_in", unet.conv_in(ops, xs), h_ref)
    # ---- encoder blocks in isolation
    gN, eps = cfg.norm_num_groups, cfg.norm_eps
    kv = {tr.prefix: tr.project_context(ops, ctx16) for tr in unet.transformers()}
    skips_ref = [h_ref]
    for (res, att, ds), i in zip(unet.down, range(len(unet.down))):
        for rb, tr in zip(res, att):
            r = O.resnet_block(usd, rb.prefix, h_ref, temb_ref, gN, eps)
            err(rb.prefix, rb(ops, nhwc16(h_ref), None, rows[rb.prefix]), r)
            h_ref = r
            if tr is not None:
                r = O.transformer2d(usd, tr.prefix, h_ref, ctx, tr.heads, gN)
                err(tr.prefix, tr(ops, nhwc16(h_ref), kv[tr.prefix], 77), r)
                h_ref = r
            skips_ref.append(h_ref)
        if ds is not None:
            pfx = f"down_blocks.{i}.downsamplers.0.conv"
            r = F.conv2d(h_ref, usd[f"{pfx}.weight"].float(), usd[f"{pfx}.bias"].float(), stride=2, padding=1)
            err(pfx, ds(ops, nhwc16(h_ref)), r)
            h_ref = r
            skips_ref.append(h_ref)
    for rb, tr in ((unet.mid_res0, None), (None, unet.mid_attn), (unet.mid_res1, None)):
        if rb is not None:
            r = O.resnet_block(usd, rb.prefix, h_ref, temb_ref, gN, eps)
            err(rb.prefix, rb(ops, nhwc16(h_ref), None, rows[rb.prefix]), r)
        else:
            r = O.transformer2d(usd, tr.prefix, h_ref, ctx, tr.heads, gN)
            err(tr.prefix, tr(ops, nhwc16(h_ref), kv[tr.prefix], 77), r)
        h_ref = r
    # ---- decoder blocks in isolation
    for i, (res, att, us) in enumerate(unet.up):
        for rb, tr in zip(res, att):
            sk = skips_ref.pop()
            r = O.resnet_block(usd, rb.prefix, torch.cat([h_ref, sk], 1), temb_ref, gN, eps)
            err(rb.prefix + " (two-source)", rb(ops, nhwc16(h_ref), nhwc16(sk), rows[rb.prefix]), r)
            h_ref = r
            if tr is not None:
                r = O.transformer2d(usd, tr.prefix, h_ref, ctx, tr.heads, gN)
                err(tr.prefix, tr(ops, nhwc16(h_ref), kv[tr.prefix], 77), r)
                h_ref = r
        if us is not None:
            pfx = f"up_blocks.{i}.upsamplers.0.conv"
            r = F.conv2d(F.interpolate(h_ref, scale_factor=2.0, mode="nearest"), usd[f"{pfx}.weight"].float(),
                         usd[f"{pfx}.bias"].float(), padding=1)
            err(pfx, us(ops, ops.upsample_nearest2x(nhwc16(h_ref))), r)
            h_ref = r
    r = F.conv2d(F.silu(F.group_norm(h_ref, gN, usd["conv_norm_out.weight"].float(), usd["conv_norm_out.bias"].float(),
                                     eps)), usd["conv_out.weight"].float(), usd["conv_out.bias"].float(), padding=1)
    n = ops.group_norm(nhwc16(h_ref), unet.out_g, unet.out_b, gN, eps, silu=True)
    out = torch.zeros(B, s, s, 8, dtype=torch.float16, device="cuda")
    err("conv_out", unet.conv_out(ops, n, out=out), r)


if __name__ == "__main__":
    main()
