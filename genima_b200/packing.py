"""One-time weight re-layouts from the upstream (diffusers / torchvision / RoboBase) state-dict layout to the layouts the
sm_100a kernels consume.  Pure data movement, run once at bind time — never on the per-step path.

  conv  [Cout, Cin, KH, KW]  ->  [Cout, KH*KW*Cp (+ Cp_extra...)]   tap-major, channels padded to 64 (zeros)
  GEGLU [2*D, K] (value rows, then gate rows) -> 64-row interleaved (value, gate) blocks
"""
from __future__ import annotations

from typing import Sequence

import torch


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def _pad_last(t: torch.Tensor, n: int) -> torch.Tensor:
    """Append n zeros along the last dim."""
    return torch.cat([t, t.new_zeros(*t.shape[:-1], n)], dim=-1)


def pack_conv_weight(w: torch.Tensor, extras: Sequence[torch.Tensor] = (), cin_layout: Sequence[int] = ()) -> torch.Tensor:
    """w: [Cout, Cin, KH, KW]; extras: 1x1 shortcut weights [Cout, Ce(,1,1)] appended along K.

    cin_layout: when the activation tensor carries padded channels in several segments (e.g. a concat of two padded
    tensors) give the (real, padded) pairs flattened: (r0, p0, r1, p1, ...); default = one segment padded to 64.
    """
    cout, cin, kh, kw = w.shape
    w = w.to(torch.float16)
    taps = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)  # [Cout, taps, Cin]
    if cin_layout:
        segs = []
        off = 0
        for i in range(0, len(cin_layout), 2):
            real, padded = cin_layout[i], cin_layout[i + 1]
            seg = taps[:, :, off:off + real]
            if padded > real:
                seg = _pad_last(seg, padded - real)
            segs.append(seg)
            off += real
        assert off == cin, (off, cin)
        taps = torch.cat(segs, dim=2)
    cp = round_up(taps.shape[2], 64)
    if cp > taps.shape[2]:
        taps = _pad_last(taps, cp - taps.shape[2])
    parts = [taps.reshape(cout, kh * kw * cp)]
    for e in extras:
        e = e.to(torch.float16).reshape(cout, -1)
        ep = round_up(e.shape[1], 64)
        if ep > e.shape[1]:
            e = _pad_last(e, ep - e.shape[1])
        parts.append(e)
    return torch.cat(parts, dim=1).contiguous()


def pack_upsample_conv_weight(w: torch.Tensor) -> torch.Tensor:
    """3x3 convolution that follows a nearest-neighbour x2 upsample (diffusers Upsample2D, TAESD) -> the four 2x2 phase
    kernels of gn_conv2d_up2x, packed [4, Cout, 4 * Cp].

    Output pixel (2y + py, 2x + px) of conv3x3(upsample(x)) reads upsampled rows 2y + py - 1 .. 2y + py + 1, i.e. input rows
    {y - 1, y, y} for py = 0 and {y, y, y + 1} for py = 1: the taps that fall on the same input pixel are summed (in fp32,
    rounded to fp16 once)."""
    assert w.shape[2:] == (3, 3), w.shape
    groups = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}   # parity -> kernel taps per effective tap
    wf = w.float()
    phases = []
    for py in (0, 1):
        for px in (0, 1):
            eff = torch.zeros(w.shape[0], w.shape[1], 2, 2, device=w.device)
            for ty, kys in enumerate(groups[py]):
                for tx, kxs in enumerate(groups[px]):
                    for ky in kys:
                        for kx in kxs:
                            eff[:, :, ty, tx] += wf[:, :, ky, kx]
            phases.append(pack_conv_weight(eff.to(torch.float16)))
    return torch.stack(phases, 0).contiguous()


def pack_geglu_weight(w: torch.Tensor, b: torch.Tensor):
    """diffusers GEGLU: proj = Linear(K, 2D); hidden, gate = proj(x).chunk(2, -1); out = hidden * gelu(gate).
    Re-order rows into 64-wide (value, gate) blocks so one 128-column accumulator group holds matching pairs."""
    two_d, k = w.shape
    d = two_d // 2
    assert d % 64 == 0, d
    wv = w[:d].reshape(d // 64, 64, k)
    wg = w[d:].reshape(d // 64, 64, k)
    wp = torch.stack([wv, wg], dim=1).reshape(two_d, k).contiguous()
    bv = b[:d].reshape(d // 64, 64)
    bg = b[d:].reshape(d // 64, 64)
    bp = torch.stack([bv, bg], dim=1).reshape(two_d).contiguous()
    return wp, bp


def pad_cols(w: torch.Tensor, mult: int = 8) -> torch.Tensor:
    """Pad the K (last) dim of a linear weight with zeros to a multiple of `mult`."""
    k = w.shape[-1]
    kp = round_up(k, mult)
    if kp == k:
        return w.contiguous()
    return _pad_last(w, kp - k).contiguous()


def fold_layer_norm(w: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bias: torch.Tensor = None):
    """LayerNorm(x) @ w^T + bias  ==  rstd * (x @ w_g^T - mean * colsum) + bias_f   with
         w_g    = fp16(w * gamma)          (what the tensor core multiplies)
         colsum = sum_k w_g[n, k]          (fp32, from the fp16-rounded w_g so that it matches the accumulation)
         bias_f = bias + w @ beta          (fp32)
    w: [N, K] (any row order, e.g. GEGLU-interleaved: the fold acts on the K axis only).  Returns (w_g, colsum, bias_f)."""
    wf = w.float()
    w_g = (wf * gamma.float()[None, :]).to(torch.float16).contiguous()
    colsum = w_g.float().sum(dim=1).contiguous()
    bias_f = wf @ beta.float()
    if bias is not None:
        bias_f = bias_f + bias.float()
    return w_g, colsum, bias_f.contiguous()
