"""Weight re-layouts consumed by the kernels (genima_b200/packing.py): pure index arithmetic, checked exactly."""
import torch

from genima_b200.packing import pack_conv_weight, pack_geglu_weight, pad_cols


def test_conv_pack_is_tap_major_with_channel_padding():
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3).half()
    p = pack_conv_weight(w)
    assert p.shape == (2, 9 * 64)
    for co in range(2):
        for ky in range(3):
            for kx in range(3):
                blk = p[co, (ky * 3 + kx) * 64:(ky * 3 + kx + 1) * 64]
                assert torch.equal(blk[:3], w[co, :, ky, kx]) and not blk[3:].any()


def test_conv_pack_extras_and_segment_layout():
    w = torch.randn(4, 6, 3, 3).half()
    e0, e1 = torch.randn(4, 70).half(), torch.randn(4, 8).half()
    p = pack_conv_weight(w, extras=[e0, e1])
    assert p.shape == (4, 9 * 64 + 128 + 64)
    assert torch.equal(p[:, 576:646], e0) and not p[:, 646:704].any()
    assert torch.equal(p[:, 704:712], e1) and not p[:, 712:].any()
    q = pack_conv_weight(w, cin_layout=(4, 8, 2, 8))       # two padded activation segments: 4 real of 8, 2 real of 8
    assert q.shape == (4, 9 * 64)
    assert torch.equal(q[:, 0:4], w[:, 0:4, 0, 0]) and not q[:, 4:8].any()
    assert torch.equal(q[:, 8:10], w[:, 4:6, 0, 0]) and not q[:, 10:64].any()


def test_geglu_pack_interleaves_value_and_gate_blocks():
    d, k = 128, 16
    w = torch.randn(2 * d, k).half()
    b = torch.randn(2 * d)
    wp, bp = pack_geglu_weight(w, b)
    for blk in range(d // 64):
        assert torch.equal(wp[blk * 128:blk * 128 + 64], w[blk * 64:(blk + 1) * 64])              # values
        assert torch.equal(wp[blk * 128 + 64:blk * 128 + 128], w[d + blk * 64:d + (blk + 1) * 64])  # gates
        assert torch.equal(bp[blk * 128 + 64:blk * 128 + 128], b[d + blk * 64:d + (blk + 1) * 64])


def test_pad_cols():
    w = torch.randn(3, 13).half()
    p = pad_cols(w, 8)
    assert p.shape == (3, 16) and torch.equal(p[:, :13], w) and not p[:, 13:].any()


def test_upsample_conv_phase_kernels_reproduce_the_3x3_convolution():
    """conv3x3(pad 1)(nearest x2 upsample(x)) == four 2x2 phase convolutions over x with summed taps (the identity behind
    gn_conv2d_up2x), checked in fp64 on the unpacked effective kernels."""
    import torch
    import torch.nn.functional as F

    from genima_b200.packing import pack_upsample_conv_weight

    g = torch.Generator().manual_seed(0)
    cin, cout, H, W = 64, 8, 6, 8
    w = torch.randn(cout, cin, 3, 3, generator=g).half()
    x = torch.randn(2, cin, H, W, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w.double(), padding=1)
    w4 = pack_upsample_conv_weight(w)
    assert tuple(w4.shape) == (4, cout, 4 * 64)
    out = torch.zeros_like(ref)
    for py in (0, 1):
        for px in (0, 1):
            eff = w4[py * 2 + px].double().reshape(cout, 2, 2, 64).permute(0, 3, 1, 2)   # [Cout, Cin, ty, tx]
            # taps read input rows {y - 1, y} (parity 0) or {y, y + 1} (parity 1): pad one row / column on that side
            xp = F.pad(x, (1 - px, px, 1 - py, py))
            out[:, :, py::2, px::2] = F.conv2d(xp, eff)
    # the only difference is the fp16 rounding of the summed taps
    assert float((out - ref).abs().max()) < 2e-2 * float(ref.abs().max())
    eff_exact = torch.zeros(cout, cin, 2, 2, dtype=torch.float64)
    eff_exact[:, :, 0, 0] = w[:, :, 0, 0].double()
    eff_exact[:, :, 0, 1] = (w[:, :, 0, 1].double() + w[:, :, 0, 2].double())
    eff_exact[:, :, 1, 0] = (w[:, :, 1, 0].double() + w[:, :, 2, 0].double())
    eff_exact[:, :, 1, 1] = w[:, :, 1:, 1:].double().sum(dim=(2, 3))
    got00 = w4[0].double().reshape(cout, 2, 2, 64).permute(0, 3, 1, 2)
    assert float((got00 - eff_exact).abs().max()) < 4e-3


def test_layer_norm_fold_is_the_same_affine_map():
    """fold_layer_norm: LayerNorm(x) @ W^T + b == rstd * (x @ Wg^T - mean * colsum) + bias_f (the epilogue's formula,
    gn_epilogue.ln_*), checked in fp64 with the fp16-rounded folded weight the tensor core would multiply."""
    import torch

    from genima_b200.packing import fold_layer_norm

    g = torch.Generator().manual_seed(0)
    K, N, M = 96, 40, 17
    x = torch.randn(M, K, generator=g, dtype=torch.float64) * 3 + 0.5
    w = (torch.randn(N, K, generator=g) * K ** -0.5).half()
    gamma, beta, bias = torch.randn(K, generator=g) * 0.3 + 1, torch.randn(K, generator=g) * 0.2, torch.randn(N, generator=g)
    w_g, colsum, bias_f = fold_layer_norm(w, gamma, beta, bias)
    assert w_g.dtype == torch.float16 and colsum.dtype == torch.float32 and bias_f.dtype == torch.float32
    mean = x.mean(dim=1, keepdim=True)
    rstd = (x.var(dim=1, unbiased=False, keepdim=True) + 1e-5).rsqrt()
    folded = rstd * (x @ w_g.double().T - mean * colsum.double()[None]) + bias_f.double()[None]
    # reference with the same fp16-rounded (w * gamma): only the algebra is under test here
    ln_rounded = (x - mean) * rstd
    ref = ln_rounded @ w_g.double().T + (w.double() @ beta.double() + bias.double())[None]
    assert torch.allclose(folded, ref, rtol=1e-5, atol=1e-5)      # (colsum / bias_f are stored in fp32)
    # and against the textbook LayerNorm with unrounded weights, to fp16 weight-rounding accuracy
    ln = torch.nn.functional.layer_norm(x, (K,), gamma.double(), beta.double(), 1e-5)
    full = ln @ w.double().T + bias.double()[None]
    assert float((folded - full).abs().max()) < 5e-3 * float(full.abs().max())
