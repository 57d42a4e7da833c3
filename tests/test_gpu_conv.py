"""Parity of the implicit-GEMM convolution (gn_conv2d: TMA-shifted NHWC tiles + tcgen05) against F.conv2d on CPU."""
import pytest
import torch

from conftest import report_close
from oracle import ops_ref

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.float16)


def _conv_case(ops, B, H, W, Cin, Cout, k=3, stride=1, pad=1, seed=0, cpad=None, **epi):
    from genima_b200.packing import pack_conv_weight

    cpad = cpad or Cin
    x = torch.zeros(B, H, W, cpad, dtype=torch.float16)
    x[..., :Cin] = _rand((B, H, W, Cin), seed)
    w = _rand((Cout, Cin, k, k), seed + 1, (Cin * k * k) ** -0.5)
    wp = pack_conv_weight(w, cin_layout=(Cin, cpad))
    gpu_epi = {kk: (v.cuda() if isinstance(v, torch.Tensor) else v) for kk, v in epi.items()}
    out = ops.conv2d(x.cuda(), wp.cuda(), Cout, ksize=k, stride=stride, pad=pad, **gpu_epi)
    ref = ops_ref.conv2d_ref(x, w, stride=stride, pad=pad, **epi)
    report_close(f"conv B{B} {H}x{W} {Cin}->{Cout} k{k} s{stride}", out, ref)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [
    (1, 64, 64, 64, 64), (1, 64, 64, 320, 320), (1, 32, 32, 640, 640), (1, 16, 16, 1280, 1280),
    (1, 8, 8, 1280, 1280), (2, 8, 8, 128, 192), (4, 64, 64, 64, 64), (1, 128, 128, 128, 64), (3, 16, 16, 64, 32),
    (1, 4, 4, 64, 64), (1, 256, 256, 64, 16),
])
def test_conv3x3(ops, B, H, W, Cin, Cout):
    _conv_case(ops, B, H, W, Cin, Cout)


def test_conv3x3_small_channels_padded(ops):
    # 3-channel image carried in a 64-channel padded tensor (ControlNet cond embedding, ResNet stem input)
    _conv_case(ops, 1, 64, 64, 3, 16, cpad=64, seed=3)
    _conv_case(ops, 1, 64, 64, 4, 320, cpad=64, seed=4)
    _conv_case(ops, 1, 32, 32, 16, 32, cpad=16, seed=5)   # C = 16 (< one 64-channel k-block): TMA clips the box
    _conv_case(ops, 1, 32, 32, 96, 96, cpad=96, seed=6)   # C = 96: second k-block half out of bounds


def test_conv3x3_out_channels_tail(ops):
    _conv_case(ops, 1, 64, 64, 320, 4, seed=7)    # U-Net conv_out
    _conv_case(ops, 1, 64, 64, 128, 3, seed=8)    # VAE conv_out


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 64, 64, 320, 320), (1, 16, 16, 1280, 1280), (2, 32, 32, 64, 128),
                                             (4, 8, 8, 64, 64)])
def test_conv3x3_stride2(ops, B, H, W, Cin, Cout):
    _conv_case(ops, B, H, W, Cin, Cout, stride=2, seed=9)


@pytest.mark.parametrize("B,H,W,Cin,Cout,stride,pads", [
    (1, 64, 64, 128, 128, 2, (0, 0, 1, 1)),     # VAE-encoder Downsample2D: F.pad(x, (0, 1, 0, 1)) + stride-2 conv
    (2, 32, 32, 64, 128, 2, (0, 0, 1, 1)),
    (1, 16, 16, 512, 512, 2, (0, 0, 1, 1)),
    (1, 32, 32, 64, 64, 1, (1, 0, 1, 2)),       # stride 1 with four different pads
    (1, 16, 16, 64, 64, 2, (1, 1, 0, 0)),
])
def test_conv3x3_asymmetric_padding(ops, B, H, W, Cin, Cout, stride, pads):
    """gn_conv2d_asym (padding = TMA zero fill at shifted coordinates) against F.pad + F.conv2d."""
    import torch.nn.functional as F

    from genima_b200.packing import pack_conv_weight

    x = _rand((B, H, W, Cin), 31)
    w = _rand((Cout, Cin, 3, 3), 32, (Cin * 9) ** -0.5)
    bias = _rand((Cout,), 33).float()
    pt, pl, pb, pr = pads
    ref = F.conv2d(F.pad(x.float().permute(0, 3, 1, 2), (pl, pr, pt, pb)), w.float(), bias, stride=stride)
    out = ops.conv2d_asym(x.cuda(), pack_conv_weight(w).cuda(), Cout, ksize=3, stride=stride, pads=pads,
                          bias=bias.cuda())
    report_close(f"asym conv B{B} {H}x{W} {Cin}->{Cout} s{stride} pads{pads}", out, ref.permute(0, 2, 3, 1))


def test_conv7x7_stride2_stem(ops):
    _conv_case(ops, 2, 64, 64, 3, 64, k=7, stride=2, pad=3, cpad=64, seed=10)


def test_conv1x1(ops):
    _conv_case(ops, 1, 32, 32, 640, 640, k=1, pad=0, seed=11)
    _conv_case(ops, 2, 16, 16, 128, 256, k=1, stride=2, pad=0, seed=12)   # ResNet downsample shortcut


def test_conv_epilogue_resblock(ops):
    # conv1 of a ResnetBlock2D: + bias + time-embedding row vector; conv2: + bias + residual
    N = 320
    bias = torch.randn(N) * 0.1
    temb = torch.randn(2, N) * 0.1
    _conv_case(ops, 2, 32, 32, 320, N, seed=13, bias=bias, rowvec=temb)
    res = _rand((2, 32, 32, N), 14)
    _conv_case(ops, 2, 32, 32, 320, N, seed=15, bias=bias, residual=res)
    scale = 1.0 + 0.1 * torch.randn(N)
    _conv_case(ops, 2, 32, 32, 320, N, seed=16, bias=bias, scale=scale, residual=res, act_post="relu")


def test_conv_fused_shortcut_two_sources(ops):
    # out = conv3x3(h) + W_sc @ concat(x0, x1) + bias: ResnetBlock2D(conv_shortcut) on a skip-concatenated input
    from genima_b200.packing import pack_conv_weight

    B, H, W, C, C0, C1, Cout = 1, 16, 16, 128, 128, 64, 128
    h = _rand((B, H, W, C), 20)
    x0 = _rand((B, H, W, C0), 21)
    x1 = _rand((B, H, W, C1), 22)
    w = _rand((Cout, C, 3, 3), 23, (C * 9) ** -0.5)
    wsc = _rand((Cout, C0 + C1), 24, (C0 + C1) ** -0.5)
    bias = torch.randn(Cout) * 0.1
    wp = pack_conv_weight(w, extras=[wsc[:, :C0], wsc[:, C0:]])
    out = ops.conv2d(h.cuda(), wp.cuda(), Cout, extras=[x0.cuda(), x1.cuda()], bias=bias.cuda())
    ref = ops_ref.conv2d_ref(h, w, extras=[x0, x1], extra_weights=[wsc[:, :C0], wsc[:, C0:]], bias=bias)
    report_close("conv + fused 1x1 shortcut over two sources", out, ref)


@pytest.mark.parametrize("splits", [2, 5])
def test_conv_split_k(ops, splits):
    ops.set_gemm_tuning(128, splits)
    try:
        bias = torch.randn(1280) * 0.1
        _conv_case(ops, 1, 8, 8, 1280, 1280, seed=30, bias=bias)
        assert ops.last_gemm_config()[1] == splits
    finally:
        ops.set_gemm_tuning(0, 0)


def test_conv_heuristic_uses_split_k_for_weight_bound_levels(ops):
    bias = torch.randn(1280) * 0.1
    _conv_case(ops, 1, 8, 8, 2560, 1280, seed=31, bias=bias)
    assert ops.last_gemm_config()[1] > 1


@pytest.mark.parametrize("B,H,Cin,Cout,bucket", [(1, 32, 640, 640, 10), (1, 8, 1280, 1280, 10), (2, 16, 128, 64, 0),
                                                 (1, 64, 256, 256, 4)])
def test_conv_up2x_matches_upsample_then_conv(ops, B, H, Cin, Cout, bucket):
    """gn_conv2d_up2x (four 2x2 phase convolutions, strided TMA stores) against conv3x3(nearest x2 upsample) of the oracle;
    with a bucket the GroupNorm statistics accumulated over the four launches are checked through gn_group_norm_apply."""
    import torch.nn.functional as F

    from genima_b200.packing import pack_upsample_conv_weight

    g = torch.Generator().manual_seed(H + Cin)
    x = torch.randn(B, H, H, Cin, generator=g).to(torch.float16)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (9 * Cin) ** -0.5).to(torch.float16)
    bias = torch.randn(Cout, generator=g) * 0.3
    ops.gn_stats_reset()
    out = ops.conv2d_up2x(x.cuda(), pack_upsample_conv_weight(w).cuda(), Cout, bias=bias.cuda(), gn_stats=bucket)
    ref = F.conv2d(F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest"), w.float(), bias,
                   padding=1).permute(0, 2, 3, 1)
    # the summed taps are rounded to fp16 once: slightly looser than the plain convolution tolerance
    report_close(f"conv_up2x B{B} {H}^2 {Cin}->{Cout}", out, ref, rtol=2e-3, atol=2e-3)
    if bucket:
        gamma, beta = 1.0 + 0.1 * torch.randn(Cout, generator=g), 0.1 * torch.randn(Cout, generator=g)
        calls0 = ops.gn_apply_calls
        n = ops.group_norm(out, gamma.cuda(), beta.cuda(), groups=32, eps=1e-5, silu=True)
        assert ops.gn_apply_calls == calls0 + 1
        report_close("GN after conv_up2x", n, ops_ref.group_norm_ref(out.cpu(), gamma, beta, 32, 1e-5, True))
