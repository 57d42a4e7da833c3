"""Parity of the normalisation / element-wise / data-movement kernels against the CPU oracle."""
import math

import numpy as np
import pytest
import torch

from conftest import report_close
from oracle import ops_ref

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0, shift=0.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale + shift).to(torch.float16)


@pytest.mark.parametrize("B,HW,C,groups,silu", [
    (1, 4096, 320, 32, True), (1, 1024, 640, 32, True), (1, 256, 1280, 32, False), (1, 64, 1280, 32, True),
    (2, 64, 2560, 32, True), (1, 16384, 128, 32, True), (1, 4096, 512, 32, False), (3, 100, 64, 32, True),
    (1, 4096, 256, 32, True),
])
def test_group_norm(ops, B, HW, C, groups, silu):
    x = _rand((B, HW, C), 1, 2.0, 0.5)
    gamma = 1.0 + 0.1 * torch.randn(C)
    beta = 0.1 * torch.randn(C)
    out = ops.group_norm(x.cuda(), gamma.cuda(), beta.cuda(), groups=groups, eps=1e-5, silu=silu)
    report_close(f"group_norm B{B} HW{HW} C{C}", out, ops_ref.group_norm_ref(x, gamma, beta, groups, 1e-5, silu))


def test_group_norm_large_mean(ops):
    # |mean| >> std: the shifted-sum statistics must not cancel catastrophically
    x = _rand((1, 1024, 320), 2, 0.05, 30.0)
    gamma = torch.ones(320)
    beta = torch.zeros(320)
    out = ops.group_norm(x.cuda(), gamma.cuda(), beta.cuda(), eps=1e-6)
    report_close("group_norm large mean", out, ops_ref.group_norm_ref(x, gamma, beta, 32, 1e-6), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("C0,C1", [(1280, 1280), (1280, 640), (640, 320), (320, 320), (64, 64)])
def test_group_norm_concat(ops, C0, C1):
    x0 = _rand((1, 8, 8, C0), 3)
    x1 = _rand((1, 8, 8, C1), 4, 1.5)
    gamma = 1.0 + 0.1 * torch.randn(C0 + C1)
    beta = 0.1 * torch.randn(C0 + C1)
    out = ops.group_norm(x0.cuda(), gamma.cuda(), beta.cuda(), silu=True, x1=x1.cuda())
    assert out.shape == (1, 8, 8, C0 + C1)
    report_close(f"group_norm concat {C0}+{C1}", out, ops_ref.group_norm_ref(x0, gamma, beta, 32, 1e-5, True, x1=x1))


@pytest.mark.parametrize("rows,C", [(4096, 320), (1024, 640), (256, 1280), (77, 1024), (258, 256), (20, 256), (3, 512)])
def test_layer_norm(ops, rows, C):
    x = _rand((rows, C), 5, 2.0, 0.3)
    gamma = 1.0 + 0.1 * torch.randn(C)
    beta = 0.1 * torch.randn(C)
    out = ops.layer_norm(x.cuda(), gamma.cuda(), beta.cuda())
    report_close(f"layer_norm {rows}x{C}", out, ops_ref.layer_norm_ref(x, gamma, beta))


def test_softmax_rows(ops):
    x = _rand((512, 4096), 6, 4.0)
    scale = 1.0 / math.sqrt(512)
    out = ops.softmax_rows(x.cuda().clone(), scale)
    report_close("softmax_rows fp16", out, torch.softmax(x.float() * scale, dim=-1), rtol=1e-3, atol=1e-6)
    x32 = x.float() * 37.0
    out = ops.softmax_rows(x32.cuda(), scale)
    assert out.dtype == torch.float16
    report_close("softmax_rows fp32 in", out, torch.softmax(x32 * scale, dim=-1), rtol=1e-3, atol=1e-6)


def test_upsample_maxpool_add_scale(ops):
    x = _rand((2, 8, 8, 64), 7)
    up = ops.upsample_nearest2x(x.cuda())
    ref = torch.nn.functional.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    assert torch.equal(up.cpu().float(), ref.permute(0, 2, 3, 1))
    y = _rand((2, 32, 32, 64), 8)
    mp = ops.maxpool3x3s2(y.cuda())
    refp = torch.nn.functional.max_pool2d(y.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(mp.cpu().float(), refp)
    a, b = _rand((1000,), 9), _rand((1000,), 10)
    report_close("add", ops.add(a.cuda(), b.cuda()), a.float() + b.float(), rtol=1e-3, atol=1e-4)
    report_close("scale", ops.scale(a.cuda(), 1 / 0.18215), a.float() / 0.18215, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("t", [999.0, 799.0, 199.0, 1.0])
def test_timestep_embedding(ops, t):
    out = ops.timestep_embedding(t, 320)
    # cos/sin of arguments up to 999 rad: fp32 argument rounding allows ~1e-4 absolute error before the fp16 rounding
    report_close(f"timestep_embedding t={t}", out, ops_ref.timestep_embedding_ref(t, 320), rtol=1e-3, atol=5e-4)


def test_euler_step(ops):
    x = _rand((1, 64, 64, 4), 11, 14.6)
    eps = _rand((1, 64, 64, 4), 12)
    sigma, sigma_next = 14.6146, 4.0
    xs = torch.empty_like(x).cuda()
    xn, _ = ops.euler_step(x.cuda(), eps.cuda(), sigma, sigma_next, x_scaled=xs)
    dsigma = float(np.float32(sigma_next) - np.float32(sigma))  # the kernel forms the step in fp32
    ref = (x.float() + dsigma * eps.float()).to(torch.float16)
    # fp32 update + one rounding to fp16; an FMA contraction may flip the last fp16 bit of a few elements
    report_close("euler x_next", xn, ref.float(), rtol=1e-3, atol=1e-4)
    assert float((xn.cpu() != ref).float().mean()) < 0.01
    report_close("euler x_scaled", xs, ref.float() / math.sqrt(sigma_next ** 2 + 1))


def test_layout_and_image_conversions(ops):
    g = torch.Generator().manual_seed(13)
    lat = torch.randn(2, 4, 16, 16, generator=g)
    nhwc = ops.nchw_to_nhwc(lat.cuda(), cpad=64)
    assert nhwc.shape == (2, 16, 16, 64)
    assert torch.equal(nhwc[..., :4].cpu(), lat.to(torch.float16).permute(0, 2, 3, 1))
    assert float(nhwc[..., 4:].abs().max()) == 0.0
    back = ops.nhwc_to_nchw(nhwc, channels=4, fp32=True)
    assert torch.equal(back.cpu(), lat.to(torch.float16).float())

    img = torch.randint(0, 256, (2, 32, 32, 3), dtype=torch.uint8, generator=g)
    f = ops.u8_to_nhwc(img.cuda(), cpad=64)
    assert torch.equal(f[..., :3].cpu(), (img.float() / 255.0).to(torch.float16))
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    fn = ops.u8_to_nhwc(img.cuda(), cpad=64, mean=mean, std=std)
    refn = (img.float() / 255.0 - torch.tensor(mean)) / torch.tensor(std)
    report_close("imagenet normalise", fn[..., :3], refn)

    x = (torch.rand(2, 32, 32, 8, generator=g) * 2.4 - 1.2).to(torch.float16)
    u8 = ops.nhwc_to_u8(x.cuda())
    refu = np.round(np.clip(x[..., :3].float().numpy() / 2 + 0.5, 0, 1) * 255).astype(np.uint8)
    assert np.array_equal(u8.cpu().numpy(), refu)


def test_tile_untile_views(ops):
    g = torch.Generator().manual_seed(14)
    views = torch.randint(0, 256, (2, 4, 256, 256, 3), dtype=torch.uint8, generator=g)
    tile = ops.tile_views(views.cuda())
    ref = torch.zeros(2, 512, 512, 3, dtype=torch.uint8)
    ref[:, :256, :256] = views[:, 0]
    ref[:, :256, 256:] = views[:, 1]
    ref[:, 256:, :256] = views[:, 2]
    ref[:, 256:, 256:] = views[:, 3]
    assert torch.equal(tile.cpu(), ref)
    assert torch.equal(ops.untile_views(tile).cpu(), views)


# ---- GroupNorm statistics accumulated by the producing GEMM epilogue (gn_epilogue.gnstats_out) + gn_group_norm_apply
def _fused_gn_check(ops, name, y, gamma, beta, groups, silu, x1=None):
    """y (and x1) carry gn_stats: the apply kernel must agree with the oracle GroupNorm of the SAME fp16 tensors."""
    assert getattr(y, "gn_stats", None) is not None, f"{name}: the producer did not attach statistics"
    calls0 = ops.gn_apply_calls
    out = ops.group_norm(y, gamma.cuda(), beta.cuda(), groups=groups, eps=1e-5, silu=silu, x1=x1)
    assert ops.gn_apply_calls == calls0 + 1, f"{name}: the statistics were not used"
    ref = ops_ref.group_norm_ref(y.cpu(), gamma, beta, groups, 1e-5, silu, x1=None if x1 is None else x1.cpu())
    report_close(name, out, ref)


@pytest.mark.parametrize("B,HW,K,N,bucket,splits", [
    (1, 4096, 320, 320, 10, 0), (1, 1024, 640, 640, 10, 0), (1, 256, 1280, 1280, 10, 0), (2, 64, 1280, 1280, 10, 0),
    (1, 64, 2560, 1280, 10, 4), (1, 256, 1280, 640, 10, 2), (1, 4096, 512, 512, 4, 0), (3, 16, 64, 96, 2, 0),
    (1, 4096, 1280, 320, 10, 2),
])
def test_fused_gn_stats_linear(ops, B, HW, K, N, bucket, splits):
    a = _rand((B * HW, K), 20)
    w = _rand((N, K), 21, K ** -0.5)
    bias = torch.randn(N) * 0.5
    res = _rand((B * HW, N), 22)
    ops.gn_stats_reset()
    if splits:
        ops.set_gemm_tuning(0, splits)
    try:
        y = ops.linear(a.cuda(), w.cuda(), bias=bias.cuda(), residual=res.cuda(), rows_per_batch=HW, gn_stats=bucket)
        cfg = ops.last_gemm_config()
    finally:
        ops.set_gemm_tuning(0, 0)
    if splits:
        assert cfg[1] == splits
    report_close("producer", y, ops_ref.linear_ref(a, w, bias=bias, residual=res))
    groups = 32 if N % 32 == 0 and (N // 32) % bucket == 0 else N // bucket // 2
    gamma = 1.0 + 0.1 * torch.randn(N)
    beta = 0.1 * torch.randn(N)
    _fused_gn_check(ops, f"fused GN linear B{B} HW{HW} N{N} splits{splits}", ops.carry_stats(y.reshape(B, HW, N), y),
                    gamma, beta, groups, True)


@pytest.mark.parametrize("B,H,Cin,Cout,stride,bucket", [
    (1, 64, 320, 320, 1, 10), (1, 32, 640, 640, 1, 10), (2, 8, 1280, 1280, 1, 10), (1, 16, 640, 1280, 1, 10),
    (1, 64, 320, 320, 2, 10), (1, 128, 128, 128, 1, 4), (1, 12, 64, 64, 1, 2), (2, 4, 64, 128, 1, 2),
])
def test_fused_gn_stats_conv(ops, B, H, Cin, Cout, stride, bucket):
    from genima_b200.packing import pack_conv_weight

    x = _rand((B, H, H, Cin), 23)
    w = _rand((Cout, Cin, 3, 3), 24, (9 * Cin) ** -0.5)
    bias = torch.randn(Cout) * 0.5
    ops.gn_stats_reset()
    y = ops.conv2d(x.cuda(), pack_conv_weight(w).cuda(), Cout, stride=stride, bias=bias.cuda(), gn_stats=bucket)
    gamma = 1.0 + 0.1 * torch.randn(Cout)
    beta = 0.1 * torch.randn(Cout)
    groups = 32 if (Cout // 32) % bucket == 0 else Cout // bucket
    _fused_gn_check(ops, f"fused GN conv B{B} {H}^2 {Cin}->{Cout} s{stride}", y, gamma, beta, groups, True)


def test_fused_gn_stats_concat(ops):
    # up-block norm1 over concat(hidden 1280, skip 640): 60 channels per group, buckets of 10 from two producers
    B, HW = 1, 1024
    ops.gn_stats_reset()
    ys = []
    for i, n in enumerate((1280, 640)):
        a = _rand((B * HW, 320), 30 + i)
        w = _rand((n, 320), 40 + i, 320 ** -0.5)
        y = ops.linear(a.cuda(), w.cuda(), rows_per_batch=HW, gn_stats=10)
        ys.append(ops.carry_stats(y.reshape(B, 32, 32, n), y))
    gamma = 1.0 + 0.1 * torch.randn(1920)
    beta = 0.1 * torch.randn(1920)
    _fused_gn_check(ops, "fused GN concat 1280+640", ys[0], gamma, beta, 32, True, x1=ys[1])


def test_fused_gn_stats_expire_on_reset(ops):
    a = _rand((256, 320), 50)
    w = _rand((320, 320), 51, 320 ** -0.5)
    ops.gn_stats_reset()
    y = ops.linear(a.cuda(), w.cuda(), gn_stats=10)
    ops.gn_stats_reset()   # the arena is recycled: y's statistics must not be used any more
    gamma, beta = torch.ones(320), torch.zeros(320)
    calls0 = ops.gn_apply_calls
    out = ops.group_norm(ops.carry_stats(y.reshape(1, 256, 320), y), gamma.cuda(), beta.cuda())
    assert ops.gn_apply_calls == calls0
    report_close("GN after reset", out, ops_ref.group_norm_ref(y.cpu().reshape(1, 256, 320), gamma, beta, 32, 1e-5))
