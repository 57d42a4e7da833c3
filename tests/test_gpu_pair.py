"""CTA-pair GEMMs (tcgen05 cta_group::2, M = 256 MMAs over two SMs; gn_set_gemm_pair): every launch that can pair must
give the result of the single-CTA kernel -- bit for bit where the summation order is the same (no split-K) -- and match
the fp32 oracle within the north-star tolerance (rtol 1e-3 / atol 1e-4)."""
import pytest
import torch

from conftest import report_close
from oracle import ops_ref

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.float16)


class _Paired:
    """Forces pairs wherever the shape allows them (mode 2), restores the default afterwards."""

    def __init__(self, ops, block_n=0, splits=0):
        self.ops, self.block_n, self.splits = ops, block_n, splits

    def __enter__(self):
        self.ops.handle.check(self.ops.lib.gn_set_gemm_pair(self.ops.h, 2), "gn_set_gemm_pair")
        self.ops.set_gemm_tuning(self.block_n, self.splits)
        return self

    def __exit__(self, *exc):
        self.ops.set_gemm_tuning(0, 0)
        self.ops.handle.check(self.ops.lib.gn_set_gemm_pair(self.ops.h, 1), "gn_set_gemm_pair")

    def used(self):
        return self.ops.lib.gn_last_gemm_pair(self.ops.h) == 1


@pytest.mark.parametrize("M,N,K,block_n", [
    (256, 128, 64, 128), (256, 256, 512, 256), (4096, 320, 320, 160), (4096, 320, 320, 64), (1024, 640, 640, 128),
    (4096, 960, 320, 192), (512, 48, 72, 48), (500, 2048, 256, 256), (4096, 512, 4096, 256), (256, 1280, 1280, 80),
])
def test_pair_linear_plain(ops, M, N, K, block_n):
    a = _rand((M, K), 1)
    w = _rand((N, K), 2, K ** -0.5)
    bias = torch.randn(N) * 0.1
    res = _rand((M, N), 3)
    ad, wd, bd, rd = a.cuda(), w.cuda(), bias.cuda(), res.cuda()
    ops.handle.check(ops.lib.gn_set_gemm_pair(ops.h, 0), "gn_set_gemm_pair")
    ops.set_gemm_tuning(block_n, 1)
    try:
        single = ops.linear(ad, wd, bias=bd, residual=rd)
    finally:
        ops.set_gemm_tuning(0, 0)
        ops.handle.check(ops.lib.gn_set_gemm_pair(ops.h, 1), "gn_set_gemm_pair")
    with _Paired(ops, block_n, 1) as pm:
        out = ops.linear(ad, wd, bias=bd, residual=rd)
        assert pm.used(), f"pairs not applied: cfg={ops.last_gemm_config()}"
    report_close(f"pair linear {M}x{N}x{K} bn{block_n}", out, ops_ref.linear_ref(a, w, bias=bias, residual=res))
    assert torch.equal(out, single), "pair and single-CTA kernels disagree (same tile, same summation order)"


@pytest.mark.parametrize("M,N,K,block_n,splits", [(256, 1280, 1280, 128, 2), (256, 1280, 5120, 160, 4),
                                                  (4096, 320, 1280, 160, 2), (1024, 640, 2560, 128, 4),
                                                  (512, 1280, 11520, 96, 3)])
def test_pair_linear_split_k(ops, M, N, K, block_n, splits):
    a = _rand((M, K), 5)
    w = _rand((N, K), 6, K ** -0.5)
    bias = torch.randn(N) * 0.1
    res = _rand((M, N), 7)
    with _Paired(ops, block_n, splits) as pm:
        out = ops.linear(a.cuda(), w.cuda(), bias=bias.cuda(), residual=res.cuda(), act_pre="silu")
        cfg = ops.last_gemm_config()
        assert pm.used() and cfg[1] == splits, f"pairs / splits not applied: cfg={cfg}"
    report_close(f"pair split-K {M}x{N}x{K} x{splits}", out,
                 ops_ref.linear_ref(a, w, bias=bias, residual=res, act_pre="silu"))


@pytest.mark.parametrize("act_pre,act_post", [("silu", None), ("gelu", None), ("relu", None), ("quick_gelu", None),
                                              (None, "relu")])
def test_pair_linear_epilogue(ops, act_pre, act_post):
    M, N, K = 512, 320, 320
    a = _rand((M, K), 10)
    w = _rand((N, K), 11, K ** -0.5)
    bias = torch.randn(N) * 0.1
    scale = 1.0 + 0.1 * torch.randn(N)
    rowvec = torch.randn(2, N) * 0.1
    res = _rand((M, N), 12)
    with _Paired(ops) as pm:
        out = ops.linear(a.cuda(), w.cuda(), bias=bias.cuda(), scale=scale.cuda(), rowvec=rowvec.cuda(),
                         rows_per_batch=256, residual=res.cuda(), act_pre=act_pre, act_post=act_post)
        assert pm.used()
    ref = ops_ref.linear_ref(a, w, bias=bias, scale=scale, rowvec=rowvec, rows_per_batch=256, residual=res,
                             act_pre=act_pre, act_post=act_post)
    report_close(f"pair linear epilogue {act_pre}/{act_post}", out, ref)


@pytest.mark.parametrize("M,D,K", [(4096, 1280, 320), (1024, 2560, 640), (256, 5120, 1280)])
def test_pair_linear_geglu(ops, M, D, K):
    from genima_b200.packing import pack_geglu_weight

    a = _rand((M, K), 15)
    w = _rand((2 * D, K), 16, K ** -0.5)
    b = torch.randn(2 * D) * 0.1
    wp, bp = pack_geglu_weight(w, b)
    with _Paired(ops) as pm:
        out = ops.linear(a.cuda(), wp.cuda(), bias=bp.cuda(), geglu=True)
        assert pm.used()
    proj = a.float() @ w.float().t() + b
    report_close(f"pair geglu {M}x{D}x{K}", out, proj[:, :D] * torch.nn.functional.gelu(proj[:, D:]))


@pytest.mark.parametrize("B,H,Cin,Cout,stride,bucket", [
    (1, 64, 320, 320, 1, 10), (1, 32, 640, 640, 1, 10), (1, 16, 1280, 1280, 1, 10), (1, 16, 640, 1280, 1, 10),
    (1, 64, 320, 320, 2, 10), (1, 128, 128, 128, 1, 4), (2, 16, 64, 64, 1, 2), (4, 8, 64, 128, 1, 2),
])
def test_pair_conv_with_gn_statistics(ops, B, H, Cin, Cout, stride, bucket):
    """3x3 convolution through pairs (each CTA of a pair has its own pixel tile and halo shifts) + the GroupNorm statistics
    its epilogue accumulates for the consumer."""
    from genima_b200.packing import pack_conv_weight

    x = _rand((B, H, H, Cin), 23)
    w = _rand((Cout, Cin, 3, 3), 24, (9 * Cin) ** -0.5)
    bias = torch.randn(Cout) * 0.5
    ops.gn_stats_reset()
    with _Paired(ops) as pm:
        y = ops.conv2d(x.cuda(), pack_conv_weight(w).cuda(), Cout, stride=stride, bias=bias.cuda(), gn_stats=bucket)
        assert pm.used(), f"pairs not applied: cfg={ops.last_gemm_config()}"
    report_close(f"pair conv B{B} {H}^2 {Cin}->{Cout} s{stride}", y, ops_ref.conv2d_ref(x, w, stride=stride, pad=1, bias=bias))
    gamma = 1.0 + 0.1 * torch.randn(Cout)
    beta = 0.1 * torch.randn(Cout)
    groups = 32 if (Cout // 32) % bucket == 0 else Cout // bucket
    assert getattr(y, "gn_stats", None) is not None
    calls0 = ops.gn_apply_calls
    out = ops.group_norm(y, gamma.cuda(), beta.cuda(), groups=groups, eps=1e-5, silu=True)
    assert ops.gn_apply_calls == calls0 + 1
    report_close("pair conv -> fused GN", out, ops_ref.group_norm_ref(y.cpu(), gamma, beta, groups, 1e-5, True))


def test_pair_autotuned_default_matches_oracle(ops):
    """Default mode (pairs are candidates of the measured tile search): whatever wins must be correct."""
    for (M, N, K) in [(4096, 320, 1280), (1024, 640, 640), (4096, 2560, 320)]:
        a = _rand((M, K), 40)
        w = _rand((N, K), 41, K ** -0.5)
        out = ops.linear(a.cuda(), w.cuda())
        print(f"autotuned {M}x{N}x{K}: cfg={ops.last_gemm_config()} pair={ops.lib.gn_last_gemm_pair(ops.h)}")
        report_close(f"autotuned {M}x{N}x{K}", out, ops_ref.linear_ref(a, w))
