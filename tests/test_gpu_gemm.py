"""Parity of the tcgen05 GEMM (gn_linear) against the fp32 CPU oracle, through the C ABI.

Tolerance (BASELINE.json north_star): rtol = 1e-3, atol = 1e-4 on fp16 outputs, operands identical (fp16-rounded).
"""
import pytest
import torch

from conftest import report_close
from oracle import ops_ref

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.float16)


@pytest.mark.parametrize("M,N,K", [
    (128, 128, 64), (256, 128, 128), (128, 256, 512), (4096, 320, 320), (1024, 640, 640), (256, 1280, 1280),
    (64, 1280, 1280), (77, 320, 1024), (77, 1280, 1024), (4096, 960, 320), (100, 48, 72), (1, 1280, 320),
    (20, 8, 256), (258, 2048, 256),
])
def test_linear_plain(ops, M, N, K):
    a = _rand((M, K), 1)
    w = _rand((N, K), 2, K ** -0.5)
    out = ops.linear(a.cuda(), w.cuda())
    report_close(f"linear {M}x{N}x{K}", out, ops_ref.linear_ref(a, w))


@pytest.mark.parametrize("block_n", [16, 32, 48, 64, 96, 128, 160, 192, 256])
def test_linear_block_n(ops, block_n):
    M, N, K = 300, 320, 256
    a = _rand((M, K), 3)
    w = _rand((N, K), 4, K ** -0.5)
    ops.set_gemm_tuning(block_n, 1)
    try:
        out = ops.linear(a.cuda(), w.cuda())
        cfg = ops.last_gemm_config()
    finally:
        ops.set_gemm_tuning(0, 0)
    assert cfg[0] == block_n
    report_close(f"linear block_n={block_n}", out, ops_ref.linear_ref(a, w))


@pytest.mark.parametrize("splits", [2, 3, 8])
def test_linear_split_k(ops, splits):
    M, N, K = 64, 1280, 2560
    a = _rand((M, K), 5)
    w = _rand((N, K), 6, K ** -0.5)
    bias = torch.randn(N) * 0.1
    res = _rand((M, N), 7)
    ops.set_gemm_tuning(128, splits)
    try:
        out = ops.linear(a.cuda(), w.cuda(), bias=bias.cuda(), residual=res.cuda())
        cfg = ops.last_gemm_config()
    finally:
        ops.set_gemm_tuning(0, 0)
    assert cfg[1] == splits
    report_close(f"linear split-K {splits}", out, ops_ref.linear_ref(a, w, bias=bias, residual=res))


def test_linear_long_k_pipeline_wrap(ops):
    # K = 8192 -> 128 k-blocks: the smem ring wraps many times (phase-bit handling)
    M, N, K = 256, 256, 8192
    a = _rand((M, K), 8)
    w = _rand((N, K), 9, K ** -0.5)
    ops.set_gemm_tuning(128, 1)
    try:
        out = ops.linear(a.cuda(), w.cuda())
    finally:
        ops.set_gemm_tuning(0, 0)
    report_close("linear K=8192", out, ops_ref.linear_ref(a, w))


@pytest.mark.parametrize("act_pre,act_post", [("silu", None), ("gelu", None), ("relu", None), ("quick_gelu", None),
                                              (None, "relu")])
def test_linear_epilogue(ops, act_pre, act_post):
    M, N, K = 512, 320, 320
    a = _rand((M, K), 10)
    w = _rand((N, K), 11, K ** -0.5)
    bias = torch.randn(N) * 0.1
    scale = 1.0 + 0.1 * torch.randn(N)
    rowvec = torch.randn(2, N) * 0.1
    res = _rand((M, N), 12)
    out = ops.linear(a.cuda(), w.cuda(), bias=bias.cuda(), scale=scale.cuda(), rowvec=rowvec.cuda(),
                     rows_per_batch=256, residual=res.cuda(), act_pre=act_pre, act_post=act_post)
    ref = ops_ref.linear_ref(a, w, bias=bias, scale=scale, rowvec=rowvec, rows_per_batch=256, residual=res,
                             act_pre=act_pre, act_post=act_post)
    report_close(f"linear epilogue {act_pre}/{act_post}", out, ref)


def test_linear_alpha_beta_fp32_out(ops):
    M, N, K = 20, 8, 256
    a = _rand((M, K), 13)
    w = _rand((N, K), 14, K ** -0.5)
    bias = torch.randn(N) * 0.1
    out = ops.linear(a.cuda(), w.cuda(), bias=bias.cuda(), out_fp32=True)
    assert out.dtype == torch.float32
    report_close("linear fp32 out", out, ops_ref.linear_ref(a, w, bias=bias), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("M,D,K", [(4096, 1280, 320), (64, 5120, 1280), (300, 128, 64)])
def test_linear_geglu(ops, M, D, K):
    from genima_b200.packing import pack_geglu_weight

    a = _rand((M, K), 15)
    w = _rand((2 * D, K), 16, K ** -0.5)
    b = torch.randn(2 * D) * 0.1
    wp, bp = pack_geglu_weight(w, b)
    out = ops.linear(a.cuda(), wp.cuda(), bias=bp.cuda(), geglu=True)
    assert out.shape == (M, D)
    af = a.float()
    proj = af @ w.float().t() + b
    ref = proj[:, :D] * torch.nn.functional.gelu(proj[:, D:])
    report_close(f"geglu {M}x{D}x{K}", out, ref)


def test_linear_strided_views(ops):
    # A is a column slice of a wider buffer (lda > K) and the output lands in a column slice (ldo > N)
    M, N, K = 200, 64, 128
    big = _rand((M, 3 * K), 17).cuda()
    w = _rand((N, K), 18, K ** -0.5)
    outbuf = torch.zeros(M, 2 * N, dtype=torch.float16, device="cuda")
    ops.linear(big[:, K:2 * K], w.cuda(), out=outbuf[:, N:])
    report_close("linear strided", outbuf[:, N:], ops_ref.linear_ref(big[:, K:2 * K].cpu(), w))
    assert float(outbuf[:, :N].abs().max()) == 0.0


@pytest.mark.parametrize("M,N,K", [(64, 1280, 1280), (256, 1280, 5120), (300, 400, 2048)])
def test_forced_tile_configurations_sweep(ops, M, N, K):
    """Every (block_n, cluster split-K, CTAs/SM) combination the autotuner may pick must give the same result: covers the
    operand-ring / fp32 staging-tile aliasing of the cluster reduction (block_n = 192 with a 2-stage ring)."""
    g = torch.Generator().manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g)).to(torch.float16)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).to(torch.float16)
    bias = torch.randn(N, generator=g) * 0.1
    res = torch.randn(M, N, generator=g).to(torch.float16)
    ref = ops_ref.linear_ref(a, w, bias=bias, residual=res)
    ad, wd, bd, rd = a.cuda(), w.cuda(), bias.cuda(), res.cuda()
    try:
        for bn in (256, 192, 160, 96, 48):
            for sp in (1, 2, 5, 8):
                for occ in (1, 2):
                    ops.set_gemm_tuning(bn, sp)
                    ops.lib.gn_set_gemm_occupancy(ops.h, occ)
                    out = ops.linear(ad, wd, bias=bd, residual=rd)
                    err = float((out.float().cpu() - ref).abs().max())
                    assert err < 2e-2, f"bn={bn} splits={sp} occ={occ} cfg={ops.last_gemm_config()}: max abs err {err}"
    finally:
        ops.set_gemm_tuning(0, 0)
        ops.lib.gn_set_gemm_occupancy(ops.h, 0)


@pytest.mark.parametrize("M,C,N,geglu", [(4096, 320, 960, False), (256, 1280, 1280, False), (300, 128, 256, True),
                                         (1024, 640, 5120, True), (64, 1280, 3840, False)])
def test_layer_norm_folded_into_gemm(ops, M, C, N, geglu):
    """BasicTransformerBlock: h = Linear(a) + res; y = Linear2(LayerNorm(h)).  The device path never materialises the
    LayerNorm: the first GEMM's epilogue writes per-row (sum, sumsq) partials of h, the second GEMM runs on the
    un-normalised h with gamma folded into its weight and applies rstd * (acc - mean * colsum) + bias' in its epilogue."""
    import torch.nn.functional as F

    from genima_b200.packing import fold_layer_norm, pack_geglu_weight

    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, C, generator=g).to(torch.float16)
    w0 = (torch.randn(C, C, generator=g) * C ** -0.5).to(torch.float16)
    b0 = torch.randn(C, generator=g) * 0.1
    res = (torch.randn(M, C, generator=g) + 0.5).to(torch.float16)          # non-zero row means
    gamma = 1.0 + 0.1 * torch.randn(C, generator=g)
    beta = 0.1 * torch.randn(C, generator=g)
    w1 = (torch.randn(N, C, generator=g) * C ** -0.5).to(torch.float16)
    b1 = torch.randn(N, generator=g) * 0.1
    # reference: fp32 LayerNorm of the fp16-rounded h, then the second linear (GEGLU: value * gelu(gate))
    h_ref = (a.float() @ w0.float().t() + b0 + res.float()).to(torch.float16)
    y = F.linear(F.layer_norm(h_ref.float(), (C,), gamma, beta, 1e-5), w1.float(), b1)
    ref = y[:, :N // 2] * F.gelu(y[:, N // 2:]) if geglu else y
    # device
    st = ops.new_row_stats(M, C)
    h = ops.linear(a.cuda(), w0.cuda(), bias=b0.cuda(), residual=res.cuda(), row_stats=st)
    assert torch.equal(h.cpu(), h_ref) or float((h.cpu().float() - h_ref.float()).abs().max()) < 2e-2
    assert 0 < st.parts <= st.capacity
    if geglu:
        w1p, b1p = pack_geglu_weight(w1, b1)
    else:
        w1p, b1p = w1, b1
    wg, colsum, bias_f = fold_layer_norm(w1p.cuda(), gamma.cuda(), beta.cuda(), bias=b1p.cuda())
    out = ops.linear(h, wg, bias=bias_f, ln=(st, colsum, 1e-5), geglu=geglu)
    err = float((out.float().cpu() - ref).abs().max() / ref.abs().max())
    print(f"LN folded into GEMM M={M} C={C} N={N} geglu={geglu}: normalised max err {err:.3e}, {st.parts} row partials")
    assert err < 3e-3


@pytest.mark.parametrize("M,N,K,residual", [(4096, 320, 320, True), (1024, 1920, 640, False), (300, 80, 256, True),
                                            (77, 1280, 1024, False), (130, 48, 72, True)])
def test_linear_direct_epilogue_matches_staged(ops, M, N, K, residual):
    """A/B: per-thread global stores (gn_set_staged_epilogue 0) and the TMA-stored staged tile give the same bits."""
    a = _rand((M, K), 30)
    w = _rand((N, K), 31, K ** -0.5)
    bias = torch.randn(N) * 0.1
    kw = dict(bias=bias.cuda())
    if residual:
        kw["residual"] = _rand((M, N), 32).cuda()
    ops.set_gemm_tuning(64 if N >= 64 else 48, 1)   # same tile configuration on both sides: same summation order
    try:
        staged = ops.linear(a.cuda(), w.cuda(), **kw)
        ops.handle.check(ops.lib.gn_set_staged_epilogue(ops.h, 0), "gn_set_staged_epilogue")
        direct = ops.linear(a.cuda(), w.cuda(), **kw)
    finally:
        ops.handle.check(ops.lib.gn_set_staged_epilogue(ops.h, 1), "gn_set_staged_epilogue")
        ops.set_gemm_tuning(0, 0)
    report_close(f"linear staged {M}x{N}x{K}", staged,
                 ops_ref.linear_ref(a, w, bias=bias, residual=kw.get("residual")))
    assert torch.equal(staged, direct)


def test_linear_strided_output_and_residual(ops):
    # out / residual are column slices of wider matrices (row stride > N): staged TMA stores must respect ldo / ldr
    M, N, K = 512, 320, 256
    a = _rand((M, K), 33)
    w = _rand((N, K), 34, K ** -0.5)
    big_res = _rand((M, 2 * N), 35).cuda()
    big_out = torch.zeros(M, 3 * N, dtype=torch.float16, device="cuda")
    ops.linear(a.cuda(), w.cuda(), residual=big_res[:, N:], out=big_out[:, N:2 * N])
    report_close("linear strided", big_out[:, N:2 * N], ops_ref.linear_ref(a, w, residual=big_res[:, N:].cpu()))
    assert float(big_out[:, :N].abs().max()) == 0.0 and float(big_out[:, 2 * N:].abs().max()) == 0.0
