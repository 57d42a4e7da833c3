"""Parity of the tcgen05 flash-attention kernel and the SIMT small-attention kernel against fp32 softmax(QK^T)V.

Tolerance: rtol 1e-3 / atol 1e-4 relative to the value scale.  P is rounded to fp16 before the PV MMA (as the
reference's fp16 flash kernels do), so atol is scaled by max|V|.
"""
import math

import pytest
import torch

from conftest import report_close
from oracle import ops_ref

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.float16)


@pytest.mark.parametrize("B,heads,Tq,Tk", [
    (1, 1, 128, 128), (1, 2, 256, 256), (1, 5, 4096, 4096), (1, 10, 1024, 1024), (1, 20, 256, 256), (1, 20, 64, 64),
    (1, 5, 4096, 77), (1, 20, 64, 77), (2, 3, 200, 300), (4, 2, 64, 77), (1, 1, 1, 1), (1, 2, 130, 129),
])
def test_attention_tc(ops, B, heads, Tq, Tk):
    d = 64
    q = _rand((B * Tq, heads * d), 1)
    k = _rand((B * Tk, heads * d), 2)
    v = _rand((B * Tk, heads * d), 3)
    scale = 1.0 / math.sqrt(d)
    out = ops.attention(q.cuda(), k.cuda(), v.cuda(), B, heads, Tq, Tk, scale)
    ref = ops_ref.attention_ref(q, k, v, B, heads, d, Tq, Tk, scale)
    report_close(f"attention B{B} h{heads} {Tq}x{Tk}", out, ref, rtol=1e-3, atol=1e-4 * float(v.abs().max()))


@pytest.mark.parametrize("B,heads,Tq,Tk", [(1, 5, 4096, 4096), (1, 20, 256, 256), (2, 3, 200, 300), (1, 2, 130, 129),
                                           (1, 1, 128, 1000), (3, 2, 64, 257)])
def test_attention_tc_kv_split(ops, B, heads, Tq, Tk):
    """Keys split across a 2-CTA cluster (partials merged through distributed shared memory), forced on for every shape
    with at least two key blocks — ragged tails, a second CTA with a single valid key, batches — and compared both with
    the fp32 reference and with the unsplit kernel."""
    d = 64
    q, k, v = _rand((B * Tq, heads * d), 11), _rand((B * Tk, heads * d), 12), _rand((B * Tk, heads * d), 13)
    scale = d ** -0.5
    ref = ops_ref.attention_ref(q, k, v, B, heads, d, Tq, Tk, scale)
    try:
        ops.set_attention_kv_split(2)
        split = ops.attention(q.cuda(), k.cuda(), v.cuda(), B, heads, Tq, Tk, scale)
        again = ops.attention(q.cuda(), k.cuda(), v.cuda(), B, heads, Tq, Tk, scale)
        ops.set_attention_kv_split(0)
        plain = ops.attention(q.cuda(), k.cuda(), v.cuda(), B, heads, Tq, Tk, scale)
    finally:
        ops.set_attention_kv_split(1)
    assert torch.equal(split, again)                                  # fixed merge order: deterministic
    report_close(f"attention kv-split B{B} h{heads} {Tq}x{Tk}", split, ref, rtol=1e-3, atol=1e-4 * float(v.abs().max()))
    assert float((split.float() - plain.float()).abs().max()) <= 2e-3 * float(v.abs().max())


def test_attention_tc_peaked_and_fused_qkv_layout(ops):
    # q/k/v are column slices of one [M, 3C] buffer (the fused QKV projection output); large logits (peaked softmax)
    B, heads, T, d = 1, 5, 512, 64
    Cn = heads * d
    qkv = _rand((B * T, 3 * Cn), 4, 3.0).cuda()
    scale = 1.0 / math.sqrt(d)
    out = ops.attention(qkv[:, :Cn], qkv[:, Cn:2 * Cn], qkv[:, 2 * Cn:], B, heads, T, T, scale)
    c = qkv.cpu()
    ref = ops_ref.attention_ref(c[:, :Cn], c[:, Cn:2 * Cn], c[:, 2 * Cn:], B, heads, d, T, T, scale)
    report_close("attention peaked / strided", out, ref, rtol=1e-3, atol=1e-4 * float(c.abs().max()))


@pytest.mark.parametrize("B,heads,d,Tq,Tk,causal", [
    (1, 8, 32, 258, 258, False), (1, 8, 32, 20, 258, False), (1, 8, 32, 20, 20, False), (2, 8, 64, 77, 77, True),
    (1, 16, 64, 77, 77, True), (3, 2, 32, 5, 70, False),
])
def test_attention_small(ops, B, heads, d, Tq, Tk, causal):
    q = _rand((B * Tq, heads * d), 5)
    k = _rand((B * Tk, heads * d), 6)
    v = _rand((B * Tk, heads * d), 7)
    scale = 1.0 / math.sqrt(d)
    out = ops.attention_small(q.cuda(), k.cuda(), v.cuda(), B, heads, d, Tq, Tk, scale, causal=causal)
    ref = ops_ref.attention_ref(q, k, v, B, heads, d, Tq, Tk, scale, causal=causal)
    report_close(f"attention_small B{B} h{heads} d{d} {Tq}x{Tk} causal={causal}", out, ref)


@pytest.mark.parametrize("B,heads,Tq,Tk,use_ln", [(1, 5, 4096, 77, True), (1, 10, 1024, 77, True), (1, 20, 256, 77, True),
                                                  (1, 20, 64, 77, True), (2, 3, 200, 77, True), (1, 2, 130, 300, False),
                                                  (1, 5, 4096, 77, False)])
def test_attention_qproj_fused(ops, B, heads, Tq, Tk, use_ln):
    """gn_attention_qproj: the cross-attention query projection (with the block's LayerNorm folded, as BasicTransformerBlock
    norm2 -> attn2.to_q) runs inside the attention kernel.  Checked against (i) the fp32 reference of LayerNorm -> Linear
    -> attention and (ii) the two-launch device path (gn_linear with ln= followed by gn_attention): the fused kernel rounds
    Q to fp16 at the same point, so the two device paths agree to the last bit or two."""
    import torch.nn.functional as F

    from genima_b200.packing import fold_layer_norm

    d = 64
    C = heads * d
    g = torch.Generator().manual_seed(B * 1000 + Tq + heads)
    a = torch.randn(B * Tq, C, generator=g).to(torch.float16)
    w0 = (torch.randn(C, C, generator=g) * C ** -0.5).to(torch.float16)
    res = (torch.randn(B * Tq, C, generator=g) + 0.5).to(torch.float16)          # non-zero row means
    gamma = 1.0 + 0.1 * torch.randn(C, generator=g)
    beta = 0.1 * torch.randn(C, generator=g)
    wq = (torch.randn(C, C, generator=g) * C ** -0.5).to(torch.float16)
    bq = torch.randn(C, generator=g) * 0.1
    k = _rand((B * Tk, C), 21)
    v = _rand((B * Tk, C), 22)
    scale = d ** -0.5
    # x = the output of a producing GEMM (that is where the row statistics of a folded LayerNorm come from)
    st = ops.new_row_stats(B * Tq, C)
    x = ops.linear(a.cuda(), w0.cuda(), residual=res.cuda(), row_stats=st)
    xf = x.cpu().float()
    if use_ln:
        q_ref = F.linear(F.layer_norm(xf, (C,), gamma, beta, 1e-5), wq.float(), bq)
        wg, colsum, bias_f = fold_layer_norm(wq.cuda(), gamma.cuda(), beta.cuda(), bias=bq.cuda())
        ln = (st, colsum, 1e-5)
    else:
        q_ref = F.linear(xf, wq.float(), bq)
        wg, bias_f, ln = wq.cuda(), bq.cuda(), None
    ref = ops_ref.attention_ref(q_ref.to(torch.float16), k, v, B, heads, d, Tq, Tk, scale)
    fused = ops.attention_qproj(x, wg, k.cuda(), v.cuda(), B, heads, Tq, Tk, scale, bias=bias_f, ln=ln)
    q_dev = ops.linear(x, wg, bias=bias_f, ln=ln)
    two = ops.attention(q_dev, k.cuda(), v.cuda(), B, heads, Tq, Tk, scale)
    tol = float(v.abs().max())
    print(f"qproj fused vs two-launch: max |diff| {float((fused.float() - two.float()).abs().max()):.3e}")
    assert float((fused.float() - two.float()).abs().max()) <= 2e-3 * tol
    report_close(f"attention+qproj B{B} h{heads} {Tq}x{Tk} ln={use_ln}", fused, ref, rtol=2e-3, atol=3e-4 * tol)
