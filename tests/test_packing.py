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
