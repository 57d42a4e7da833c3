"""Runs one op a few times; the LAST call sits between cudaProfilerStart/Stop so that
`ncu --profile-from-start off --set full ...` captures the tuned configuration, not an autotuning candidate.
Usage: python tools/one_gemm.py linear M N K [geglu|fp32out|bias_res|plain]   |   conv B H Cin Cout [stride]
       |   attn Tq Tk heads   |   xattn Tq Tk heads (query projection inside)   |   gnapply hw C bucket   |   gn hw C
       |   ln rows C   |   softmax rows cols   |   u8 (uint8 image -> NHWC fp16 and back)   |   tile (tile / untile views)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genima_b200.ops import Ops  # noqa: E402
from genima_b200.packing import pack_conv_weight, pack_geglu_weight  # noqa: E402

ops = Ops(0, workspace_mb=256)
kind = sys.argv[1]
if kind == "linear":
    M, N, K = (int(v) for v in sys.argv[2:5])
    mode = sys.argv[5] if len(sys.argv) > 5 else "plain"
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    b = torch.randn(N, device="cuda")
    kw = {}
    if mode == "geglu":
        w, b = pack_geglu_weight(w, b)
        kw = dict(bias=b, geglu=True)
    elif mode == "fp32out":
        kw = dict(out_fp32=True)
    elif mode == "bias_res":
        kw = dict(bias=b, residual=torch.randn(M, N, device="cuda").half())
    fn = lambda: ops.linear(a, w, **kw)
elif kind == "attn":
    tq, tk, heads = (int(v) for v in sys.argv[2:5])
    c = heads * 64
    q = torch.randn(tq, c, device="cuda").half()
    k = torch.randn(tk, c, device="cuda").half()
    v = torch.randn(tk, c, device="cuda").half()
    fn = lambda: ops.attention(q, k, v, 1, heads, tq, tk, 0.125)
elif kind == "xattn":
    tq, tk, heads = (int(v) for v in sys.argv[2:5])
    c = heads * 64
    x = torch.randn(tq, c, device="cuda").half()
    wq = (torch.randn(c, c, device="cuda") * c ** -0.5).half()
    k = torch.randn(tk, c, device="cuda").half()
    v = torch.randn(tk, c, device="cuda").half()
    bq = torch.randn(c, device="cuda")
    fn = lambda: ops.attention_qproj(x, wq, k, v, 1, heads, tq, tk, 0.125, bias=bq)
elif kind == "ln":
    rows, c = int(sys.argv[2]), int(sys.argv[3])
    x = torch.randn(rows, c, device="cuda").half()
    g, b = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    fn = lambda: ops.layer_norm(x, g, b)
elif kind == "softmax":
    rows, cols = int(sys.argv[2]), int(sys.argv[3])
    x = torch.randn(rows, cols, device="cuda")
    fn = lambda: ops.softmax_rows(x, 0.044)
elif kind == "u8":
    img = torch.randint(0, 255, (1, 512, 512, 3), dtype=torch.uint8, device="cuda")
    fn = lambda: ops.nhwc_to_u8(ops.u8_to_nhwc(img, cpad=8)[..., :8].contiguous())
elif kind == "tile":
    views = torch.randint(0, 255, (1, 4, 256, 256, 3), dtype=torch.uint8, device="cuda")
    fn = lambda: ops.untile_views(ops.tile_views(views))
elif kind == "gnapply":
    # GroupNorm from statistics accumulated by the producing GEMM epilogue: `gnapply hw C bucket`
    hw, c, bucket = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    a = torch.randn(hw * hw, 64, device="cuda").half()
    w = torch.randn(c, 64, device="cuda").half() * 0.125
    ops.gn_stats_reset()
    y = ops.linear(a, w, gn_stats=bucket)
    x = ops.carry_stats(y.reshape(1, hw, hw, c), y)
    g, b = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    fn = lambda: ops.group_norm(x, g, b, 32, 1e-5, silu=True)
elif kind == "gn":
    hw, c = int(sys.argv[2]), int(sys.argv[3])
    x = torch.randn(1, hw, hw, c, device="cuda").half()
    g, b = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    fn = lambda: ops.group_norm(x, g, b, 32, 1e-5, silu=True)
else:
    B, H, Cin, Cout = (int(v) for v in sys.argv[2:6])
    s = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    wp = pack_conv_weight((torch.randn(Cout, Cin, 3, 3) * (Cin * 9) ** -0.5).half()).cuda()
    bias = torch.randn(Cout, device="cuda")
    fn = lambda: ops.conv2d(x, wp, Cout, stride=s, bias=bias)
for _ in range(4):
    out = fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(ops.last_gemm_config(), float(out.float().abs().mean()))
