"""Weight-streaming GEMM-class launches with L2-warm vs L2-cold weights, inside a CUDA graph (as in the agent step, where
2.5 GB of weights pass through the 126 MB L2 per denoise iteration).  Usage: python tools/cold_bench.py [prefetch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genima_b200.ops import Ops  # noqa: E402
from genima_b200.packing import pack_conv_weight  # noqa: E402


def graph_time(fns, reps=5):
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(fns))


ops = Ops(0)
NW = 10
for (H, Cin, Cout) in [(8, 1280, 1280), (8, 2560, 1280), (16, 1280, 1280), (16, 2560, 1280), (32, 640, 640), (32, 1280, 640)]:
    x = torch.randn(1, H, H, Cin, device="cuda").half()
    ws = [pack_conv_weight((torch.randn(Cout, Cin, 3, 3) * (Cin * 9) ** -0.5).half()).cuda() for _ in range(NW)]
    bias = torch.randn(Cout, device="cuda")
    out = ops.conv2d(x, ws[0], Cout, bias=bias)
    warm = graph_time([lambda: ops.conv2d(x, ws[0], Cout, bias=bias, out=out)] * 20)
    cold = graph_time([(lambda w=w: ops.conv2d(x, w, Cout, bias=bias, out=out)) for w in ws] * 2)
    mb = ws[0].numel() * 2 / 1e6
    print(f"conv {H:2d}^2 {Cin:4d}->{Cout:4d}: {mb:5.1f} MB of weights | warm {warm:6.2f} us | cold {cold:6.2f} us ({mb / cold * 1e-3:.2f} TB/s) "
          f"cfg={ops.last_gemm_config()}", flush=True)
for (M, N, K) in [(64, 1280, 5120), (256, 1280, 5120), (64, 10240, 1280), (256, 10240, 1280), (64, 3840, 1280), (1024, 5120, 640)]:
    a = torch.randn(M, K, device="cuda").half()
    ws = [(torch.randn(N, K, device="cuda") * K ** -0.5).half() for _ in range(NW)]
    out = ops.linear(a, ws[0])
    warm = graph_time([lambda: ops.linear(a, ws[0], out=out)] * 20)
    cold = graph_time([(lambda w=w: ops.linear(a, w, out=out)) for w in ws] * 2)
    mb = N * K * 2 / 1e6
    print(f"linear {M:4d}x{N:5d}x{K:4d}: {mb:5.1f} MB of weights | warm {warm:6.2f} us | cold {cold:6.2f} us ({mb / cold * 1e-3:.2f} TB/s) "
          f"cfg={ops.last_gemm_config()}", flush=True)
