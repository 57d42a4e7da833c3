"""Phase timeline of CTA (0,0,0) of one convolution (gn_set_gemm_trace).  Usage: python tools/conv_trace.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genima_b200.ops import Ops  # noqa: E402
from genima_b200.packing import pack_conv_weight  # noqa: E402

ops = Ops(0, workspace_mb=256)
tr = torch.zeros(16, dtype=torch.int64, device="cuda")
names = ["start", "setup", "tma0", "ops0", "mma_issued", "acc_done", "epi_done", "exit", "s8", "s9", "s10", "s11", "s12"]
for (H, Cin, Cout) in [(64, 320, 320), (32, 640, 640), (16, 1280, 1280), (8, 1280, 1280), (256, 256, 256), (128, 512, 512)]:
    x = torch.randn(1, H, H, Cin, device="cuda").half()
    w = (torch.randn(Cout, Cin, 3, 3) * (Cin * 9) ** -0.5).half()
    wp = pack_conv_weight(w).cuda()
    bias = torch.randn(Cout, device="cuda")
    out = ops.conv2d(x, wp, Cout, bias=bias)
    for _ in range(2):
        ops.lib.gn_set_gemm_trace(ops.h, tr.data_ptr())
        torch.cuda.synchronize()
        ops.conv2d(x, wp, Cout, bias=bias, out=out)
        torch.cuda.synchronize()
        ops.lib.gn_set_gemm_trace(ops.h, None)
    t = tr.cpu().tolist()[:len(names)]
    rel = [(v - t[0]) / 1e3 if v else float("nan") for v in t]
    tr.zero_()
    kb = 9 * Cin // 64
    print(f"conv {H}^2 {Cin}->{Cout} cfg{ops.last_gemm_config()} kblocks {kb} | "
          + " ".join(f"{n}={r:.2f}" for n, r in zip(names, rel)), flush=True)
