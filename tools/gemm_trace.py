"""Phase timeline of CTA (0,0,0) of one GEMM (gn_set_gemm_trace).  Usage: python tools/gemm_trace.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genima_b200.ops import Ops  # noqa: E402

ops = Ops(0, workspace_mb=256)
tr = torch.zeros(16, dtype=torch.int64, device="cuda")
names = ["start", "setup", "tma0", "ops0", "mma_issued", "acc_done", "epi_done", "exit", "chunks", "bar", "st_issued", "gn", "st_read"]
for (M, N, K, res) in [(4096, 320, 320, True), (4096, 320, 1280, True), (4096, 960, 320, False), (1024, 640, 640, True),
                       (256, 1280, 1280, True), (64, 1280, 1280, True), (4096, 512, 512, False), (64, 1280, 5120, True),
                       (256, 1280, 5120, True), (64, 1280, 11520, False), (256, 1280, 11520, False)]:
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    b = torch.randn(N, device="cuda")
    kw = dict(bias=b)
    if res:
        kw["residual"] = torch.randn(M, N, device="cuda").half()
    out = ops.linear(a, w, **kw)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for cold in (False, True):
        ops.lib.gn_set_gemm_trace(ops.h, tr.data_ptr())
        if cold:
            flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.linear(a, w, out=out, **kw)
        e1.record()
        torch.cuda.synchronize()
        ops.lib.gn_set_gemm_trace(ops.h, None)
        full = tr.cpu().tolist()
        t = full[:len(names)]
        rel = [(x - t[0]) / 1e3 if x else float("nan") for x in t]
        tr.zero_()
        print(f"M{M} N{N} K{K} cfg{ops.last_gemm_config()} {'cold' if cold else 'warm'} events {e0.elapsed_time(e1) * 1e3:.1f} us | "
              + " ".join(f"{n}={r:.2f}" for n, r in zip(names, rel)))
