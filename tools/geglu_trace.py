"""Phase timeline of CTA (0,0,0) of the GEGLU GEMMs (gn_set_gemm_trace).  Usage: python tools/geglu_trace.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genima_b200.ops import Ops  # noqa: E402
from genima_b200.packing import pack_geglu_weight  # noqa: E402

ops = Ops(0, workspace_mb=256)
tr = torch.zeros(16, dtype=torch.int64, device="cuda")
names = ["start", "setup", "tma0", "ops0", "mma_issued", "acc_done", "epi_done", "exit", "chunks", "bar", "st_issued", "gn", "st_read"]
for (M, N, K) in [(4096, 2560, 320), (1024, 5120, 640), (256, 10240, 1280), (64, 10240, 1280)]:
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    b = torch.randn(N, device="cuda")
    wg, bg = pack_geglu_weight(w, b)
    out = ops.linear(a, wg, bias=bg, geglu=True)
    for _ in range(2):
        ops.lib.gn_set_gemm_trace(ops.h, tr.data_ptr())
        torch.cuda.synchronize()
        ops.linear(a, wg, bias=bg, geglu=True, out=out)
        torch.cuda.synchronize()
        ops.lib.gn_set_gemm_trace(ops.h, None)
    t = tr.cpu().tolist()[:len(names)]
    rel = [(v - t[0]) / 1e3 if v else float("nan") for v in t]
    tr.zero_()
    print(f"geglu M{M} N{N} K{K} cfg{ops.last_gemm_config()} | " + " ".join(f"{n}={r:.2f}" for n, r in zip(names, rel)), flush=True)
