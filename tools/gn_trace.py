import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genima_b200.ops import Ops
ops = Ops(0)
tr = torch.zeros(8, dtype=torch.int64, device="cuda")
for (hw, c) in [(8, 1280), (64, 320), (32, 640), (256, 256)]:
    x = torch.randn(1, hw, hw, c, device="cuda").half()
    g, b = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    out = torch.empty_like(x)
    for i in range(3):
        ops.lib.gn_set_gemm_trace(ops.h, tr.data_ptr())
        ops.group_norm(x, g, b, 32, 1e-5, silu=True, out=out)
        torch.cuda.synchronize()
        ops.lib.gn_set_gemm_trace(ops.h, None)
        t = tr.cpu().tolist()
        print(hw, c, [round((v - t[0]) / 1e3, 2) for v in t[:6]])
