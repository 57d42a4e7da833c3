import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genima_b200.ops import Ops
ops = Ops(0)
def bench(B, heads, T, Tk, mode):
    ops.set_attention_kv_split(mode)
    q = torch.randn(B * T, heads * 64, device="cuda").half(); k = torch.randn(B * Tk, heads * 64, device="cuda").half(); v = torch.randn_like(k)
    for _ in range(5): ops.attention(q, k, v, B, heads, T, Tk, 0.125)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): ops.attention(q, k, v, B, heads, T, Tk, 0.125)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 50 * 1e3
for shape in [(1, 5, 4096, 4096), (2, 5, 4096, 4096), (4, 5, 4096, 4096), (1, 10, 1024, 1024), (1, 20, 256, 256)]:
    print(shape, "us: off %.1f auto %.1f forced %.1f" % tuple(bench(*shape, m) for m in (0, 1, 2)), flush=True)
