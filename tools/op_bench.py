"""Per-op timings inside a CUDA graph (no launch overhead from Python): N back-to-back calls captured, replayed, timed.
Usage: python tools/op_bench.py gn|ln|attn|gemm|conv"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genima_b200.ops import Ops  # noqa: E402
from genima_b200.packing import pack_conv_weight, pack_geglu_weight  # noqa: E402


def graph_time(fn, n=20, reps=5):
    """us per call of fn() when n calls run back to back inside one CUDA graph."""
    fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * n)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "gn"
    ops = Ops(0, workspace_mb=64)
    if which == "gn":
        for (hw, c0, c1) in [(64, 320, 0), (64, 320, 320), (64, 640, 320), (32, 640, 0), (32, 1280, 640), (16, 1280, 0),
                             (16, 1280, 1280), (8, 1280, 0), (8, 1280, 1280), (64, 512, 0), (128, 512, 0), (256, 512, 0),
                             (256, 256, 0), (512, 256, 0), (512, 128, 0)]:
            x0 = torch.randn(1, hw, hw, c0, device="cuda").half()
            x1 = torch.randn(1, hw, hw, c1, device="cuda").half() if c1 else None
            g, b = torch.ones(c0 + c1, device="cuda"), torch.zeros(c0 + c1, device="cuda")
            out = torch.empty(1, hw, hw, c0 + c1, device="cuda", dtype=torch.float16)
            t = graph_time(lambda: ops.group_norm(x0, g, b, 32, 1e-5, silu=True, x1=x1, out=out))
            mb = hw * hw * (c0 + c1) * 4 / 1e6
            print(f"gn {hw:3d}^2 C={c0}+{c1}: {t:7.2f} us  ({mb:6.1f} MB moved -> {mb / t * 1e3:7.1f} GB/s)", flush=True)
    if which == "gnapply":
        # GroupNorm from statistics accumulated by the producing GEMM epilogue (gn_group_norm_apply)
        for (hw, c, bucket) in [(64, 320, 10), (32, 640, 10), (16, 1280, 10), (8, 1280, 10), (64, 512, 4), (128, 512, 4),
                                (256, 256, 4), (512, 128, 4)]:
            a = torch.randn(hw * hw, 64, device="cuda").half()
            w = torch.randn(c, 64, device="cuda").half() * 0.125
            ops.gn_stats_reset()
            y = ops.linear(a, w, gn_stats=bucket)
            x = ops.carry_stats(y.reshape(1, hw, hw, c), y)
            g, b = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
            out = torch.empty(1, hw, hw, c, device="cuda", dtype=torch.float16)
            n0 = ops.gn_apply_calls
            t = graph_time(lambda: ops.group_norm(x, g, b, 32, 1e-5, silu=True, out=out))
            assert ops.gn_apply_calls > n0
            mb = hw * hw * c * 4 / 1e6
            print(f"gn_apply {hw:3d}^2 C={c}: {t:7.2f} us  ({mb:6.1f} MB moved -> {mb / t * 1e3:7.1f} GB/s)", flush=True)
    if which == "ln":
        for (rows, c) in [(4096, 320), (1024, 640), (256, 1280), (64, 1280), (258, 256), (77, 1024)]:
            x = torch.randn(rows, c, device="cuda").half()
            g, b = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
            out = torch.empty_like(x)
            t = graph_time(lambda: ops.layer_norm(x, g, b, out=out))
            print(f"ln {rows}x{c}: {t:7.2f} us", flush=True)
    if which == "attn":
        for (tq, tk, heads) in [(4096, 4096, 5), (1024, 1024, 10), (256, 256, 20), (64, 64, 20), (4096, 77, 5),
                                (1024, 77, 10), (256, 77, 20), (64, 77, 20)]:
            c = heads * 64
            q = torch.randn(tq, c, device="cuda").half()
            k = torch.randn(tk, c, device="cuda").half()
            v = torch.randn(tk, c, device="cuda").half()
            out = torch.empty(tq, c, device="cuda", dtype=torch.float16)
            t = graph_time(lambda: ops.attention(q, k, v, 1, heads, tq, tk, 0.125, out=out))
            fl = 4.0 * heads * tq * tk * 64
            print(f"attn Tq={tq} Tk={tk} h={heads}: {t:7.2f} us  {fl / t / 1e6:7.1f} TF/s", flush=True)
    if which == "gemm":
        shapes = [(4096, 320, 320, "bias_res"), (4096, 960, 320, "plain"), (4096, 2560, 320, "geglu"),
                  (4096, 320, 1280, "bias_res"), (1024, 640, 640, "bias_res"), (1024, 1920, 640, "plain"),
                  (1024, 5120, 640, "geglu"), (1024, 640, 2560, "bias_res"), (256, 1280, 1280, "bias_res"),
                  (256, 3840, 1280, "plain"), (256, 10240, 1280, "geglu"), (256, 1280, 5120, "bias_res"),
                  (64, 1280, 1280, "bias_res"), (64, 3840, 1280, "plain"), (64, 10240, 1280, "geglu"),
                  (64, 1280, 5120, "bias_res"), (77, 640, 1024, "plain"), (77, 2560, 1024, "plain"),
                  (4096, 4096, 512, "fp32out"), (4096, 512, 4096, "bias"), (4096, 512, 512, "bias")]
        for M, N, K, mode in shapes:
            a = torch.randn(M, K, device="cuda").half()
            w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
            b = torch.randn(N, device="cuda")
            kw = {}
            if mode in ("bias", "bias_res"):
                kw["bias"] = b
            if mode == "bias_res":
                kw["residual"] = torch.randn(M, N, device="cuda").half()
            if mode == "geglu":
                w, b2 = pack_geglu_weight(w, b)
                kw.update(bias=b2, geglu=True)
            if mode == "fp32out":
                kw["out_fp32"] = True
            out = ops.linear(a, w, **kw)
            t = graph_time(lambda: ops.linear(a, w, out=out, **kw))
            cfg = ops.last_gemm_config() + (("P",) if ops.lib.gn_last_gemm_pair(ops.h) else ())
            tb = graph_time(lambda: torch.matmul(a, w.t()))
            fl = 2.0 * M * N * K
            extra = ""
            if "--occ" in sys.argv:
                for occ in (1, 2):
                    ops.lib.gn_set_gemm_occupancy(ops.h, occ)
                    t2 = graph_time(lambda: ops.linear(a, w, out=out, **kw))
                    extra += f" | occ{occ} {t2:6.2f} {ops.last_gemm_config()}"
                ops.lib.gn_set_gemm_occupancy(ops.h, 0)
            if "--pair" in sys.argv:
                for pm in (0, 2):
                    ops.lib.gn_set_gemm_pair(ops.h, pm)
                    ops.linear(a, w, out=out, **kw)   # re-tune among the allowed candidates
                    t2 = graph_time(lambda: ops.linear(a, w, out=out, **kw))
                    extra += f" | pair{pm} {t2:6.2f} {ops.last_gemm_config()}"
                ops.lib.gn_set_gemm_pair(ops.h, 1)
            print(f"{M:5d} {N:6d} {K:5d} {mode:>9} | {t:7.2f} us {fl / t / 1e6:7.1f} TF/s {str(cfg):>22} | cuBLAS {tb:7.2f} us{extra}", flush=True)
    if which == "conv":
        import torch.nn.functional as F
        torch.backends.cudnn.benchmark = True
        shapes = [(64, 320, 320, 1), (64, 640, 320, 1), (64, 960, 320, 1), (32, 640, 640, 1), (32, 1280, 640, 1),
                  (32, 1920, 640, 1), (16, 1280, 1280, 1), (16, 2560, 1280, 1), (8, 1280, 1280, 1), (8, 2560, 1280, 1),
                  (64, 320, 320, 2), (32, 640, 640, 2), (16, 1280, 1280, 2), (64, 512, 512, 1), (128, 512, 512, 1),
                  (256, 512, 512, 1), (256, 256, 256, 1), (512, 256, 256, 1), (512, 128, 128, 1), (512, 128, 3, 1)]
        for H, Cin, Cout, s in shapes:
            x = torch.randn(1, H, H, Cin, device="cuda").half()
            w = (torch.randn(Cout, Cin, 3, 3) * (Cin * 9) ** -0.5).half()
            wp = pack_conv_weight(w).cuda()
            bias = torch.randn(Cout, device="cuda")
            out = ops.conv2d(x, wp, Cout, stride=s, bias=bias)
            t = graph_time(lambda: ops.conv2d(x, wp, Cout, stride=s, bias=bias, out=out), n=10)
            cfg = ops.last_gemm_config() + (("P",) if ops.lib.gn_last_gemm_pair(ops.h) else ())
            xc = x.permute(0, 3, 1, 2)
            wc = w.cuda().to(memory_format=torch.channels_last)
            bh = bias.half()
            tb = graph_time(lambda: F.conv2d(xc, wc, bh, stride=s, padding=1), n=10)
            fl = 2.0 * (H // s) ** 2 * Cout * Cin * 9
            extra = ""
            if "--occ" in sys.argv:
                for occ in (1, 2):
                    ops.lib.gn_set_gemm_occupancy(ops.h, occ)
                    t2 = graph_time(lambda: ops.conv2d(x, wp, Cout, stride=s, bias=bias, out=out), n=10)
                    extra += f" | occ{occ} {t2:6.2f} {ops.last_gemm_config()}"
                ops.lib.gn_set_gemm_occupancy(ops.h, 0)
            if "--pair" in sys.argv:
                for pm in (0, 2):
                    ops.lib.gn_set_gemm_pair(ops.h, pm)
                    ops.conv2d(x, wp, Cout, stride=s, bias=bias, out=out)   # re-tune among the allowed candidates
                    t2 = graph_time(lambda: ops.conv2d(x, wp, Cout, stride=s, bias=bias, out=out), n=10)
                    extra += f" | pair{pm} {t2:6.2f} {ops.last_gemm_config()}"
                ops.lib.gn_set_gemm_pair(ops.h, 1)
            print(f"conv {H:3d}^2 {Cin:4d}->{Cout:4d} s{s} | {t:7.2f} us {fl / t / 1e6:7.1f} TF/s {str(cfg):>22} | cuDNN {tb:7.2f} us{extra}", flush=True)


if __name__ == "__main__":
    main()
