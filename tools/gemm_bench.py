"""Micro-benchmark of gn_linear / gn_conv2d on the shapes of the agent step (L2-warm, CUDA events, 30 reps), next to
cuBLAS / cuDNN through torch as a yardstick.  Usage: python tools/gemm_bench.py [linear|conv|all] [--sweep]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genima_b200.ops import Ops  # noqa: E402
from genima_b200.packing import pack_conv_weight, pack_geglu_weight  # noqa: E402


def timeit(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


LINEAR = [  # (M, N, K, mode)
    (4096, 320, 320, "bias_res"), (4096, 960, 320, "plain"), (4096, 2560, 320, "geglu"), (4096, 320, 1280, "bias_res"),
    (1024, 640, 640, "bias_res"), (1024, 1920, 640, "plain"), (1024, 5120, 640, "geglu"), (1024, 640, 2560, "bias_res"),
    (256, 1280, 1280, "bias_res"), (256, 3840, 1280, "plain"), (256, 10240, 1280, "geglu"), (256, 1280, 5120, "bias_res"),
    (64, 1280, 1280, "bias_res"), (64, 3840, 1280, "plain"), (64, 10240, 1280, "geglu"), (64, 1280, 5120, "bias_res"),
    (77, 640, 1024, "plain"), (4096, 4096, 512, "fp32out"), (4096, 512, 4096, "bias"), (4096, 512, 512, "bias"),
    (1, 1280, 320, "bias"), (258, 768, 256, "plain"), (20, 2048, 256, "bias"),
]
CONV = [  # (B, H, W, Cin, Cout, stride)
    (1, 64, 64, 320, 320, 1), (1, 64, 64, 640, 320, 1), (1, 64, 64, 960, 320, 1), (1, 32, 32, 640, 640, 1),
    (1, 32, 32, 1280, 640, 1), (1, 16, 16, 1280, 1280, 1), (1, 16, 16, 2560, 1280, 1), (1, 8, 8, 1280, 1280, 1),
    (1, 8, 8, 2560, 1280, 1), (1, 64, 64, 320, 320, 2), (1, 32, 32, 640, 640, 2), (1, 16, 16, 1280, 1280, 2),
    (1, 64, 64, 512, 512, 1), (1, 128, 128, 512, 512, 1), (1, 256, 256, 512, 512, 1), (1, 256, 256, 512, 256, 1),
    (1, 256, 256, 256, 256, 1), (1, 512, 512, 256, 256, 1), (1, 512, 512, 256, 128, 1), (1, 512, 512, 128, 128, 1),
    (1, 512, 512, 128, 3, 1), (4, 64, 64, 64, 64, 1), (4, 32, 32, 128, 128, 1), (4, 16, 16, 256, 256, 1),
    (4, 8, 8, 512, 512, 1),
]


def bench_linear(ops, sweep):
    print(f"{'M':>6} {'N':>6} {'K':>6} {'mode':>9} | {'ours us':>9} {'TF/s':>7} {'cfg(bn,split,stg,ctas)':>24} | {'cuBLAS us':>9} {'TF/s':>7}")
    for M, N, K, mode in LINEAR:
        a = torch.randn(M, K, device="cuda").half()
        w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
        b = torch.randn(N, device="cuda")
        res = torch.randn(M, N // 2 if mode == "geglu" else N, device="cuda").half()
        kw = {}
        if mode in ("bias", "bias_res"):
            kw["bias"] = b
        if mode == "bias_res":
            kw["residual"] = res
        if mode == "geglu":
            w, b2 = pack_geglu_weight(w, b)
            kw.update(bias=b2, geglu=True)
        if mode == "fp32out":
            kw["out_fp32"] = True
        out = ops.linear(a, w, **kw)
        t = timeit(lambda: ops.linear(a, w, out=out, **kw))
        cfg = ops.last_gemm_config()
        tb = timeit(lambda: torch.matmul(a, w.t()))
        fl = 2.0 * M * N * K
        print(f"{M:6d} {N:6d} {K:6d} {mode:>9} | {t:9.1f} {fl / t / 1e6:7.1f} {str(cfg):>24} | {tb:9.1f} {fl / tb / 1e6:7.1f}", flush=True)
        if sweep:
            for bn in (64, 128, 256):
                if mode == "geglu" and bn % 128:
                    continue
                for sp in (1, 2, 4):
                    try:
                        ops.set_gemm_tuning(bn, sp)
                        t2 = timeit(lambda: ops.linear(a, w, out=out, **kw), reps=10)
                        print(f"{'':31} bn={bn:3d} split={sp}: {t2:8.1f} us {ops.last_gemm_config()}")
                    except Exception as e:
                        print(f"{'':31} bn={bn} split={sp}: {e}")
                    finally:
                        ops.set_gemm_tuning(0, 0)


def bench_conv(ops, sweep):
    print(f"{'B':>2} {'H':>4} {'Cin':>5} {'Cout':>5} {'s':>2} | {'ours us':>9} {'TF/s':>7} {'cfg':>24} | {'cuDNN us':>9} {'TF/s':>7}")
    for B, H, Wd, Cin, Cout, s in CONV:
        x = torch.randn(B, H, Wd, Cin, device="cuda").half()
        w = (torch.randn(Cout, Cin, 3, 3) * (Cin * 9) ** -0.5).half()
        wp = pack_conv_weight(w).cuda()
        bias = torch.randn(Cout, device="cuda")
        out = ops.conv2d(x, wp, Cout, stride=s, bias=bias)
        t = timeit(lambda: ops.conv2d(x, wp, Cout, stride=s, bias=bias, out=out), reps=20)
        cfg = ops.last_gemm_config()
        xc = x.permute(0, 3, 1, 2)  # channels_last view
        wc = w.cuda().to(memory_format=torch.channels_last)
        bh = bias.half()
        tb = timeit(lambda: F.conv2d(xc, wc, bh, stride=s, padding=1), reps=20)
        fl = 2.0 * B * (H // s) * (Wd // s) * Cout * Cin * 9
        print(f"{B:2d} {H:4d} {Cin:5d} {Cout:5d} {s:2d} | {t:9.1f} {fl / t / 1e6:7.1f} {str(cfg):>24} | {tb:9.1f} {fl / tb / 1e6:7.1f}", flush=True)
        if sweep:
            for bn in (64, 128, 256):
                for sp in (1, 2, 4, 8):
                    try:
                        ops.set_gemm_tuning(bn, sp)
                        t2 = timeit(lambda: ops.conv2d(x, wp, Cout, stride=s, bias=bias, out=out), reps=10)
                        print(f"{'':24} bn={bn:3d} split={sp}: {t2:8.1f} us {ops.last_gemm_config()}")
                    except Exception as e:
                        print(f"{'':24} bn={bn} split={sp}: {str(e)[:80]}")
                    finally:
                        ops.set_gemm_tuning(0, 0)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    sweep = "--sweep" in sys.argv
    ops = Ops(0, workspace_mb=256)
    torch.backends.cudnn.benchmark = True
    if which in ("linear", "all"):
        bench_linear(ops, sweep)
    if which in ("conv", "all"):
        bench_conv(ops, sweep)
