"""Correctness sweep of gn_linear / gn_conv2d over forced tile configurations (block_n x splits x CTAs/SM) on the
full-size shapes of the agent step, against torch fp32 on the GPU.  Prints only failures + a summary."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genima_b200.ops import Ops  # noqa: E402
from genima_b200.packing import pack_conv_weight, pack_geglu_weight  # noqa: E402

ops = Ops(0)
bad = total = 0


def check(name, out, ref):
    global bad, total
    total += 1
    err = float((out.float() - ref).abs().max() / ref.abs().max())
    if not (err < 3e-3):
        bad += 1
        print(f"BAD {name}: err {err:.3e} cfg {ops.last_gemm_config()}", flush=True)


LIN = [(4096, 320, 320, "bias_res"), (4096, 2560, 320, "geglu"), (4096, 320, 1280, "bias_res"), (1024, 640, 2560, "bias_res"),
       (1024, 5120, 640, "geglu"), (256, 1280, 5120, "bias_res"), (256, 10240, 1280, "geglu"), (64, 1280, 1280, "bias_res"),
       (64, 3840, 1280, "plain"), (64, 10240, 1280, "geglu"), (64, 1280, 5120, "bias_res"), (77, 2560, 1024, "plain"),
       (1, 1280, 320, "bias"), (258, 768, 256, "plain")]
for M, N, K, mode in LIN:
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    b = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").half()
    acc = a.float() @ w.float().t()
    kw = {}
    if mode == "geglu":
        ref = (acc[:, :N // 2] + b[:N // 2]) * F.gelu(acc[:, N // 2:] + b[N // 2:])
        wp, bp = pack_geglu_weight(w, b)
        kw = dict(bias=bp, geglu=True)
        w_use = wp
    else:
        ref = acc.clone()
        w_use = w
        if mode in ("bias", "bias_res"):
            ref += b
            kw["bias"] = b
        if mode == "bias_res":
            ref += res.float()
            kw["residual"] = res
    for bn in (0, 256, 192, 160, 128, 96, 64, 48, 32):
        if mode == "geglu" and bn % 128:
            continue
        for sp in (0, 1, 2, 3, 4, 5, 6, 7, 8):
            for occ in (1, 2):
                ops.set_gemm_tuning(bn, sp)
                ops.lib.gn_set_gemm_occupancy(ops.h, occ)
                try:
                    out = ops.linear(a, w_use, **kw)
                    check(f"linear {M}x{N}x{K} {mode} bn={bn} sp={sp} occ={occ}", out, ref)
                except Exception as e:
                    print("EXC", M, N, K, mode, bn, sp, occ, str(e)[:100])
ops.set_gemm_tuning(0, 0)
CONV = [(64, 320, 320, 1), (64, 960, 320, 1), (32, 1920, 640, 1), (16, 2560, 1280, 1), (8, 1280, 1280, 1), (8, 2560, 1280, 1),
        (64, 320, 320, 2), (16, 1280, 1280, 2), (32, 640, 640, 2)]
for H, Cin, Cout, s in CONV:
    x = torch.randn(1, H, H, Cin, device="cuda").half()
    w = (torch.randn(Cout, Cin, 3, 3) * (Cin * 9) ** -0.5).half()
    wp = pack_conv_weight(w).cuda()
    bias = torch.randn(Cout, device="cuda")
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.cuda().float(), bias, stride=s, padding=1).permute(0, 2, 3, 1)
    for bn in (0, 256, 224, 160, 128, 96, 80, 64):
        for sp in (0, 1, 2, 3, 4, 5, 6, 7, 8):
            for occ in (1, 2):
                ops.set_gemm_tuning(bn, sp)
                ops.lib.gn_set_gemm_occupancy(ops.h, occ)
                try:
                    out = ops.conv2d(x, wp, Cout, stride=s, bias=bias)
                    check(f"conv {H}^2 {Cin}->{Cout} s{s} bn={bn} sp={sp} occ={occ}", out, ref)
                except Exception as e:
                    print("EXC conv", H, Cin, Cout, s, bn, sp, occ, str(e)[:100])
ops.set_gemm_tuning(0, 0)
ops.lib.gn_set_gemm_occupancy(ops.h, 0)
print(f"checked {total} configurations, {bad} bad")
