"""Thin tensor-level wrappers over the C ABI (genima_b200._cabi).

torch is used for device memory and streams only; every arithmetic op below is one call into libgenima_b200.so on
`torch.cuda.current_stream()`, so a chain of calls can be captured with `torch.cuda.graph`.

Layout convention: activations are fp16, channels-last.  An image is a contiguous [B, H, W, C] tensor, which is also
the row-major [B*H*W, C] matrix the GEMM kernels see.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import torch

from . import _cabi
from ._cabi import ACT_GELU, ACT_NONE, ACT_QUICKGELU, ACT_RELU, ACT_SILU, GnEpilogue  # noqa: F401

_ACTS = {None: ACT_NONE, "none": ACT_NONE, "silu": ACT_SILU, "gelu": ACT_GELU, "relu": ACT_RELU,
         "quick_gelu": ACT_QUICKGELU}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f16(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float16 or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA float16 tensor, got {t.dtype} on {t.device}")
    return t


def _f32(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise TypeError(f"{name}: expected a contiguous CUDA float32 tensor")
    return t


class RowStats:
    """Per-row (sum, sumsq) partials of a GEMM output: float2 [M][parts] (parts is set by the producing call)."""

    def __init__(self, M: int, N: int, device):
        self.capacity = max(4, (N + 15) // 16 + 2)
        self.buf = torch.empty(M, self.capacity, 2, dtype=torch.float32, device=device)
        self.parts = 0


class GNStats:
    """GroupNorm statistics of one fp16 tensor, accumulated by the GEMM epilogue that produced it: uint64
    [B][C / bucket][2] fixed-point (sum, sumsq) per (image, bucket of channels) in the owning Ops' statistics arena."""

    __slots__ = ("buf", "bucket", "generation", "owner")

    def __init__(self, buf: torch.Tensor, bucket: int, generation: int, owner=None):
        self.buf, self.bucket, self.generation, self.owner = buf, bucket, generation, owner


def gn_bucket_for(channels: Sequence[int], groups: int) -> int:
    """Channels per statistics bucket such that every GroupNorm over any of `channels` (or a concat of two of them) with
    `groups` groups covers whole buckets: gcd(channels) / groups; 0 (fusion off) when that is not an even integer."""
    import math

    g = 0
    for c in channels:
        g = math.gcd(g, int(c))
    if g == 0 or g % groups:
        return 0
    b = g // groups
    return b if b >= 2 and b % 2 == 0 else 0


class Ops:
    """One instance per device; owns the gn_handle and the scratch workspace (L2-flush buffer of the autotuner)."""

    GN_ARENA_WORDS = 1 << 19  # 4 MiB of uint64 statistics accumulators: far more than one agent step needs

    def __init__(self, device: int = 0, workspace_mb: int = 160, autotune: bool = True):
        if not torch.cuda.is_available():
            raise _cabi.GenimaB200Error("CUDA is not available: genima_b200 has no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.handle = _cabi.Handle(device)
        self.lib = self.handle.lib
        self.h = self.handle.ptr
        self.workspace = torch.empty(workspace_mb * 1024 * 1024 // 4, dtype=torch.float32, device=self.device)
        self.handle.check(self.lib.gn_set_workspace(self.h, self.workspace.data_ptr(), self.workspace.numel() * 4),
                          "gn_set_workspace")
        # tile configurations are measured once per problem shape (first eager call) and cached in the handle
        self.handle.check(self.lib.gn_set_autotune(self.h, 1 if autotune else 0), "gn_set_autotune")
        self.autotune = bool(autotune)
        if os.environ.get("GENIMA_B200_PDL", "1") != "1":   # A/B: 0 = ordinary launches, 2 = PDL without weight prefetch
            self.handle.check(self.lib.gn_set_pdl(self.h, int(os.environ["GENIMA_B200_PDL"])), "gn_set_pdl")
        if os.environ.get("GENIMA_B200_STAGED", "1") == "0":   # A/B switch for the TMA-stored GEMM epilogue
            self.handle.check(self.lib.gn_set_staged_epilogue(self.h, 0), "gn_set_staged_epilogue")
        if os.environ.get("GENIMA_B200_PAIR", "1") != "1":   # A/B: 0 = no CTA pairs, 2 = pairs wherever possible
            self.handle.check(self.lib.gn_set_gemm_pair(self.h, int(os.environ["GENIMA_B200_PAIR"])), "gn_set_gemm_pair")
        if os.environ.get("GENIMA_B200_OCC"):   # A/B: operand-ring sizing for 1 or 2 resident CTAs per SM everywhere
            self.handle.check(self.lib.gn_set_gemm_occupancy(self.h, int(os.environ["GENIMA_B200_OCC"])),
                              "gn_set_gemm_occupancy")
        if os.environ.get("GENIMA_B200_ATTN_SPLIT", "1") != "1":   # A/B: 0 = no KV split, 2 = split whenever possible
            self.set_attention_kv_split(int(os.environ["GENIMA_B200_ATTN_SPLIT"]))
        # GroupNorm statistics fused into the producing GEMM epilogues (A/B switch: GENIMA_B200_GNFUSE=0)
        self.gn_fuse = os.environ.get("GENIMA_B200_GNFUSE", "1") != "0"
        # cross-attention query projection inside the attention kernel (A/B: GENIMA_B200_QPROJ=0)
        self.fuse_qproj = os.environ.get("GENIMA_B200_QPROJ", "1") != "0"
        # nearest-upsample folded into the following convolution (four 2x2 phase kernels; A/B: GENIMA_B200_UPFOLD=0)
        self.fold_upsample = os.environ.get("GENIMA_B200_UPFOLD", "1") != "0"
        self._gn_arena = torch.zeros(self.GN_ARENA_WORDS, dtype=torch.int64, device=self.device)
        self._gn_used = 0
        self._gn_generation = 0
        self.gn_apply_calls = 0   # group_norm calls served by gn_group_norm_apply (fused statistics)

    # ------------------------------------------------------------------------------------------------ helpers
    @staticmethod
    def _stream() -> int:
        return torch.cuda.current_stream().cuda_stream

    def _epilogue(self, M: int, N: int, bias=None, scale=None, rowvec=None, rows_per_batch: int = 0, residual=None,
                  act_pre=None, act_post=None, alpha: float = 1.0, beta: float = 1.0, geglu: bool = False,
                  out_fp32: bool = False, ln=None, row_stats=None, gn_stats: Optional[GNStats] = None,
                  w_dynamic: bool = False) -> GnEpilogue:
        e = GnEpilogue()
        e.scale = _ptr(_f32(scale, "scale"))
        e.bias = _ptr(_f32(bias, "bias"))
        e.rowvec = _ptr(_f32(rowvec, "rowvec"))
        if bias is not None and bias.numel() != N:
            raise ValueError(f"bias has {bias.numel()} elements, expected {N}")
        if scale is not None and scale.numel() != N:
            raise ValueError(f"scale has {scale.numel()} elements, expected {N}")
        if rowvec is not None:
            if rowvec.shape[-1] != N or rows_per_batch <= 0:
                raise ValueError("rowvec needs shape [B, N] and rows_per_batch > 0")
        n_out = N // 2 if geglu else N
        if residual is not None:
            _f16(residual, "residual")
            r2 = residual.reshape(-1, residual.shape[-1])
            if r2.shape[0] != M or r2.shape[1] < n_out or r2.stride(1) != 1:
                raise ValueError(f"residual shape {tuple(residual.shape)} does not match output [{M}, {n_out}]")
            e.residual = r2.data_ptr()
            e.ldr = r2.stride(0)
        e.rows_per_batch = rows_per_batch
        e.act_pre = _ACTS[act_pre]
        e.act_post = _ACTS[act_post]
        e.alpha = alpha
        e.beta = beta
        e.geglu = 1 if geglu else 0
        e.out_fp32 = 1 if out_fp32 else 0
        if ln is not None:
            # ln = (RowStats of the A operand, colsum [N] fp32, eps): LayerNorm folded into this GEMM
            stats, colsum, eps = ln
            if stats.buf.shape[0] != M or colsum.numel() != N:
                raise ValueError("folded LayerNorm: row statistics / colsum do not match the GEMM")
            e.ln_stats = stats.buf.data_ptr()
            e.ln_colsum = _f32(colsum, "ln colsum").data_ptr()
            e.ln_parts = stats.parts
            e.ln_eps = float(eps)
        if row_stats is not None:
            e.rowstats_out = row_stats.buf.data_ptr()
            e.rowstats_capacity = row_stats.capacity
        if gn_stats is not None:
            e.gnstats_out = gn_stats.buf.data_ptr()
            e.gn_bucket = gn_stats.bucket
        e.w_dynamic = 1 if w_dynamic else 0   # `w` is an activation of this stream, not a constant weight matrix
        return e

    # ------------------------------------------------------------------------------------------------ GroupNorm statistics
    def gn_stats_reset(self) -> None:
        """Zero the statistics arena on the current stream and start allocating from its beginning again.  Call once at
        the top of every forward pass (it is one memset node in a captured graph); statistics handed out before the
        reset are no longer used (group_norm falls back to its own reduction for them)."""
        self._gn_arena.zero_()
        self._gn_used = 0
        self._gn_generation += 1

    def _gn_stats_alloc(self, images: int, n_out: int, bucket: int) -> GNStats:
        words = images * (n_out // bucket) * 2
        if self._gn_used + words > self.GN_ARENA_WORDS:
            # arena exhausted (nobody called gn_stats_reset): slow path, a private zeroed buffer that never expires
            return GNStats(torch.zeros(words, dtype=torch.int64, device=self.device), bucket, 0, None)
        buf = self._gn_arena[self._gn_used:self._gn_used + words]
        self._gn_used += (words + 1) // 2 * 2
        return GNStats(buf, bucket, self._gn_generation, self)

    @staticmethod
    def _gn_rows_ok(rows_per_image: int) -> bool:
        return (rows_per_image >= 128 and rows_per_image % 128 == 0) or rows_per_image in (16, 32, 64)

    def carry_stats(self, view: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
        """`view` is a reshape of `src`: keep the GroupNorm statistics attached."""
        st = getattr(src, "gn_stats", None)
        if st is not None:
            view.gn_stats = st
        return view

    def set_attention_kv_split(self, mode: int) -> None:
        """0: never split the keys of gn_attention across a 2-CTA cluster, 1: when it pays (default), 2: whenever possible."""
        self.handle.check(self.lib.gn_set_attention_kv_split(self.h, int(mode)), "gn_set_attention_kv_split")

    def set_gemm_tuning(self, block_n: int = 0, splits: int = 0) -> None:
        self.handle.check(self.lib.gn_set_gemm_tuning(self.h, block_n, splits), "gn_set_gemm_tuning")

    def set_autotune(self, enable: bool) -> None:
        self.handle.check(self.lib.gn_set_autotune(self.h, 1 if enable else 0), "gn_set_autotune")
        self.autotune = bool(enable)

    def tune_cache_export(self) -> bytes:
        """Measured tile configurations of this handle as text (gn_tune_cache_export)."""
        n = int(self.lib.gn_tune_cache_export(self.h, None, 0))
        if n < 0:
            self.handle.check(n, "gn_tune_cache_export")
        buf = C.create_string_buffer(max(n, 1))
        self.lib.gn_tune_cache_export(self.h, buf, n)
        return buf.raw[:n]

    def tune_cache_import(self, text: bytes, replace: bool = False) -> None:
        """Adopt tile configurations measured elsewhere (another handle / rank 0): identical launches, identical
        summation order, bit-identical results."""
        self.handle.check(self.lib.gn_tune_cache_import(self.h, text, len(text), 1 if replace else 0),
                          "gn_tune_cache_import")

    def last_gemm_config(self) -> Tuple[int, int, int, int]:
        out = (C.c_int32 * 4)()
        self.lib.gn_get_last_gemm_config(self.h, out)
        return tuple(out)

    def launch_count(self) -> int:
        return self.handle.launch_count()

    def num_sms(self) -> int:
        return torch.cuda.get_device_properties(self.device).multi_processor_count

    def set_gn_max_ctas(self, n: int) -> None:
        """Cap the grid of the following gn_group_norm launches (0 = one CTA per SM)."""
        self.handle.check(self.lib.gn_set_gn_max_ctas(self.h, int(n)), "gn_set_gn_max_ctas")

    PROF_CLASSES = ("linear", "conv", "attention", "attention_small", "norm", "elementwise")

    def profile_begin(self) -> None:
        """Start per-call CUDA-event timing inside the library (eager launches only, not under graph capture)."""
        self.handle.check(self.lib.gn_profile_begin(self.h), "gn_profile_begin")

    def profile_end(self) -> dict:
        """-> {class: dict(ms=, calls=, flops=, bytes=)} summed over the calls since profile_begin (synchronises)."""
        n = len(self.PROF_CLASSES)
        ms, calls, flops, byts = (C.c_double * n)(), (C.c_int64 * n)(), (C.c_double * n)(), (C.c_double * n)()
        self.handle.check(self.lib.gn_profile_end(self.h, ms, calls, flops, byts), "gn_profile_end")
        return {name: dict(ms=ms[i], calls=int(calls[i]), flops=flops[i], bytes=byts[i])
                for i, name in enumerate(self.PROF_CLASSES)}

    # ------------------------------------------------------------------------------------------------ contractions
    def new_row_stats(self, M: int, N: int) -> "RowStats":
        """Buffer for the per-row (sum, sumsq) partials a GEMM with N output columns writes (row_stats=...)."""
        return RowStats(M, N, self.device)

    def linear(self, a: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None, **epi) -> torch.Tensor:
        """out[..., N'] = epilogue(a[..., K] @ w[N, K]^T); N' = N/2 with geglu=True.
        row_stats=RowStats: also write the row statistics of the fp16 output (for a LayerNorm folded into the consumer);
        ln=(RowStats, colsum, eps): `a` is un-normalised, `w` is pre-multiplied by gamma (packing.fold_layer_norm)."""
        _f16(a, "a")
        _f16(w, "w")
        K = a.shape[-1]
        a2 = a.reshape(-1, K)
        if a2.stride(1) != 1:
            a2 = a2.contiguous()
        M = a2.shape[0]
        N = w.shape[0]
        if w.shape[1] != K or not w.is_contiguous():
            raise ValueError(f"w must be contiguous [N, K={K}], got {tuple(w.shape)}")
        n_out = N // 2 if epi.get("geglu") else N
        out_dtype = torch.float32 if epi.get("out_fp32") else torch.float16
        if out is None:
            out = torch.empty(*a.shape[:-1], n_out, dtype=out_dtype, device=a.device)
        o2 = out.reshape(-1, out.shape[-1])
        if o2.shape[0] != M or o2.shape[1] < n_out or o2.stride(1) != 1 or out.dtype != out_dtype:
            raise ValueError("bad `out` tensor for linear")
        st = self._gn_request(epi, M, n_out, epi.get("rows_per_batch") or M, out_dtype)
        e = self._epilogue(M, N, **epi)
        rc = self.lib.gn_linear(self.h, a2.data_ptr(), a2.stride(0), M, K, w.data_ptr(), N, o2.data_ptr(),
                                o2.stride(0), C.byref(e), self._stream())
        self.handle.check(rc, "gn_linear")
        if epi.get("row_stats") is not None:
            epi["row_stats"].parts = int(self.lib.gn_get_last_rowstats_parts(self.h))
        if st is not None:
            out.gn_stats = st
        return out

    def _gn_request(self, epi: dict, M: int, n_out: int, rows_per_image: int, out_dtype) -> Optional[GNStats]:
        """Turn the `gn_stats=<bucket>` keyword of linear / conv2d into an allocated GNStats (or drop it when the layout
        is not supported by the fused statistics pass: the consumer then reduces on its own)."""
        bucket = epi.pop("gn_stats", 0)
        if not bucket or not self.gn_fuse:
            return None
        if (out_dtype != torch.float16 or bucket % 2 or n_out % bucket or n_out % 2 or M % rows_per_image
                or not self._gn_rows_ok(rows_per_image)):
            return None
        st = self._gn_stats_alloc(M // rows_per_image, n_out, bucket)
        epi["gn_stats"] = st
        return st

    def conv2d(self, x: torch.Tensor, w: torch.Tensor, cout: int, ksize: int = 3, stride: int = 1, pad: int = 1,
               extras: Sequence[torch.Tensor] = (), out: Optional[torch.Tensor] = None, **epi) -> torch.Tensor:
        """NHWC implicit-GEMM convolution.  `w` is packed by packing.pack_conv_weight ([Cout, K_total])."""
        _f16(x, "x")
        _f16(w, "w")
        if x.dim() != 4 or not x.is_contiguous():
            raise ValueError("x must be a contiguous [B, H, W, C] tensor")
        B, H, W, Cin = x.shape
        Ho = (H + 2 * pad - ksize) // stride + 1
        Wo = (W + 2 * pad - ksize) // stride + 1
        cp = (Cin + 63) // 64 * 64
        ktot = ksize * ksize * cp
        ex = [None, None]
        exc = [0, 0]
        for i, t in enumerate(extras):
            _f16(t, "extra")
            if t.dim() != 4 or not t.is_contiguous() or t.shape[:3] != (B, Ho, Wo):
                raise ValueError("extra source must be contiguous [B, Ho, Wo, C]")
            ex[i] = t.data_ptr()
            exc[i] = t.shape[3]
            ktot += (t.shape[3] + 63) // 64 * 64
        if tuple(w.shape) != (cout, ktot) or not w.is_contiguous():
            raise ValueError(f"packed conv weight must be [{cout}, {ktot}], got {tuple(w.shape)}")
        out_dtype = torch.float32 if epi.get("out_fp32") else torch.float16
        if out is None:
            out = torch.empty(B, Ho, Wo, cout, dtype=out_dtype, device=x.device)
        if out.shape[:3] != (B, Ho, Wo) or out.shape[3] < cout or out.stride(3) != 1 or out.dtype != out_dtype:
            raise ValueError("bad `out` tensor for conv2d")
        if out.stride(2) * Wo != out.stride(1) or out.stride(1) * Ho != out.stride(0):
            raise ValueError("`out` must be pixel-contiguous")
        M = B * Ho * Wo
        epi.setdefault("rows_per_batch", Ho * Wo)
        # rows of one image inside a 128-pixel tile (bw x bh box of powers of two, see gn_conv2d)
        bw = min(128, 1 << max(0, (Wo - 1).bit_length()))
        bh = min(128 // bw, 1 << max(0, (Ho - 1).bit_length()))
        st = self._gn_request(epi, B * 128, cout, 128, out_dtype) if bw * bh >= 16 and epi.get("gn_stats") else None
        if st is None:
            epi.pop("gn_stats", None)
        e = self._epilogue(M, cout, **epi)
        rc = self.lib.gn_conv2d(self.h, x.data_ptr(), B, H, W, Cin, w.data_ptr(), cout, ksize, ksize, stride, pad,
                                ex[0], exc[0], ex[1], exc[1], out.data_ptr(), out.stride(2), C.byref(e),
                                self._stream())
        self.handle.check(rc, "gn_conv2d")
        if st is not None:
            out.gn_stats = st
        return out

    def conv2d_asym(self, x: torch.Tensor, w: torch.Tensor, cout: int, ksize: int = 3, stride: int = 2,
                    pads: Sequence[int] = (0, 0, 1, 1), out: Optional[torch.Tensor] = None, **epi) -> torch.Tensor:
        """Convolution with (top, left, bottom, right) zero padding (gn_conv2d_asym); the default is the VAE encoder's
        Downsample2D: F.pad(x, (0, 1, 0, 1)) + 3x3 stride-2 convolution."""
        _f16(x, "x")
        _f16(w, "w")
        if x.dim() != 4 or not x.is_contiguous():
            raise ValueError("x must be a contiguous [B, H, W, C] tensor")
        B, H, W, Cin = x.shape
        pt, pl, pb, pr = (int(v) for v in pads)
        Ho = (H + pt + pb - ksize) // stride + 1
        Wo = (W + pl + pr - ksize) // stride + 1
        cp = (Cin + 63) // 64 * 64
        if tuple(w.shape) != (cout, ksize * ksize * cp) or not w.is_contiguous():
            raise ValueError(f"packed conv weight must be [{cout}, {ksize * ksize * cp}], got {tuple(w.shape)}")
        if out is None:
            out = torch.empty(B, Ho, Wo, cout, dtype=torch.float16, device=x.device)
        if tuple(out.shape) != (B, Ho, Wo, cout) or not out.is_contiguous() or out.dtype != torch.float16:
            raise ValueError("bad `out` tensor for conv2d_asym")
        epi.setdefault("rows_per_batch", Ho * Wo)
        bw = min(128, 1 << max(0, (Wo - 1).bit_length()))
        bh = min(128 // bw, 1 << max(0, (Ho - 1).bit_length()))
        st = self._gn_request(epi, B * 128, cout, 128, torch.float16) if bw * bh >= 16 and epi.get("gn_stats") else None
        if st is None:
            epi.pop("gn_stats", None)
        e = self._epilogue(B * Ho * Wo, cout, **epi)
        rc = self.lib.gn_conv2d_asym(self.h, x.data_ptr(), B, H, W, Cin, w.data_ptr(), cout, ksize, ksize, stride,
                                     pt, pl, pb, pr, out.data_ptr(), cout, C.byref(e), self._stream())
        self.handle.check(rc, "gn_conv2d_asym")
        if st is not None:
            out.gn_stats = st
        return out

    def conv2d_up2x(self, x: torch.Tensor, w4: torch.Tensor, cout: int, out: Optional[torch.Tensor] = None,
                    **epi) -> torch.Tensor:
        """conv3x3(pad 1)(nearest x2 upsample(x)) as four 2x2 phase convolutions over x (gn_conv2d_up2x); `w4` from
        packing.pack_upsample_conv_weight.  Returns [B, 2H, 2W, cout]."""
        _f16(x, "x")
        _f16(w4, "w4")
        if x.dim() != 4 or not x.is_contiguous():
            raise ValueError("x must be a contiguous [B, H, W, C] tensor")
        B, H, W, Cin = x.shape
        cp = (Cin + 63) // 64 * 64
        if tuple(w4.shape) != (4, cout, 4 * cp) or not w4.is_contiguous():
            raise ValueError(f"packed phase weights must be [4, {cout}, {4 * cp}], got {tuple(w4.shape)}")
        if out is None:
            out = torch.empty(B, 2 * H, 2 * W, cout, dtype=torch.float16, device=x.device)
        if tuple(out.shape) != (B, 2 * H, 2 * W, cout) or not out.is_contiguous() or out.dtype != torch.float16:
            raise ValueError("bad `out` tensor for conv2d_up2x")
        epi.setdefault("rows_per_batch", H * W)
        bw = min(128, 1 << max(0, (W - 1).bit_length()))
        bh = min(128 // bw, 1 << max(0, (H - 1).bit_length()))
        st = self._gn_request(epi, B * 128, cout, 128, torch.float16) if bw * bh >= 16 and epi.get("gn_stats") else None
        if st is None:
            epi.pop("gn_stats", None)
        e = self._epilogue(B * H * W, cout, **epi)
        rc = self.lib.gn_conv2d_up2x(self.h, x.data_ptr(), B, H, W, Cin, w4.data_ptr(), cout, out.data_ptr(), cout,
                                     C.byref(e), self._stream())
        self.handle.check(rc, "gn_conv2d_up2x")
        if st is not None:
            out.gn_stats = st
        return out

    # ------------------------------------------------------------------------------------------------ attention
    def attention(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, B: int, heads: int, Tq: int, Tk: int,
                  scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """tcgen05 flash attention, head_dim 64.  q/k/v are 2D row views (last-dim stride 1) with heads*64 columns."""
        for name, t in (("q", q), ("k", k), ("v", v)):
            _f16(t, name)
            if t.dim() != 2 or t.stride(1) != 1 or t.shape[1] != heads * 64:
                raise ValueError(f"{name} must be a 2D view with {heads * 64} unit-stride columns")
        if q.shape[0] != B * Tq or k.shape[0] != B * Tk or v.shape[0] != B * Tk:
            raise ValueError("row counts do not match B*Tq / B*Tk")
        if out is None:
            out = torch.empty(B * Tq, heads * 64, dtype=torch.float16, device=q.device)
        rc = self.lib.gn_attention(self.h, q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(),
                                   v.stride(0), out.data_ptr(), out.stride(0), B, heads, Tq, Tk, float(scale),
                                   self._stream())
        self.handle.check(rc, "gn_attention")
        return out

    def attention_qproj(self, x: torch.Tensor, wq: torch.Tensor, k: torch.Tensor, v: torch.Tensor, B: int, heads: int,
                        Tq: int, Tk: int, scale: float, bias: Optional[torch.Tensor] = None, ln=None,
                        out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """softmax(scale (x Wq^T + bias) K^T) V with the query projection inside the attention kernel (cross-attention:
        K / V are cached per prompt).  x [B*Tq, C] fp16, wq [heads*64, C] fp16; ln = (RowStats of x, colsum, eps) folds a
        LayerNorm of x into the projection exactly as `linear(..., ln=)` does."""
        _f16(x, "x")
        _f16(wq, "wq")
        C = x.shape[1]
        if x.dim() != 2 or x.stride(1) != 1 or x.shape[0] != B * Tq:
            raise ValueError("x must be a [B*Tq, C] view with unit-stride columns")
        if tuple(wq.shape) != (heads * 64, C) or not wq.is_contiguous():
            raise ValueError(f"wq must be a contiguous [{heads * 64}, {C}] matrix")
        for name, t in (("k", k), ("v", v)):
            _f16(t, name)
            if t.dim() != 2 or t.stride(1) != 1 or t.shape[1] != heads * 64 or t.shape[0] != B * Tk:
                raise ValueError(f"{name} must be a [B*Tk, {heads * 64}] view with unit-stride columns")
        if out is None:
            out = torch.empty(B * Tq, heads * 64, dtype=torch.float16, device=x.device)
        st_ptr, parts, eps, cs_ptr = None, 0, 0.0, None
        if ln is not None:
            stats, colsum, eps = ln
            if stats.buf.shape[0] != B * Tq or colsum.numel() != heads * 64:
                raise ValueError("folded LayerNorm: row statistics / colsum do not match the projection")
            st_ptr, parts, cs_ptr = stats.buf.data_ptr(), stats.parts, _f32(colsum, "ln colsum").data_ptr()
        rc = self.lib.gn_attention_qproj(self.h, x.data_ptr(), x.stride(0), C, wq.data_ptr(), _ptr(_f32(bias, "bias")),
                                         st_ptr, parts, float(eps), cs_ptr, k.data_ptr(), k.stride(0), v.data_ptr(),
                                         v.stride(0), out.data_ptr(), out.stride(0), B, heads, Tq, Tk, float(scale),
                                         self._stream())
        self.handle.check(rc, "gn_attention_qproj")
        return out

    def attention_small(self, q, k, v, B: int, heads: int, head_dim: int, Tq: int, Tk: int, scale: float,
                        causal: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        for name, t in (("q", q), ("k", k), ("v", v)):
            _f16(t, name)
            if t.dim() != 2 or t.stride(1) != 1 or t.shape[1] != heads * head_dim:
                raise ValueError(f"{name} must be a 2D view with {heads * head_dim} unit-stride columns")
        if out is None:
            out = torch.empty(B * Tq, heads * head_dim, dtype=torch.float16, device=q.device)
        rc = self.lib.gn_attention_small(self.h, q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(),
                                         v.stride(0), out.data_ptr(), out.stride(0), B, heads, head_dim, Tq, Tk,
                                         float(scale), 1 if causal else 0, self._stream())
        self.handle.check(rc, "gn_attention_small")
        return out

    # ------------------------------------------------------------------------------------------------ normalisation
    def group_norm(self, x0: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int = 32,
                   eps: float = 1e-5, silu: bool = False, x1: Optional[torch.Tensor] = None,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """GroupNorm(+SiLU) over NHWC; with x1 the input is concat([x0, x1], channel) without materialising it."""
        _f16(x0, "x0")
        if not x0.is_contiguous():
            raise ValueError("x0 must be contiguous NHWC")
        B = x0.shape[0]
        C0 = x0.shape[-1]
        HW = x0.numel() // (B * C0)
        C1 = 0
        if x1 is not None:
            _f16(x1, "x1")
            if not x1.is_contiguous() or x1.shape[:-1] != x0.shape[:-1]:
                raise ValueError("x1 must be contiguous and match x0's spatial shape")
            C1 = x1.shape[-1]
        _f32(gamma, "gamma")
        _f32(beta, "beta")
        if gamma.numel() != C0 + C1 or beta.numel() != C0 + C1:
            raise ValueError("gamma/beta size mismatch")
        if out is None:
            out = torch.empty(*x0.shape[:-1], C0 + C1, dtype=torch.float16, device=x0.device)
        st0 = getattr(x0, "gn_stats", None)
        st1 = getattr(x1, "gn_stats", None) if x1 is not None else None
        if (st0 is not None and (x1 is None or st1 is not None) and self._gn_stats_live(st0) and
                (st1 is None or (self._gn_stats_live(st1) and st1.bucket == st0.bucket)) and
                (C0 + C1) % groups == 0 and ((C0 + C1) // groups) % st0.bucket == 0 and C0 % st0.bucket == 0):
            rc = self.lib.gn_group_norm_apply(self.h, x0.data_ptr(), C0, st0.buf.data_ptr(), _ptr(x1), C1,
                                              None if st1 is None else st1.buf.data_ptr(), st0.bucket, B, HW, groups,
                                              float(eps), gamma.data_ptr(), beta.data_ptr(), 1 if silu else 0,
                                              out.data_ptr(), self._stream())
            self.handle.check(rc, "gn_group_norm_apply")
            self.gn_apply_calls += 1
            return out
        rc = self.lib.gn_group_norm(self.h, x0.data_ptr(), C0, _ptr(x1), C1, B, HW, groups, float(eps),
                                    gamma.data_ptr(), beta.data_ptr(), 1 if silu else 0, None, out.data_ptr(),
                                    self._stream())
        self.handle.check(rc, "gn_group_norm")
        return out

    @staticmethod
    def _gn_stats_live(st: GNStats) -> bool:
        """Statistics from an arena are valid until that arena's next reset (any Ops: the producer may be another handle)."""
        return st.owner is None or st.owner._gn_generation == st.generation

    def layer_norm(self, x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
        _f16(x, "x")
        Cn = x.shape[-1]
        x2 = x.reshape(-1, Cn)
        if x2.stride(1) != 1:
            raise ValueError("x must have unit stride in the last dim")
        if out is None:
            out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
        o2 = out.reshape(-1, Cn)
        rc = self.lib.gn_layer_norm(self.h, x2.data_ptr(), x2.stride(0), x2.shape[0], Cn, float(eps),
                                    _f32(gamma, "gamma").data_ptr(), _f32(beta, "beta").data_ptr(), o2.data_ptr(),
                                    o2.stride(0), self._stream())
        self.handle.check(rc, "gn_layer_norm")
        return out

    def softmax_rows(self, x: torch.Tensor, scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """softmax(scale * x) over the last dim of a 2D fp32/fp16 matrix -> fp16 (in place for fp16 input by default)."""
        if x.dim() != 2 or x.stride(1) != 1 or not x.is_cuda or x.dtype not in (torch.float16, torch.float32):
            raise ValueError("x must be a 2D CUDA fp16/fp32 matrix with unit column stride")
        if out is None:
            out = x if x.dtype == torch.float16 else torch.empty(x.shape, dtype=torch.float16, device=x.device)
        rc = self.lib.gn_softmax_rows(self.h, x.data_ptr(), 1 if x.dtype == torch.float32 else 0, x.stride(0),
                                      out.data_ptr(), out.stride(0), x.shape[0], x.shape[1], float(scale),
                                      self._stream())
        self.handle.check(rc, "gn_softmax_rows")
        return out

    # ------------------------------------------------------------------------------------------------ elementwise
    def upsample_nearest2x(self, x: torch.Tensor) -> torch.Tensor:
        _f16(x, "x")
        B, H, W, Cn = x.shape
        y = torch.empty(B, 2 * H, 2 * W, Cn, dtype=torch.float16, device=x.device)
        self.handle.check(self.lib.gn_upsample_nearest2x(self.h, x.data_ptr(), B, H, W, Cn, y.data_ptr(),
                                                         self._stream()), "gn_upsample_nearest2x")
        return y

    def maxpool3x3s2(self, x: torch.Tensor) -> torch.Tensor:
        _f16(x, "x")
        B, H, W, Cn = x.shape
        y = torch.empty(B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, Cn, dtype=torch.float16, device=x.device)
        self.handle.check(self.lib.gn_maxpool3x3s2(self.h, x.data_ptr(), B, H, W, Cn, y.data_ptr(), self._stream()),
                          "gn_maxpool3x3s2")
        return y

    def add(self, a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        _f16(a, "a")
        _f16(b, "b")
        if a.shape != b.shape or not a.is_contiguous() or not b.is_contiguous():
            raise ValueError("add: shapes must match and be contiguous")
        if out is None:
            out = torch.empty_like(a)
        self.handle.check(self.lib.gn_add(self.h, a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(),
                                          self._stream()), "gn_add")
        return out

    def scale(self, x: torch.Tensor, s: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        _f16(x, "x")
        if out is None:
            out = torch.empty_like(x)
        self.handle.check(self.lib.gn_scale(self.h, x.data_ptr(), float(s), out.data_ptr(), x.numel(),
                                            self._stream()), "gn_scale")
        return out

    def tanh_clamp(self, x: torch.Tensor, mag: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """tanh(x / mag) * mag (AutoencoderTiny's latent clamp)."""
        _f16(x, "x")
        if out is None:
            out = torch.empty_like(x)
        self.handle.check(self.lib.gn_tanh_clamp(self.h, x.data_ptr(), float(mag), out.data_ptr(), x.numel(),
                                                 self._stream()), "gn_tanh_clamp")
        return out

    def timestep_embedding(self, t: float, dim: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty(1, dim, dtype=torch.float16, device=self.device)
        self.handle.check(self.lib.gn_timestep_embedding(self.h, float(t), dim, out.data_ptr(), self._stream()),
                          "gn_timestep_embedding")
        return out

    def euler_step(self, x: torch.Tensor, eps: torch.Tensor, sigma: float, sigma_next: float,
                   x_next: Optional[torch.Tensor] = None, x_scaled: Optional[torch.Tensor] = None):
        _f16(x, "x")
        _f16(eps, "eps")
        if x_next is None:
            x_next = torch.empty_like(x)
        rc = self.lib.gn_euler_step(self.h, x.data_ptr(), eps.data_ptr(), float(sigma), float(sigma_next),
                                    x_next.data_ptr(), _ptr(x_scaled), x.numel(), self._stream())
        self.handle.check(rc, "gn_euler_step")
        return x_next, x_scaled

    def euler_ancestral_step(self, x: torch.Tensor, eps: torch.Tensor, noise: torch.Tensor, sigma: float,
                             sigma_down: float, sigma_up: float, sigma_next: float,
                             x_next: Optional[torch.Tensor] = None, x_scaled: Optional[torch.Tensor] = None):
        """EulerAncestralDiscreteScheduler.step: x + (sigma_down - sigma) eps + sigma_up noise (+ next step's scaling)."""
        _f16(x, "x")
        _f16(eps, "eps")
        _f16(noise, "noise")
        if noise.shape != x.shape or not noise.is_contiguous():
            raise ValueError("noise must be contiguous and shaped like x")
        if x_next is None:
            x_next = torch.empty_like(x)
        rc = self.lib.gn_euler_ancestral_step(self.h, x.data_ptr(), eps.data_ptr(), noise.data_ptr(), float(sigma),
                                              float(sigma_down), float(sigma_up), float(sigma_next),
                                              x_next.data_ptr(), _ptr(x_scaled), x.numel(), self._stream())
        self.handle.check(rc, "gn_euler_ancestral_step")
        return x_next, x_scaled

    def nchw_to_nhwc(self, src: torch.Tensor, cpad: Optional[int] = None, mean=None, std=None) -> torch.Tensor:
        if not src.is_cuda or not src.is_contiguous() or src.dtype not in (torch.float16, torch.float32, torch.uint8):
            raise TypeError("nchw_to_nhwc: contiguous CUDA fp16/fp32/uint8 tensor expected")
        B, Cn, H, W = src.shape
        cpad = cpad or Cn
        dst = torch.empty(B, H, W, cpad, dtype=torch.float16, device=src.device)
        m = (C.c_float * 3)(*mean) if mean is not None else None
        s = (C.c_float * 3)(*std) if std is not None else None
        code = {torch.float16: 0, torch.float32: 1, torch.uint8: 2}[src.dtype]
        rc = self.lib.gn_nchw_to_nhwc(self.h, src.data_ptr(), code, B, Cn, H, W,
                                      cpad, m, s, dst.data_ptr(), self._stream())
        self.handle.check(rc, "gn_nchw_to_nhwc")
        return dst

    def nhwc_to_nchw(self, src: torch.Tensor, channels: Optional[int] = None, fp32: bool = False) -> torch.Tensor:
        _f16(src, "src")
        B, H, W, cpad = src.shape
        Cn = channels or cpad
        dst = torch.empty(B, Cn, H, W, dtype=torch.float32 if fp32 else torch.float16, device=src.device)
        rc = self.lib.gn_nhwc_to_nchw(self.h, src.data_ptr(), B, Cn, H, W, cpad, dst.data_ptr(), 1 if fp32 else 0,
                                      self._stream())
        self.handle.check(rc, "gn_nhwc_to_nchw")
        return dst

    def u8_to_nhwc(self, src: torch.Tensor, cpad: int = 64, mean=None, std=None,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if src.dtype != torch.uint8 or not src.is_cuda or not src.is_contiguous() or src.shape[-1] != 3:
            raise TypeError("u8_to_nhwc: contiguous CUDA uint8 [B, H, W, 3] expected")
        B, H, W, _ = src.shape
        if out is None:
            out = torch.empty(B, H, W, cpad, dtype=torch.float16, device=src.device)
        m = (C.c_float * 3)(*mean) if mean is not None else None
        s = (C.c_float * 3)(*std) if std is not None else None
        rc = self.lib.gn_u8_to_nhwc(self.h, src.data_ptr(), B, H, W, cpad, m, s, out.data_ptr(), self._stream())
        self.handle.check(rc, "gn_u8_to_nhwc")
        return out

    def nhwc_to_u8(self, src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        _f16(src, "src")
        B, H, W, cpad = src.shape
        if out is None:
            out = torch.empty(B, H, W, 3, dtype=torch.uint8, device=src.device)
        rc = self.lib.gn_nhwc_to_u8(self.h, src.data_ptr(), B, H, W, cpad, out.data_ptr(), self._stream())
        self.handle.check(rc, "gn_nhwc_to_u8")
        return out

    def tile_views(self, views: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[B, 4, S, S, 3] u8 -> [B, 2S, 2S, 3] u8 (controller/utils/misc.py:6-19; S = 256 there)."""
        if (views.dtype != torch.uint8 or not views.is_cuda or views.dim() != 5 or views.shape[1] != 4
                or views.shape[2] != views.shape[3] or views.shape[4] != 3):
            raise TypeError("tile_views: CUDA uint8 [B, 4, S, S, 3] expected")
        B, S = views.shape[0], views.shape[2]
        if out is None:
            out = torch.empty(B, 2 * S, 2 * S, 3, dtype=torch.uint8, device=views.device)
        self.handle.check(self.lib.gn_tile_views(self.h, views.contiguous().data_ptr(), B, S, out.data_ptr(),
                                                 self._stream()), "gn_tile_views")
        return out

    def untile_views(self, tile: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[B, 2S, 2S, 3] u8 -> [B, 4, S, S, 3] u8 (controller/utils/misc.py:22-47; S = 256 there)."""
        if (tile.dtype != torch.uint8 or not tile.is_cuda or tile.dim() != 4 or tile.shape[1] != tile.shape[2]
                or tile.shape[1] % 2 or tile.shape[3] != 3):
            raise TypeError("untile_views: CUDA uint8 [B, 2S, 2S, 3] expected")
        B, S = tile.shape[0], tile.shape[1] // 2
        if out is None:
            out = torch.empty(B, 4, S, S, 3, dtype=torch.uint8, device=tile.device)
        self.handle.check(self.lib.gn_untile_views(self.h, tile.contiguous().data_ptr(), B, S, out.data_ptr(),
                                                   self._stream()), "gn_untile_views")
        return out

    def embed_tokens(self, ids: torch.Tensor, tok_emb: torch.Tensor, pos_emb: torch.Tensor) -> torch.Tensor:
        if ids.dtype != torch.int64 or not ids.is_cuda or ids.dim() != 2:
            raise TypeError("embed_tokens: CUDA int64 [B, T] ids expected")
        _f16(tok_emb, "tok_emb")
        _f16(pos_emb, "pos_emb")
        B, T = ids.shape
        V, D = tok_emb.shape
        out = torch.empty(B, T, D, dtype=torch.float16, device=ids.device)
        rc = self.lib.gn_embed_tokens(self.h, ids.contiguous().data_ptr(), tok_emb.data_ptr(), pos_emb.data_ptr(), B,
                                      T, D, V, out.data_ptr(), self._stream())
        self.handle.check(rc, "gn_embed_tokens")
        return out

    def film_fold(self, film: torch.Tensor, bn_scale: torch.Tensor, bn_shift: torch.Tensor):
        Cn = bn_scale.numel()
        _f32(film, "film")
        if film.numel() != 2 * Cn:
            raise ValueError("film must hold [gamma | beta] = 2*C floats")
        scale = torch.empty(Cn, dtype=torch.float32, device=film.device)
        shift = torch.empty(Cn, dtype=torch.float32, device=film.device)
        rc = self.lib.gn_film_fold(self.h, film.data_ptr(), _f32(bn_scale, "bn_scale").data_ptr(),
                                   _f32(bn_shift, "bn_shift").data_ptr(), scale.data_ptr(), shift.data_ptr(), Cn,
                                   self._stream())
        self.handle.check(rc, "gn_film_fold")
        return scale, shift
