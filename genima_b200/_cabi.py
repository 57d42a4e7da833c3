"""ctypes binding of libgenima_b200.so (the C ABI declared in include/genima_b200.h).

This is the ONLY route from Python to the CUDA kernels.  There is no CPU or PyTorch fallback: if the shared library
is missing, or no sm_100 device is present, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgenima_b200.so")

GN_OK = 0
GN_ERR_INVALID = -1
GN_ERR_CUDA = -2
GN_ERR_NOMEM = -3
GN_ERR_NODRIVER = -4

ACT_NONE, ACT_SILU, ACT_GELU, ACT_RELU, ACT_QUICKGELU = 0, 1, 2, 3, 4


class GnEpilogue(C.Structure):
    """Mirror of `struct gn_epilogue`."""

    _fields_ = [
        ("scale", C.c_void_p),
        ("bias", C.c_void_p),
        ("rowvec", C.c_void_p),
        ("residual", C.c_void_p),
        ("ldr", C.c_int64),
        ("rows_per_batch", C.c_int32),
        ("act_pre", C.c_int32),
        ("act_post", C.c_int32),
        ("alpha", C.c_float),
        ("beta", C.c_float),
        ("geglu", C.c_int32),
        ("out_fp32", C.c_int32),
        ("ln_stats", C.c_void_p),
        ("ln_colsum", C.c_void_p),
        ("ln_parts", C.c_int32),
        ("ln_eps", C.c_float),
        ("rowstats_out", C.c_void_p),
        ("rowstats_capacity", C.c_int32),
        ("gn_bucket", C.c_int32),
        ("gnstats_out", C.c_void_p),
        ("w_dynamic", C.c_int32),
        ("reserved", C.c_int32),
    ]


# name -> (restype, argtypes); must list EVERY symbol include/genima_b200.h declares (tests/test_cabi.py checks).
_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
SIGNATURES = {
    "gn_create": (_i, [_i, C.POINTER(_vp)]),
    "gn_destroy": (_i, [_vp]),
    "gn_last_error": (C.c_char_p, [_vp]),
    "gn_version": (C.c_char_p, []),
    "gn_set_workspace": (_i, [_vp, _vp, _i64]),
    "gn_set_gemm_tuning": (_i, [_vp, _i, _i]),
    "gn_set_pdl": (_i, [_vp, _i]),
    "gn_set_gn_max_ctas": (_i, [_vp, _i]),
    "gn_set_staged_epilogue": (_i, [_vp, _i]),
    "gn_set_autotune": (_i, [_vp, _i]),
    "gn_tune_cache_export": (_i64, [_vp, C.c_char_p, _i64]),
    "gn_tune_cache_import": (_i, [_vp, C.c_char_p, _i64, _i]),
    "gn_set_gemm_occupancy": (_i, [_vp, _i]),
    "gn_set_attention_kv_split": (_i, [_vp, _i]),
    "gn_set_gemm_pair": (_i, [_vp, _i]),
    "gn_last_gemm_pair": (_i, [_vp]),
    "gn_set_gemm_trace": (_i, [_vp, _vp]),
    "gn_get_last_gemm_config": (_i, [_vp, C.POINTER(C.c_int32)]),
    "gn_get_last_rowstats_parts": (_i, [_vp]),
    "gn_launch_count": (_i64, [_vp]),
    "gn_profile_begin": (_i, [_vp]),
    "gn_profile_end": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double),
                            C.POINTER(C.c_double)]),
    "gn_linear": (_i, [_vp, _vp, _i64, _i, _i, _vp, _i, _vp, _i64, C.POINTER(GnEpilogue), _vp]),
    "gn_conv2d": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i64,
                       C.POINTER(GnEpilogue), _vp]),
    "gn_conv2d_asym": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i64,
                            C.POINTER(GnEpilogue), _vp]),
    "gn_conv2d_up2x": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _i64, C.POINTER(GnEpilogue), _vp]),
    "gn_attention": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _f, _vp]),
    "gn_attention_qproj": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _vp, _i, _f, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i,
                                _f, _vp]),
    "gn_attention_small": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "gn_group_norm": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _f, _vp, _vp, _i, _vp, _vp, _vp]),
    "gn_group_norm_apply": (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _f, _vp, _vp, _i, _vp, _vp]),
    "gn_layer_norm": (_i, [_vp, _vp, _i64, _i, _i, _f, _vp, _vp, _vp, _i64, _vp]),
    "gn_softmax_rows": (_i, [_vp, _vp, _i, _i64, _vp, _i64, _i, _i, _f, _vp]),
    "gn_upsample_nearest2x": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "gn_maxpool3x3s2": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "gn_add": (_i, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "gn_timestep_embedding": (_i, [_vp, _f, _i, _vp, _vp]),
    "gn_euler_step": (_i, [_vp, _vp, _vp, _f, _f, _vp, _vp, _i64, _vp]),
    "gn_euler_ancestral_step": (_i, [_vp, _vp, _vp, _vp, _f, _f, _f, _f, _vp, _vp, _i64, _vp]),
    "gn_scale": (_i, [_vp, _vp, _f, _vp, _i64, _vp]),
    "gn_tanh_clamp": (_i, [_vp, _vp, _f, _vp, _i64, _vp]),
    "gn_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _vp, _vp]),
    "gn_nhwc_to_nchw": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "gn_u8_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _vp, _vp]),
    "gn_nhwc_to_u8": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "gn_tile_views": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "gn_untile_views": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "gn_embed_tokens": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "gn_film_fold": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
}

_lib: Optional[C.CDLL] = None


class GenimaB200Error(RuntimeError):
    pass


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """Load the shared library and bind every declared symbol.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise GenimaB200Error(
            f"{path} not found: build it with `python -m genima_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class Handle:
    """Owns one `gn_handle` (one per device / host thread)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self._h = C.c_void_p()
        rc = self.lib.gn_create(int(device), C.byref(self._h))
        if rc != GN_OK:
            self._h = C.c_void_p()
            reason = {GN_ERR_NODRIVER: "no CUDA driver/device", GN_ERR_INVALID: "not an sm_100 device (or bad index)",
                      GN_ERR_CUDA: "CUDA error", GN_ERR_NOMEM: "out of memory"}.get(rc, "unknown")
            raise GenimaB200Error(f"gn_create(device={device}) failed with {rc}: {reason}")
        self.device = device

    def check(self, rc: int, what: str = "") -> None:
        if rc != GN_OK:
            msg = self.lib.gn_last_error(self._h)
            raise GenimaB200Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    @property
    def ptr(self) -> C.c_void_p:
        return self._h

    def launch_count(self) -> int:
        return int(self.lib.gn_launch_count(self._h))

    def close(self) -> None:
        if self._h:
            self.lib.gn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
