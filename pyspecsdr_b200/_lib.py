"""ctypes binding of libpss.so (C ABI declared in include/pss.h).

The product path has NO CPU fallback: if the shared library is missing or does not load, importing
this module raises, loudly.  Build it with `python __graft_entry__.py build` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PSS_LIB: another build of the same library (a compile-time variant under measurement), never a fallback
LIB_PATH = os.environ.get("PSS_LIB") or os.path.join(_HERE, "libpss.so")


class PssError(RuntimeError):
    pass


class _Sized(C.Structure):
    """Every ABI struct starts with `size_t struct_size` = sizeof(struct); set here at construction."""

    def __init__(self, *args, **kw):
        super().__init__(C.sizeof(type(self)), *args, **kw)


class PsdOut(_Sized):
    """Mirror of pss_psd_out (include/pss.h)."""
    _fields_ = [("struct_size", C.c_size_t), ("db", C.c_void_p), ("cols", C.c_void_p), ("W", C.c_int),
                ("stats", C.c_void_p), ("moments", C.c_void_p)]


class DisplayOut(_Sized):
    """Mirror of pss_display_out (include/pss.h)."""
    _fields_ = [("struct_size", C.c_size_t), ("norm", C.c_void_p), ("minmax", C.c_void_p), ("norm64", C.c_void_p),
                ("minmax64", C.c_void_p), ("plane_a", C.c_void_p), ("plane_b", C.c_void_p), ("n_rows", C.c_void_p)]


class DemodDesc(_Sized):
    """Mirror of pss_demod_desc (include/pss.h)."""
    _dp = C.POINTER(C.c_double)
    _fields_ = [
        ("struct_size", C.c_size_t), ("kind", C.c_int), ("mode", C.c_int), ("N", C.c_int),
        ("q", C.c_int), ("n_out", C.c_int), ("lead", C.c_int), ("SF", C.c_int), ("SB", C.c_int),
        ("n_body", C.c_int), ("m_tail", C.c_int), ("tail_start", C.c_int), ("tail_len", C.c_int),
        ("scale", C.c_float), ("norm", C.c_float), ("iq_correct", C.c_int),
        ("body", _dp), ("BF", _dp), ("BB", _dp), ("G", _dp), ("CR", _dp),
        ("CB", _dp), ("DB", C.c_double), ("head", _dp), ("tail_T", _dp), ("tail_M", _dp),
        ("taps", _dp), ("n_taps", C.c_int),
        ("sos", _dp), ("n_sections", C.c_int),
    ]


class PipelineIO(_Sized):
    """Mirror of pss_pipeline_io (include/pss.h)."""
    _fields_ = [("struct_size", C.c_size_t), ("N_block", C.c_int), ("N_fft", C.c_int), ("W", C.c_int),
                ("rows_max", C.c_int), ("plan", C.c_void_p), ("audio", C.c_void_p), ("cols", C.c_void_p),
                ("stats", C.c_void_p), ("db", C.c_void_p), ("norm", C.c_void_p), ("minmax", C.c_void_p),
                ("display_stream", C.c_int), ("plane_a", C.c_void_p), ("plane_b", C.c_void_p)]


def _signatures():
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
    return {
        "pss_init": (i32, [i32, C.POINTER(vp)]),
        "pss_destroy": (None, [vp]),
        "pss_strerror": (C.c_char_p, [i32]),
        "pss_last_error": (C.c_char_p, [vp]),
        "pss_version": (i32, []),
        "pss_set_stream": (i32, [vp, vp]),
        "pss_use_own_stream": (i32, [vp]),
        "pss_sync": (i32, [vp]),
        "pss_kernel_launches": (i64, [vp]),
        "pss_host_alloc": (vp, [C.c_size_t]),
        "pss_host_free": (None, [vp]),
        "pss_psd_c64": (i32, [vp, vp, i32, i64, i32, i32, i32, C.POINTER(PsdOut)]),
        "pss_psd_c64_dev": (i32, [vp, vp, i32, i64, i32, i32, i32, C.POINTER(PsdOut)]),
        "pss_scan_c64": (i32, [vp, vp, i32, i64, i32, f32, vp, vp, vp]),
        "pss_scan_c64_dev": (i32, [vp, vp, i32, i64, i32, f32, vp, vp, vp]),
        "pss_display_render": (i32, [vp, vp, vp, i32, i64, i32, i64, i64, i64, i32, vp, vp]),
        "pss_display_render_dev": (i32, [vp, vp, vp, i32, i64, i32, i64, i64, i64, i32, vp, vp]),
        "pss_pipeline_c64": (i32, [vp, vp, i64, C.POINTER(PipelineIO)]),
        "pss_iq_correct_c64": (i32, [vp, vp, i32, i64, vp]),
        "pss_sosfilt_f32": (i32, [vp, vp, i32, i64, vp, i32, vp]),
        "pss_power_c64": (i32, [vp, vp, i32, i64, vp]),
        "pss_audio_to_int16": (i32, [vp, vp, i64, vp]),
        "pss_classify_c64": (i32, [vp, vp, i32, i64, C.c_double, vp, vp]),
        "pss_classify_c64_dev": (i32, [vp, vp, i32, i64, C.c_double, vp, vp]),
        "pss_display_quantise": (i32, [vp, vp, i64, i32, i32, vp, vp]),
        "pss_display_open": (i32, [vp, i32, i32, i32, i32, i32]),
        "pss_display_close": (i32, [vp, i32]),
        "pss_display_rows": (i32, [vp, i32]),
        "pss_display_accumulate_f64": (i32, [vp, i32, vp, i32, i64, C.POINTER(DisplayOut)]),
        "pss_display_accumulate_dev": (i32, [vp, i32, vp, vp, i64, i64, i64, i64, C.POINTER(DisplayOut)]),
        "pss_spectrum_normalise": (i32, [vp, vp, i32, i64, i32, vp, vp]),
        "pss_spectrum_normalise_f64": (i32, [vp, vp, i32, i64, i32, vp, vp]),
        "pss_audio_to_int16_f64": (i32, [vp, vp, i64, vp]),
        "pss_demod_plan_block_len": (i32, [vp]),
        "pss_demod_plan_create": (i32, [vp, C.POINTER(DemodDesc), C.POINTER(vp)]),
        "pss_demod_plan_destroy": (None, [vp, vp]),
        "pss_demod_plan_out_len": (i32, [vp]),
        "pss_demod_plan_channels": (i32, [vp]),
        "pss_demod_c64": (i32, [vp, vp, vp, i64, vp]),
        "pss_demod_c64_dev": (i32, [vp, vp, vp, i64, vp]),
        "pss_demod_c64_dev_moments": (i32, [vp, vp, vp, i64, vp, vp, i32, i32]),
    }


def _load():
    if not os.path.exists(LIB_PATH):
        raise PssError(
            f"{LIB_PATH} not found: the CUDA library is not built. Run `python __graft_entry__.py build` "
            "(there is deliberately no numpy fallback in the product path).")
    handle = C.CDLL(LIB_PATH)
    sigs = _signatures()
    for name, (res, args) in sigs.items():
        fn = getattr(handle, name)       # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    return handle, sigs


lib, SIGNATURES = _load()


def check(ctx_handle, rc: int, what: str):
    if rc != 0:
        msg = lib.pss_strerror(rc).decode()
        detail = lib.pss_last_error(ctx_handle).decode() if ctx_handle else ""
        raise PssError(f"{what}: {msg}" + (f" [{detail}]" if detail else ""))
