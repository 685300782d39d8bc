"""ctypes binding of the xrft_b200 C-ABI (include/xrft_b200.h).

The library is built in-tree (``xrft_b200/libxrftb200.so``) by ``__graft_entry__.build()`` or
``make -C xrft_b200/csrc``.  There is no CPU fallback: if the library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("XRFTB_LIB", os.path.join(_HERE, "libxrftb200.so"))

# every symbol include/xrft_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "xrftb_version",
    "xrftb_last_error",
    "xrftb_device_info",
    "xrftb_launch_count",
    "xrftb_profile_begin",
    "xrftb_profile_end",
    "xrftb_set_option",
    "xrftb_get_option",
    "xrftb_spectrum2d_last_path",
    "xrftb_fftn_workspace",
    "xrftb_fftn",
    "xrftb_fft2r_workspace",
    "xrftb_fft2r",
    "xrftb_moments",
    "xrftb_detrend_window",
    "xrftb_spectral_post",
    "xrftb_spectral_post_segmean",
    "xrftb_roll_scale",
    "xrftb_permute",
    "xrftb_pad",
    "xrftb_binned_sum",
    "xrftb_spectrum2d_workspace",
    "xrftb_spectrum2d",
    "xrftb_comm_unique_id",
    "xrftb_comm_init",
    "xrftb_comm_destroy",
    "xrftb_comm_nccl_version",
    "xrftb_allreduce_bins",
]

F32, F64 = 0, 1
C2C_FWD, C2C_INV, R2C, C2R = 0, 1, 2, 3
EPI_COMPLEX, EPI_POWER, EPI_CROSS, EPI_PHASE, EPI_BINS_POWER, EPI_BINS_CROSS = range(6)


class Spectrum2dDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int),
        ("batch", C.c_int64),
        ("ny", C.c_int),
        ("nx", C.c_int),
        ("in1", C.c_void_p),
        ("in2", C.c_void_p),
        ("detrend", C.c_int),
        ("win_y", C.c_void_p),
        ("win_x", C.c_void_p),
        ("mode", C.c_int),
        ("keep_half", C.c_int),
        ("shift_y", C.c_int),
        ("shift_x", C.c_int),
        ("scale", C.c_double),
        ("ramp_y", C.c_void_p),
        ("ramp_x", C.c_void_p),
        ("weight_x", C.c_void_p),
        ("out", C.c_void_p),
        ("lut", C.c_void_p),
        ("bins", C.c_void_p),
        ("nbins", C.c_int),
        ("lut_symmetric", C.c_int),
        ("work", C.c_void_p),
        ("work_bytes", C.c_size_t),
        ("out2", C.c_void_p),
    ]


class Fft2rDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("inverse", C.c_int), ("batch", C.c_int64), ("ny", C.c_int64), ("nx", C.c_int64),
        ("in_", C.c_void_p), ("out", C.c_void_p),
        ("in_ny", C.c_int64), ("in_nx", C.c_int64), ("in_off_y", C.c_int64), ("in_off_x", C.c_int64),
        ("ramp_y", C.c_void_p), ("ramp_x", C.c_void_p), ("scale", C.c_double),
        ("in_roll_y", C.c_int64), ("out_roll_y", C.c_int64), ("out_roll_x", C.c_int64),
        ("out_ny", C.c_int64), ("out_nx", C.c_int64), ("out_off_y", C.c_int64), ("out_off_x", C.c_int64),
        ("work", C.c_void_p), ("work_bytes", C.c_size_t),
    ]


_lib = None


class XrftbError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XrftbError(
            f"xrft_b200 CUDA library not found at {LIB_PATH}; build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C xrft_b200/csrc`. "
            "There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, i64p, ip = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int)
    lib.xrftb_version.restype = C.c_int
    lib.xrftb_last_error.restype = C.c_char_p
    lib.xrftb_device_info.argtypes = [ip, ip, ip, C.POINTER(C.c_size_t)]
    lib.xrftb_launch_count.restype = C.c_long
    lib.xrftb_launch_count.argtypes = [C.c_int]
    lib.xrftb_profile_end.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_long)]
    lib.xrftb_fftn_workspace.restype = C.c_size_t
    lib.xrftb_fftn_workspace.argtypes = [C.c_int, C.c_int, C.c_int, i64p, C.c_int, ip]
    lib.xrftb_fftn.argtypes = [vp, vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, i64p, C.c_int, ip, vp]
    lib.xrftb_moments.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, vp]
    lib.xrftb_detrend_window.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, vp]
    lib.xrftb_spectral_post.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                        C.c_int, ip, C.POINTER(vp), vp, C.c_double, vp]
    lib.xrftb_spectral_post_segmean.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                                C.c_int, ip, C.POINTER(vp), vp, C.c_double, C.c_int64, C.c_int64, vp]
    lib.xrftb_spectral_post_segmean.restype = C.c_int
    lib.xrftb_roll_scale.argtypes = [vp, vp, C.c_int, C.c_int] + [C.c_int64] * 7 + [C.c_double, vp]
    lib.xrftb_fft2r_workspace.restype = C.c_size_t
    lib.xrftb_fft2r_workspace.argtypes = [C.POINTER(Fft2rDesc)]
    lib.xrftb_fft2r.restype = C.c_int
    lib.xrftb_fft2r.argtypes = [C.POINTER(Fft2rDesc), vp]
    lib.xrftb_permute.argtypes = [vp, vp, C.c_int, C.c_int, i64p, ip, ip, vp]
    lib.xrftb_permute.restype = C.c_int
    lib.xrftb_pad.argtypes = [vp, vp, C.c_int, C.c_int, i64p, i64p, i64p, C.c_int, vp, vp]
    lib.xrftb_pad.restype = C.c_int
    lib.xrftb_binned_sum.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, vp]
    lib.xrftb_spectrum2d_workspace.restype = C.c_size_t
    lib.xrftb_spectrum2d_workspace.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]
    lib.xrftb_spectrum2d.argtypes = [C.POINTER(Spectrum2dDesc), vp]
    lib.xrftb_set_option.argtypes = [C.c_char_p, C.c_int]
    lib.xrftb_set_option.restype = C.c_int
    lib.xrftb_get_option.argtypes = [C.c_char_p, ip]
    lib.xrftb_get_option.restype = C.c_int
    lib.xrftb_comm_unique_id.argtypes = [vp]
    lib.xrftb_comm_init.argtypes = [C.POINTER(vp), C.c_int, C.c_int, vp]
    lib.xrftb_comm_destroy.argtypes = [vp]
    lib.xrftb_comm_nccl_version.restype = C.c_int
    lib.xrftb_allreduce_bins.argtypes = [vp, vp, C.c_size_t, vp]
    for name in ["xrftb_comm_unique_id", "xrftb_comm_init", "xrftb_comm_destroy", "xrftb_allreduce_bins"]:
        getattr(lib, name).restype = C.c_int
    for name in ["xrftb_device_info", "xrftb_fftn", "xrftb_moments", "xrftb_detrend_window", "xrftb_spectral_post",
                 "xrftb_permute",
    "xrftb_pad",
    "xrftb_binned_sum", "xrftb_spectrum2d", "xrftb_roll_scale"]:
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def set_option(name: str, value: int):
    """xrftb_set_option: choose a kernel chain explicitly (see include/xrft_b200.h for the names)"""
    check(load().xrftb_set_option(name.encode(), int(value)), "xrftb_set_option")


def get_option(name: str) -> int:
    v = C.c_int(0)
    check(load().xrftb_get_option(name.encode(), C.byref(v)), "xrftb_get_option")
    return v.value


def check(rc: int, what: str):
    if rc != 0:
        msg = load().xrftb_last_error().decode("utf-8", "replace")
        if rc == -2:
            raise NotImplementedError(f"{what}: {msg}")
        raise XrftbError(f"{what} failed (code {rc}): {msg}")
