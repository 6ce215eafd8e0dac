"""ctypes binding of libmcquic_b200.so (include/mcquic_b200.h).  There is NO fallback: if the CUDA
library cannot be loaded every op raises."""
import ctypes
import os

from .csrc import build as _build

_c = ctypes
_LIB = None


class ConvParams(_c.Structure):
    """mirror of `mcq_conv_params` (include/mcquic_b200.h)"""
    _fields_ = [
        ("a_hi", _c.c_void_p), ("a_lo", _c.c_void_p),
        ("n", _c.c_int32), ("hin", _c.c_int32), ("win", _c.c_int32), ("cin", _c.c_int32),
        ("w_hi", _c.c_void_p), ("w_lo", _c.c_void_p),
        ("cout", _c.c_int32), ("cout_pad", _c.c_int32), ("ksize", _c.c_int32), ("stride", _c.c_int32),
        ("w_scale", _c.c_float),
        ("bias", _c.c_void_p),
        ("mode", _c.c_int32), ("store", _c.c_int32),
        ("res1", _c.c_void_p), ("res1_scale", _c.c_float),
        ("res2", _c.c_void_p), ("aux", _c.c_void_p),
        ("out_f32", _c.c_void_p),
        ("out0_hi", _c.c_void_p), ("out0_lo", _c.c_void_p), ("out0_act", _c.c_int32),
        ("out1_hi", _c.c_void_p), ("out1_lo", _c.c_void_p), ("out1_act", _c.c_int32),
        ("passes", _c.c_int32), ("impl", _c.c_int32),
        ("ev_start", _c.c_void_p), ("ev_stop", _c.c_void_p),
        ("gn_partials", _c.c_void_p), ("gn_groups", _c.c_int32),
        ("out_u8", _c.c_void_p),
        ("dev_scale", _c.c_void_p),
    ]


class WgradParams(_c.Structure):
    """mirror of `mcq_wgrad_params` (include/mcquic_b200.h)"""
    _fields_ = [
        ("x_hi", _c.c_void_p), ("n", _c.c_int32), ("hin", _c.c_int32), ("win", _c.c_int32), ("cin", _c.c_int32),
        ("dy_hi", _c.c_void_p), ("cout", _c.c_int32), ("ksize", _c.c_int32), ("stride", _c.c_int32),
        ("dw", _c.c_void_p), ("scale", _c.c_float), ("dev_scale", _c.c_void_p), ("accumulate", _c.c_int32),
        ("workspace", _c.c_void_p), ("workspace_bytes", _c.c_int64),
    ]


EPI_LINEAR, EPI_GATE, EPI_GDN, EPI_IGDN = 0, 1, 2, 3
ACT_NONE, ACT_SILU, ACT_SQUARE = 0, 1, 2
SQUARE_SCALE = 2.0 ** -6      # MCQ_SQUARE_SCALE: ACT_SQUARE planes hold y^2 * 2^-6 (fp16 range)
STORE_NHWC, STORE_SHUFFLE_NHWC, STORE_SHUFFLE_NCHW = 0, 1, 2
IMPL_TCGEN05, IMPL_SIMT = 0, 1
ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_DRIVER, ERR_CODE_RANGE, ERR_WATCHDOG = -1, -2, -3, -4, -5

# every symbol include/mcquic_b200.h declares: name -> (restype, argtypes)
_i32, _i64, _p, _f = _c.c_int32, _c.c_int64, _c.c_void_p, _c.c_float
SYMBOLS = {
    "mcq_conv2d": (_c.c_int, [_c.POINTER(ConvParams), _p]),
    "mcq_conv_chain": (_c.c_int, [_c.POINTER(ConvParams), _i32, _p]),
    "mcq_conv_chain_max_layers": (_i32, []),
    "mcq_debug_timeline": (None, [_p]),
    "mcq_stem_conv": (_c.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _i32, _p, _p, _p, _i32, _p]),
    "mcq_stem_conv_tc": (_c.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _f, _p, _i32, _i32, _p, _p, _p,
                                    _i32, _p]),
    "mcq_vq_assign": (_c.c_int, [_p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "mcq_vq_assign_tc": (_c.c_int, [_p, _p, _p, _f, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _i64, _p]),
    "mcq_vq_assign_fused": (_c.c_int, [_p, _p, _f, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "mcq_vq_fused_supported": (_c.c_int, [_i32, _i32, _i32, _i32]),
    "mcq_vq_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "mcq_vq_dequant": (_c.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _i32, _p, _p, _i32, _p, _p]),
    "mcq_code_histogram": (_c.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p]),
    "mcq_groupnorm": (_c.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _p, _p, _f, _p, _p, _p, _i32, _p]),
    "mcq_conv_gn_layout": (_c.c_int, [_c.POINTER(ConvParams), _c.POINTER(_i32), _c.POINTER(_i32)]),
    "mcq_groupnorm_apply": (_c.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _f, _p, _p, _p, _p, _i32, _p]),
    "mcq_add_scaled": (_c.c_int, [_p, _p, _f, _i64, _p, _p, _p, _i32, _p]),
    "mcq_split_planes": (_c.c_int, [_p, _i64, _i32, _p, _p, _p, _p]),
    "mcq_conv_wgrad_workspace_bytes": (_i64, [_c.POINTER(WgradParams)]),
    "mcq_conv_wgrad": (_c.c_int, [_c.POINTER(WgradParams), _p]),
    "mcq_nchw_to_nhwc": (_c.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p, _p, _i32, _p, _p, _i32, _p]),
    "mcq_nhwc_to_nchw": (_c.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p]),
    "mcq_error_string": (_c.c_char_p, [_c.c_int]),
    "mcq_version": (_c.c_int, []),
    "mcq_set_option": (_c.c_int, [_c.c_char_p, _i32]),
    "mcq_get_option": (_i32, [_c.c_char_p]),
    "mcq_device_error_flag": (_c.c_int, []),
    "mcq_kernel_launch_count": (_c.c_int, []),
}


def library_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load (building first if the .so is absent or stale and nvcc is available)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if build_if_missing:
        try:
            _build.build()
        except Exception as e:  # no nvcc on this box: use the shipped .so if there is one
            if not os.path.exists(path):
                raise RuntimeError(f"mcquic_b200: CUDA library missing and cannot be built: {e}") from e
    if not os.path.exists(path):
        raise RuntimeError(f"mcquic_b200: {path} not found; run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = _c.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().mcq_error_string(rc).decode()
        raise RuntimeError(f"{what}: {msg} (code {rc})")


def set_option(name: str, value: int):
    """explicit tuning / A-B knob of the library (include/mcquic_b200.h: mcq_set_option)"""
    check(load().mcq_set_option(name.encode(), int(value)), f"mcq_set_option({name})")


def apply_options(spec: str):
    """"name=value,name=value" -> set_option for each (measurement scripts under tools/ take such a string)"""
    for item in filter(None, (t.strip() for t in (spec or "").split(","))):
        name, _, value = item.partition("=")
        set_option(name.strip(), int(value))


def get_option(name: str) -> int:
    return int(load().mcq_get_option(name.encode()))


def launch_count() -> int:
    return int(load().mcq_kernel_launch_count())
