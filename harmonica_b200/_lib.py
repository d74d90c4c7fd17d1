"""
ctypes binding of ``libharmonica_b200.so`` (C ABI in ``include/harmonica_b200.h``).

There is no CPU fallback: if the shared library is missing or no sm_100 device
is usable, every compute entry point raises.
"""

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libharmonica_b200.so")

HB200_OK = 0
HB200_EZERODIV = -5
FLAG_SINGULAR = 1
FLAG_ZERO_DIV = 2
FLAG_TESS_STACK = 4
FLAG_TESS_LEAVES = 8
FLAG_TESS_INSIDE = 16
SHARD_AUTO, SHARD_OBSERVERS, SHARD_SOURCES = 0, 1, 2
MAG_DEFAULT_RULES = 3

# field ids == bit positions of field_mask (include/harmonica_b200.h)
FIELD_IDS = {
    "potential": 0, "g_e": 1, "g_n": 2, "g_z": 3,
    "g_ee": 4, "g_nn": 5, "g_zz": 6, "g_en": 7, "g_ez": 8, "g_nz": 9,
}  # fmt: skip
MASK_ACCEL = 0x00E
MASK_TENSOR = 0x3F0

_dp = ctypes.POINTER(ctypes.c_double)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i64 = ctypes.c_int64
_u32 = ctypes.c_uint32
_int = ctypes.c_int
_vp = ctypes.c_void_p
_sz = ctypes.c_size_t
# hb200_density_fn: void (*)(const double* radius, double* density_out, int64_t n, void* user)
DENSITY_FN = ctypes.CFUNCTYPE(None, _dp, _dp, _i64, _vp)

# every symbol include/harmonica_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "hb200_version": (_int, []),
    "hb200_device_count": (_int, []),
    "hb200_init": (_int, [ctypes.POINTER(_int), _int]),
    "hb200_num_devices": (_int, []),
    "hb200_shutdown": (None, []),
    "hb200_last_error": (ctypes.c_char_p, []),
    "hb200_set_variant": (_int, [_int]),
    "hb200_get_variant": (_int, []),
    "hb200_set_tesseroid_variant": (_int, [_int]),
    "hb200_get_tesseroid_variant": (_int, []),
    "hb200_set_tile_mode": (_int, [_int]),
    "hb200_get_tile_mode": (_int, []),
    "hb200_set_source_chunks": (_int, [_int]),
    "hb200_get_source_chunks": (_int, []),
    "hb200_set_fit_rcond": (_int, [ctypes.c_double]),
    "hb200_get_fit_rcond": (ctypes.c_double, []),
    "hb200_launch_count": (ctypes.c_uint64, []),
    "hb200_prism_gravity": (
        _int, [_dp, _dp, _dp, _i64, _dp, _dp, _i64, _u32, _int, _dp, _u32p]),
    "hb200_prism_singular_scan": (_int, [_dp, _dp, _dp, _i64, _dp, _i64, _int, _u32p]),
    "hb200_prism_magnetic": (
        _int, [_dp, _dp, _dp, _i64, _dp, _dp, _dp, _dp, _i64, _u32, _u32, _int, _dp, _u32p]),
    "hb200_prism_layer_gravity": (
        _int, [_dp, _dp, _dp, _i64, _dp, _i64, _dp, _i64, _dp, _dp, _dp, ctypes.c_double, _u32,
               _int, _dp, _u32p]),
    "hb200_point_gravity": (
        _int, [_dp, _dp, _dp, _i64, _dp, _dp, _dp, _dp, _i64, _u32, _int, _int, _dp, _u32p]),
    "hb200_eqs_predict": (_int, [_dp, _dp, _dp, _i64, _dp, _dp, _dp, _dp, _i64, _int, _dp, _u32p]),
    "hb200_eqs_predict_spherical": (
        _int, [_dp, _dp, _dp, _i64, _dp, _dp, _dp, _dp, _i64, _int, _dp, _u32p]),
    "hb200_dipole_magnetic": (
        _int, [_dp, _dp, _dp, _i64, _dp, _dp, _dp, _dp, _dp, _dp, _i64, _u32, _int, _dp, _u32p]),
    "hb200_eqs_jacobian": (_int, [_dp, _dp, _dp, _i64, _dp, _dp, _dp, _i64, _dp]),
    "hb200_eqs_jacobian_spherical": (_int, [_dp, _dp, _dp, _i64, _dp, _dp, _dp, _i64, _dp]),
    "hb200_eqs_fit": (
        _int, [_dp, _dp, _dp, _i64, _dp, _dp, _dp, _i64, _dp, _dp, ctypes.c_double, _int, _dp,
               ctypes.POINTER(_int)]),
    "hb200_eqs_fit_gb": (
        _int, [_dp, _dp, _dp, _i64, _dp, _dp, _dp, _i64, _dp, _dp, ctypes.c_double, _int, _i64,
               _i64p, _i64p, _i64p, _i64p, _dp, _dp]),
    "hb200_tesseroid_gravity": (
        _int, [_dp, _dp, _dp, _i64, _dp, _dp, _i64, _int, _int, _int, _dp, _u32p]),
    "hb200_tesseroid_gravity_variable_density": (
        _int, [_dp, _dp, _dp, _i64, _dp, _dp, _dp, _i64, _int, _int, _dp, _u32p]),
    "hb200_tesseroid_gravity_density_function": (
        _int, [_dp, _dp, _dp, _i64, _dp, _dp, _dp, _i64, _int, _vp, _vp, _dp, _u32p]),
    "hb200_tesseroid_inside_scan": (_int, [_dp, _dp, _dp, _i64, _dp, _i64, _u32p]),
    "hb200_tesseroid_ws_bytes": (_sz, [_i64, _i64]),
    "hb200_tesseroid_gravity_dev": (
        _int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _int, _int, _vp, _vp, _vp, _sz, _vp]),
    "hb200_prism_ws_bytes": (_sz, [_i64, _i64, _int]),
    "hb200_point_ws_bytes": (_sz, [_i64, _i64]),
    "hb200_prism_gravity_dev": (
        _int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _u32, _vp, _vp, _vp, _sz, _vp]),
    "hb200_prism_magnetic_dev": (
        _int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _u32, _u32, _vp, _vp, _vp, _sz, _vp]),
    "hb200_prism_layer_gravity_dev": (
        _int, [_vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp, ctypes.c_double, _u32,
               _vp, _vp, _vp, _sz, _vp]),
    "hb200_point_gravity_dev": (
        _int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _u32, _int, _int, _vp, _vp, _vp,
               _sz, _vp]),
    "hb200_fp64_peak": (_int, [_int, _dp, _dp]),
}  # fmt: skip

_LIB = None


class HarmonicaB200Error(RuntimeError):
    """Raised when libharmonica_b200.so reports an error."""


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise HarmonicaB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (harmonica_b200 has no CPU fallback)"
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _LIB = lib
    return _LIB


def check(rc):
    if rc == HB200_EZERODIV:
        raise ZeroDivisionError("division by zero")
    if rc != HB200_OK:
        msg = load().hb200_last_error()
        raise HarmonicaB200Error(f"libharmonica_b200 error {rc}: {msg.decode() if msg else ''}")


def init(devices=None):
    """Select the GPUs used by the host entry points (default: all visible)."""
    lib = load()
    if devices is None:
        env = os.environ.get("HARMONICA_B200_DEVICES")
        if env:
            devices = [int(d) for d in env.split(",") if d.strip()]
    if devices is None:
        check(lib.hb200_init(None, 0))
    else:
        arr = (_int * len(devices))(*devices)
        check(lib.hb200_init(arr, len(devices)))
    return lib.hb200_num_devices()


def ensure_init():
    lib = load()
    if lib.hb200_num_devices() == 0:
        init()
    return lib


def f64(a):
    """C-contiguous float64 view/copy (the drop-in casts everything to float64)."""
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    return a.ctypes.data_as(_dp)


def shard_mode(name):
    modes = {None: SHARD_AUTO, "auto": SHARD_AUTO, "observers": SHARD_OBSERVERS,
             "sources": SHARD_SOURCES}  # fmt: skip
    if name not in modes:
        raise ValueError(f"Invalid shard '{name}'. Choose one of 'auto', 'observers', 'sources'.")
    return modes[name]
