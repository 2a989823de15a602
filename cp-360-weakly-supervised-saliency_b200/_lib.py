"""ctypes binding of libcp360.so — exactly the entry points include/cp360.h declares.

No torch types cross this boundary: callers pass ``tensor.data_ptr()``, sizes and the raw
``cudaStream_t``. A missing library is a hard error (there is no CPU or PyTorch fallback).
"""
import ctypes
import os
import threading

from ._build import LIB_PATH

c_i32, c_i64, c_vp, c_dbl = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_double
p_i32 = ctypes.POINTER(ctypes.c_int)

# name -> (restype, argtypes); mirrors include/cp360.h one to one
SIGNATURES = {
    "cp360_version": (c_i32, []),
    "cp360_status_string": (ctypes.c_char_p, [c_i32]),
    "cp360_last_error": (ctypes.c_char_p, []),
    "cp360_launch_count": (ctypes.c_uint64, []),
    "cp360_cubepad_out_shape": (c_i32, [c_i32] * 6 + [p_i32, p_i32]),
    "cp360_cubepad_build_map": (c_i32, [c_i32] * 6 + [c_vp]),
    "cp360_cubepad_fwd": (c_i32, [c_vp, c_vp, c_i64, c_i64] + [c_i32] * 7 + [c_vp]),
    "cp360_cubepad_fwd_algo": (c_i32, [c_vp, c_vp, c_i64, c_i64] + [c_i32] * 8 + [c_vp]),
    "cp360_cubepad_pick_algo": (c_i32, [c_i64, c_i64] + [c_i32] * 8),
    "cp360_cubepad_fused_fwd": (c_i32, [c_vp, c_vp, c_i64, c_i64] + [c_i32] * 6 + [c_vp, c_vp, c_i32, c_i64, c_i64, c_vp]),
    "cp360_cubepad_autotune": (c_i32, [c_vp, c_vp, c_i64, c_i64] + [c_i32] * 7 + [c_vp]),
    "cp360_cubepad_set_tiling": (c_i32, [c_i64, c_i64] + [c_i32] * 14),
    "cp360_cubepad_tune_info": (c_i32, [c_i64, c_i64] + [c_i32] * 6 + [ctypes.c_char_p, c_i32]),
    "cp360_cubepad_build_inverse_map": (c_i32, [c_i32] * 6 + [c_vp, c_vp]),
    "cp360_cubepad_bwd_f32": (c_i32, [c_vp, c_vp, c_i64, c_i64] + [c_i32] * 6 + [c_vp]),
    "cp360_e2c_map_words": (c_i64, [c_i32, c_i32, c_i32]),
    "cp360_e2c_build_map": (c_i32, [c_i32, c_i32, c_i32, c_dbl, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cp360_e2c_fwd": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "cp360_e2c_fwd_u8": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, ctypes.c_float, c_vp, c_vp, c_vp]),
    "cp360_e2c_cubepad_fwd": (c_i32, [c_vp, c_i32, c_vp, c_vp, c_i64] + [c_i32] * 8 + [ctypes.c_float, c_vp, c_vp, c_vp]),
    "cp360_c2e_build_map": (c_i32, [c_i32, c_vp, c_vp]),
    "cp360_c2e_build_plan": (c_i32, [c_i32, c_i32, c_vp, c_vp, c_vp]),
    "cp360_c2e_fwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp]),
    "cp360_c2e_max_fwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp]),
    "cp360_c2e_max_arg_fwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp]),
    "cp360_c2e_build_bwd_plan": (c_i32, [c_i32, c_i32, c_vp, c_vp, c_vp]),
    "cp360_c2e_max_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp]),
    "cp360_c2e_build_cubic_plan": (c_i32, [c_i32, c_vp]),
    "cp360_c2e_cubic_fwd": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp]),
    "cp360_c2e_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_i64, c_i64, c_i32, c_vp]),
    "cp360_host_alloc": (c_i32, [ctypes.c_uint64, c_i32, ctypes.POINTER(c_vp)]),
    "cp360_host_free": (c_i32, [c_vp]),
    "cp360_npy_read_header": (c_i32, [ctypes.c_char_p, ctypes.c_char_p, c_i32, p_i32, c_vp, c_i32, c_vp, p_i32]),
    "cp360_npy_read_f32": (c_i32, [ctypes.c_char_p, c_vp, c_i64]),
    "cp360_npy_write_f32": (c_i32, [ctypes.c_char_p, c_vp, c_i32, c_vp]),
}

CP360_OK = 0
CP360_ERR_GROUP = 2
LAYOUT_NCHW, LAYOUT_NHWC = 0, 1
ALGO_AUTO, ALGO_GENERIC, ALGO_BAND_STG, ALGO_BAND_BULK, ALGO_CUBE, ALGO_ROW, ALGO_CUBE2 = range(7)

_lock = threading.Lock()
_lib = None


class CP360Error(RuntimeError):
    def __init__(self, status, what, detail):
        self.status = status
        super().__init__("libcp360: %s (status %d)%s" % (what, status, ": " + detail if detail else ""))


def lib():
    """Load libcp360.so once. Raises if it has not been built — never falls back."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        "libcp360.so is not built (%s). Run `python -c \"import __graft_entry__ as g; "
                        "g.build()\"` at the repo root. This package has no CPU/PyTorch fallback." % LIB_PATH)
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)     # AttributeError if the .so lacks a symbol
                    fn.restype, fn.argtypes = res, args
                _lib = handle
    return _lib


def check(status):
    if status != CP360_OK:
        handle = lib()
        what = handle.cp360_status_string(status).decode()
        detail = handle.cp360_last_error().decode()
        if status == CP360_ERR_GROUP:
            raise ValueError("CubePad size mismatch! " + detail)    # cube_pad.py:33-35
        raise CP360Error(status, what, detail)


def launch_count():
    return int(lib().cp360_launch_count())
