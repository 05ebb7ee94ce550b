"""ctypes binding of the TEST-ONLY library libsupcon_b200_test.so (csrc/supcon_debug.h): the product
library plus the tcgen05 one-tile diagnostic and the host-side plan introspection.  Loaded by tests/ and
tools/ only; the product package never touches it."""
import ctypes
import os
from ctypes import POINTER, c_int32, c_int64, c_void_p

from wav2vec_contr_loss_b200 import _cabi
from wav2vec_contr_loss_b200 import build as _build

_lib = None


def load():
    global _lib
    if _lib is None:
        _build.build()
        lib = ctypes.CDLL(_build.TEST_LIB_PATH)
        P = POINTER(_cabi.Problem)
        lib.supcon_debug_last_error.restype = ctypes.c_char_p
        lib.supcon_debug_tc_tile.restype = c_int32
        lib.supcon_debug_tc_tile.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
        lib.supcon_debug_plan.restype = c_int32
        lib.supcon_debug_plan.argtypes = [P, c_void_p, c_int32]
        lib.supcon_debug_sched.restype = c_int32
        lib.supcon_debug_sched.argtypes = [c_int32, c_int32, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                           c_void_p]
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {load().supcon_debug_last_error().decode()}")
