"""ctypes binding of include/amrvw_b200.h.  There is no fallback: if the CUDA library is missing
or no device is present, loading/creating fails loudly."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))


class AmrvwError(RuntimeError):
    pass


class Options(ctypes.Structure):
    _fields_ = [("schedule", ctypes.c_int32), ("lanes_per_poly", ctypes.c_int32), ("shift_reuse", ctypes.c_int32),
                ("small_window", ctypes.c_int32), ("shift_solver", ctypes.c_int32), ("it_count", ctypes.c_int32), ("seed", ctypes.c_uint64),
                ("it_max_factor", ctypes.c_int32), ("stream_ctas_per_sm", ctypes.c_int32), ("tol", ctypes.c_double)]


class Stats(ctypes.Structure):
    _fields_ = [("turnovers", ctypes.c_int64), ("sweeps", ctypes.c_int64), ("exceptional", ctypes.c_int64),
                ("unconverged", ctypes.c_int64), ("fat_steps", ctypes.c_int64), ("launches", ctypes.c_int64),
                ("device_ms", ctypes.c_double), ("lanes", ctypes.c_int64), ("schedule_used", ctypes.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


SCHEDULE_AUTO, SCHEDULE_REFERENCE, SCHEDULE_PIPELINED, SCHEDULE_STREAMED = 0, 1, 2, 3
SHIFTS_AUTO, SHIFTS_CHASE, SHIFTS_DENSE = 0, 1, 2

_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_SIGS = {
    "amrvw_version": (ctypes.c_int, []),
    "amrvw_create": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.c_int, ctypes.c_uint64]),
    "amrvw_create_multi": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_uint64]),
    "amrvw_device_count": (ctypes.c_int, [_P]),
    "amrvw_partition_batch": (ctypes.c_int, [_P, _I64, ctypes.c_int, _P]),
    "amrvw_partition_hessenberg": (ctypes.c_int, [_P, _I64, ctypes.c_int, _P]),
    "amrvw_destroy": (ctypes.c_int, [_P]),
    "amrvw_last_error": (ctypes.c_char_p, [_P]),
    "amrvw_default_options": (None, [ctypes.POINTER(Options)]),
    "amrvw_set_options": (ctypes.c_int, [_P, ctypes.POINTER(Options)]),
    "amrvw_set_stream": (ctypes.c_int, [_P, _P]),
    "amrvw_use_own_stream": (ctypes.c_int, [_P]),
    "amrvw_roots_rds_f64": (ctypes.c_int, [_P, _P, _P, _I64, _P, _P, _P, ctypes.POINTER(Stats)]),
    "amrvw_roots_css_c64": (ctypes.c_int, [_P, _P, _P, _I64, _P, _P, _P, ctypes.POINTER(Stats)]),
    "amrvw_roots_pencil_f64": (ctypes.c_int, [_P, _P, _P, _P, _I64, _P, _P, _P, ctypes.POINTER(Stats)]),
    "amrvw_roots_pencil_c64": (ctypes.c_int, [_P, _P, _P, _P, _I64, _P, _P, _P, ctypes.POINTER(Stats)]),
    "amrvw_count_real_roots_rds_f64": (ctypes.c_int, [_P, _P, _P, _I64, _P, _P, ctypes.POINTER(Stats)]),
    "amrvw_eigvals_hessenberg_f64": (ctypes.c_int, [_P, _P, _P, _I64, ctypes.c_int, _P, _P, ctypes.POINTER(Stats)]),
    "amrvw_eigvals_hessenberg_c64": (ctypes.c_int, [_P, _P, _P, _I64, ctypes.c_int, _P, _P, ctypes.POINTER(Stats)]),
    "amrvw_eigvals_hessenberg_f64_dev": (ctypes.c_int, [_P, _P, _P, _P, _P, _I64, ctypes.c_int32, ctypes.c_int, _P, _P, ctypes.POINTER(Stats)]),
    "amrvw_eigvals_hessenberg_c64_dev": (ctypes.c_int, [_P, _P, _P, _P, _P, _I64, ctypes.c_int32, ctypes.c_int, _P, _P, ctypes.POINTER(Stats)]),
    "amrvw_roots_rds_f64_dev": (ctypes.c_int, [_P, _P, _P, _I64, _I64, _I64, _P, _P, _P, ctypes.POINTER(Stats)]),
    "amrvw_roots_css_c64_dev": (ctypes.c_int, [_P, _P, _P, _I64, _I64, _I64, _P, _P, _P, ctypes.POINTER(Stats)]),
    "amrvw_roots_pencil_f64_dev": (ctypes.c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _P, _P, _P, ctypes.POINTER(Stats)]),
    "amrvw_roots_pencil_c64_dev": (ctypes.c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _P, _P, _P, ctypes.POINTER(Stats)]),
    "amrvw_count_real_roots_dev": (ctypes.c_int, [_P, _P, _P, _I64, ctypes.c_int32, _P, _P]),
    "amrvw_measure_fp64_peak": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_double)]),
    "amrvw_selftest": (ctypes.c_int, [_P, _I64, ctypes.POINTER(ctypes.c_double)]),
    "amrvw_last_profile": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int64)]),
    "amrvw_debug_counters": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int64)]),
    "amrvw_selftest_complex": (ctypes.c_int, [_P, _I64, ctypes.POINTER(ctypes.c_double)]),
    "amrvw_selftest_robust": (ctypes.c_int, [_P, _I64, ctypes.POINTER(ctypes.c_double)]),
}
EXPORTS = tuple(_SIGS)

_libs = {}


def lib_path(strict=False):
    return os.path.join(HERE, "libamrvw_b200_strict.so" if strict else "libamrvw_b200.so")


def load(strict=False, path=None):
    """dlopen the in-tree library and type its exports.  Raises AmrvwError if it was not built.
    `path` (tools/ only) loads another build of the same ABI, e.g. a kernel variant under build/."""
    key = path or strict
    if key in _libs:
        return _libs[key]
    path = path or lib_path(strict)
    if not os.path.exists(path):
        raise AmrvwError(f"{path} is missing: run `python __graft_entry__.py build` (there is no CPU fallback)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _libs[key] = lib
    return lib
