"""ctypes binding of libphox.so - the C ABI declared in include/phox.h.

This is the stub a Python-side maintainer of the reference would write.
There is deliberately no fallback: if the CUDA library is missing or no device is usable the
calls raise, they never route to a CPU implementation.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("PHOX_LIB") or os.path.join(CSRC, "libphox.so")   # PHOX_LIB: tuning builds only

PHOX_OK = 0
MODE_MINIMAL, MODE_HITPHOTON, MODE_HITPHOTONSEQ, MODE_DEBUGLITE, MODE_DEBUGHEAVY = range(5)
RNG_PRODUCTION, RNG_DEBUG_TAG = 0, 1
ACCEL_BVH, ACCEL_BRUTE, ACCEL_BVH_NOHOME = 0, 1, 2
KERNEL_AUTO, KERNEL_PERSISTENT, KERNEL_WAVEFRONT = 0, 1, 2


class PhoxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("phox error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    """phox_config (include/phox.h); defaults are SEventConfig's (sysrap/SEventConfig.cc:37-115)."""
    _fields_ = [
        ("max_bounce", C.c_int32), ("event_mode", C.c_int32), ("max_record", C.c_int32),
        ("rng_mode", C.c_int32), ("accel", C.c_int32),
        ("hit_mask", C.c_uint32), ("epsilon0_mask", C.c_uint32), ("propagate_refine", C.c_uint32),
        ("propagate_epsilon", C.c_float), ("propagate_epsilon0", C.c_float),
        ("refine_distance", C.c_float), ("tmax", C.c_float), ("max_time", C.c_float),
        ("kernel_mode", C.c_uint32),
        ("rng_seed", C.c_uint64), ("rng_offset", C.c_uint64), ("skipahead_event_offset", C.c_uint64),
        ("max_slot", C.c_int64),
        ("mode_lite", C.c_uint32), ("reserved0", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("num_photon", C.c_uint64), ("num_hit", C.c_uint64), ("num_ray", C.c_uint64),
        ("num_launch", C.c_uint64), ("num_kernel", C.c_uint64),
        ("launch_seconds", C.c_double), ("upload_seconds", C.c_double), ("gather_seconds", C.c_double),
        ("simulate_kernel_seconds", C.c_double), ("compact_kernel_seconds", C.c_double),
        ("trace_kernel_seconds", C.c_double), ("propagate_kernel_seconds", C.c_double), ("num_trace_launch", C.c_uint64),
        ("num_home_ray", C.c_uint64),
    ]


# every symbol include/phox.h declares, with its signature
SYMBOLS = {
    "phox_default_config": (None, [C.POINTER(Config)]),
    "phox_create": (C.c_void_p, [C.c_int]),
    "phox_destroy": (None, [C.c_void_p]),
    "phox_last_error": (C.c_char_p, [C.c_void_p]),
    "phox_desc": (C.c_char_p, [C.c_void_p]),
    "phox_set_geometry": (C.c_int, [C.c_void_p] + [C.c_void_p, C.c_int64] * 6),
    "phox_set_tables": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_float,
                                  C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32]),
    "phox_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "phox_set_config": (C.c_int, [C.c_void_p, C.POINTER(Config)]),
    "phox_get_config": (C.c_int, [C.c_void_p, C.POINTER(Config)]),
    "phox_simulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                C.c_uint64, C.POINTER(C.c_double)]),
    "phox_simulate_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                       C.c_uint64, C.POINTER(C.c_double)]),
    "phox_num_photon": (C.c_int64, [C.c_void_p]),
    "phox_num_hit": (C.c_int64, [C.c_void_p]),
    "phox_get_hits": (C.c_int, [C.c_void_p, C.c_void_p]),
    "phox_hits_device": (C.c_void_p, [C.c_void_p]),
    "phox_get_hits_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "phox_get_hits_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "phox_hits_wait": (C.c_int, [C.c_void_p]),
    "phox_host_alloc": (C.c_void_p, [C.c_int64]),
    "phox_host_free": (None, [C.c_void_p]),
    "phox_device_count": (C.c_int, []),
    "phox_get_array": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "phox_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "phox_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "phox_reset": (None, [C.c_void_p]),
    "phox_intersect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]),
    "phox_simtrace": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "phox_merge_hits": (C.c_int64, [C.c_void_p, C.c_float, C.c_void_p, C.c_int64]),
    "phox_merge": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_float, C.c_void_p, C.c_int64]),
    "phox_get_hits_lite": (C.c_int, [C.c_void_p, C.c_void_p]),
    "phox_merge_hits_lite": (C.c_int64, [C.c_void_p, C.c_float, C.c_void_p, C.c_int64]),
    "phox_boundary_lookup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "phox_rng_sequence": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_int32]),
}

_lib = None


def build(verbose=False):
    """Compile libphox.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building libphox.so failed")
    return LIB_PATH


def load():
    """Load libphox.so and bind every declared symbol; raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def default_config():
    cfg = Config()
    load().phox_default_config(C.byref(cfg))
    return cfg
