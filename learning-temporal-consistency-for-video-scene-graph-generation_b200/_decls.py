"""ctypes signatures for the non-GEMM entry points of include/b200vsgg.h (kept next to the header
order so the symbol-export test can diff them)."""
import ctypes as C

i32, i64, vp, f32, u64 = C.c_int32, C.c_int64, C.c_void_p, C.c_float, C.c_uint64

# name -> argtypes (restype is always int32)
SIGNATURES = {}


def declare(lib):
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = i32
        fn.argtypes = argtypes
