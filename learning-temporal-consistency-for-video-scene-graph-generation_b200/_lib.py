"""ctypes binding of the C-ABI in include/b200vsgg.h.  No CPU fallback: if the shared library is
missing, ``lib()`` raises; if a call returns non-zero, ``check()`` raises RuntimeError."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200vsgg.so")

_lib = None


class GemmEpilogue(C.Structure):
    """Mirror of struct b200vsgg_gemm_epilogue (include/b200vsgg.h)."""
    _fields_ = [
        ("bias", C.c_void_p),
        ("residual", C.c_void_p),
        ("residual_is_bf16", C.c_int32),
        ("ldr", C.c_int32),
        ("mask_src", C.c_void_p),
        ("ldm", C.c_int32),
        ("mask_mode", C.c_int32),
        ("act", C.c_int32),
        ("out_f32", C.c_void_p),
        ("ld_f32", C.c_int32),
        ("out_bf16", C.c_void_p),
        ("ld_bf16", C.c_int32),
        ("accumulate", C.c_int32),
        ("alpha", C.c_float),
        ("dropout_p", C.c_float),
        ("dropout_seed", C.c_uint64),
        ("split_k", C.c_int32),
        ("a_k_period", C.c_int32),
    ]


def _declare(lib):
    i32, vp, f32, u64 = C.c_int32, C.c_void_p, C.c_float, C.c_uint64
    lib.b200vsgg_version.restype = C.c_char_p
    lib.b200vsgg_version.argtypes = []
    lib.b200vsgg_last_error.restype = C.c_char_p
    lib.b200vsgg_last_error.argtypes = []
    lib.b200vsgg_gemm_bf16.restype = i32
    lib.b200vsgg_gemm_bf16.argtypes = [vp, i32, i32, vp, i32, i32, i32, i32, i32, C.POINTER(GemmEpilogue), vp]
    from . import _decls
    _decls.declare(lib)


def lib():
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "b200vsgg: %s not found. Build it with `python -m b200vsgg.build` "
                "(there is no CPU fallback for the hot path)." % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().b200vsgg_last_error().decode()
        raise RuntimeError("b200vsgg %s failed (code %d): %s" % (what, rc, msg))
