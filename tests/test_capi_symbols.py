"""The C-ABI library loads without a GPU and exports every function include/b200vsgg.h declares;
the ctypes table (_decls.SIGNATURES + _lib) lists exactly the same set (no compute calls here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200vsgg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200vsgg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from b200vsgg import _lib, _decls, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), "header declares %s but the library does not export it" % n
    bound = set(_decls.SIGNATURES) | {"b200vsgg_version", "b200vsgg_last_error", "b200vsgg_gemm_bf16"}
    assert bound == set(names), (bound ^ set(names))
    assert lib.b200vsgg_version().decode().startswith("b200vsgg")


def test_bad_arguments_return_error_codes_without_gpu():
    import ctypes as C
    from b200vsgg import _lib
    lib = _lib.lib()
    ep = _lib.GemmEpilogue()
    rc = lib.b200vsgg_gemm_bf16(None, 0, 0, None, 0, 0, 0, 0, 0, C.byref(ep), None)
    assert rc == -1 and b"gemm" in lib.b200vsgg_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "gemm_bf16")


def test_model_refuses_cpu_tensors():
    """No CPU fallback: the product path must fail loudly off-GPU."""
    import torch
    from b200vsgg import synthetic, tempura
    m = tempura.TEMPURA(mode="predcls", attention_class_num=3, spatial_class_num=6, contact_class_num=17,
                        obj_classes=synthetic.ag_object_classes(), enc_layer_num=1, dec_layer_num=1, K=2,
                        rel_mem_compute=None)
    e = synthetic.make_video_entry(0, 3, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(e, phase="test")
