"""CPU: the C-ABI library builds/loads and exports every symbol the header declares (no compute)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from vision_assist_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "vision_assist_b200.h")).read()
    declared = set(re.findall(r"VA_API\s+[\w\s\*]+?\b(va_\w+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.va_abi_version() == 2


def test_struct_sizes_match_header():
    from vision_assist_b200 import _lib
    assert C.sizeof(_lib.VaConfig) == 40
    assert C.sizeof(_lib.VaLayout) == 64
    assert C.sizeof(_lib.VaGridInput) == 32


def test_no_cpu_fallback_without_gpu():
    import torch
    from vision_assist_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    ctx = C.c_void_p()
    cfg = _lib.VaConfig(0, 640, 640, 160, 160, 32, 8, 20, 4, 0)
    assert lib.va_create(C.byref(ctx), C.byref(cfg)) == _lib.VA_ERR_CUDA
    assert b"no CPU fallback" in lib.va_last_error(None)
    from vision_assist_b200.engine import MaskGridEngine
    with pytest.raises(RuntimeError):
        MaskGridEngine(H=640, W=640, mh=160, mw=160)


def test_invalid_config_rejected():
    from vision_assist_b200 import _lib
    lib = _lib.load()
    ctx = C.c_void_p()
    for bad in (dict(K=16), dict(max_n=33), dict(mw=162), dict(gs=2), dict(max_batch=0)):
        kw = dict(device=0, H=640, W=640, mh=160, mw=160, K=32, max_n=8, gs=20, max_batch=4, flags=0)
        kw.update(bad)
        cfg = _lib.VaConfig(*[kw[k] for k, _ in _lib.VaConfig._fields_])
        assert lib.va_create(C.byref(ctx), C.byref(cfg)) == _lib.VA_ERR_INVALID, bad


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "vision_assist_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_integration_md_binding_matches_the_library():
    """The reference-side ctypes stub shown in INTEGRATION.md declares the same structures (field order, sizes) as
    the product's own binding and only names entry points the library exports."""
    from vision_assist_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = text[text.index("```python\n# vision_assist/_va_sm100.py"):]
    code = code[len("```python\n"):code.index("\n```")]
    # the two structure definitions, executed as written
    decl = code[code.index("class VaConfig"):code.index("vp = C.c_void_p")]
    ns = {"C": C}
    exec(decl, ns)
    assert [f[0] for f in ns["VaConfig"]._fields_] == [f[0] for f in _lib.VaConfig._fields_]
    assert C.sizeof(ns["VaLayout"]) == C.sizeof(_lib.VaLayout) == 64
    mine = [f[0] for f in _lib.VaLayout._fields_]
    theirs = [f[0] for f in ns["VaLayout"]._fields_]
    assert [t if t != "alg_bytes_n1" else "algorithmic_bytes_per_frame_n1" for t in theirs] == mine
    for name in set(re.findall(r"lib\.(va_\w+)", code)):
        assert name in _lib.EXPORTS, name
