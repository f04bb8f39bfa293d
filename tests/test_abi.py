"""The C-ABI shared library: loads without a GPU, exports every symbol include/sqt.h declares, and refuses to
compute (no fallback) when there is no B200.  CPU only; no compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

import pysqt

HEADER = os.path.join(pysqt.ROOT, "include", "sqt.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sqt_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = pysqt.b200()
    names = declared_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), "include/sqt.h declares %s but libsqt_b200.so does not export it" % n
    assert sorted(pysqt.ABI_SYMBOLS) == names
    assert lib.sqt_abi_version() == 1


def test_exports_are_plain_c():
    out = subprocess.run(["nm", "-D", "--defined-only", pysqt.LIB_B200], capture_output=True, text=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    ours = [s for s in syms if s.startswith("sqt_")]
    assert sorted(ours) == declared_functions()


def test_struct_sizes_match_header():
    assert C.sizeof(pysqt.Node) == 16 and C.sizeof(pysqt.Tri) == 48 and C.sizeof(pysqt.Material) == 32
    assert C.sizeof(pysqt.Camera) == 48 and C.sizeof(pysqt.RenderParams) == 48


def test_library_is_built_for_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", pysqt.LIB_B200], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_device():
    """On a box without a B200 sqt_create must fail with SQT_E_NO_DEVICE; with one it must succeed."""
    lib = pysqt.b200()
    h = C.c_void_p()
    rc = lib.sqt_create(0, C.byref(h))
    if rc == 0:
        lib.sqt_destroy(h)
        pytest.skip("a CUDA device is present")
    assert rc == 3
    msg = lib.sqt_last_error(None).decode()
    assert "no CPU fallback" in msg or "sm_100a" in msg
    with pytest.raises(pysqt.SqtError):
        pysqt.Context(0)


def test_product_does_not_reference_the_oracle():
    """The product path must not import, link or execute anything under oracle/ or tests/emu."""
    for root, _, files in os.walk(os.path.join(pysqt.ROOT, "squigly-trace_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "liboracle" not in text and "libsqt_emu" not in text and "import oracle" not in text and "from oracle" not in text, f
    for lib in (pysqt.LIB_B200, pysqt.LIB_HOST):
        ldd = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
        assert "oracle" not in ldd and "emu" not in ldd
