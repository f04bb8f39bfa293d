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


def test_struct_layout_matches_the_binding_documentation(tmp_path):
    """The C structs of include/sqt.h (compiled with gcc, offsetof), the ctypes mirror used by every test, and the byte
    offsets INTEGRATION.md gives the Haskell Storable instances must be the same numbers."""
    import ctypes as C
    import pysqt
    src = tmp_path / "layout.c"
    fields = {
        "sqt_scene_desc": ["root_bounds", "nodes", "n_nodes", "tris", "n_tris", "mats", "n_mats"],
        "sqt_camera": ["position", "rotation"],
        "sqt_render_params": ["rows", "cols", "xdiv", "ydiv", "seed_stride", "spp", "max_depth", "mode", "seed", "flags", "reserved"],
        "sqt_node": ["lmax", "rmin", "a", "b"],
        "sqt_tri": ["v0", "e1", "e2", "material", "orig_index", "pad"],
        "sqt_material": ["reflective", "surf_color", "emissive", "emit_color"],
        "sqt_stats": ["device_ms", "primary_ms", "paths_ms", "tonemap_ms", "reduce_ms", "h2d_ms", "d2h_ms", "rays_traced"],
    }
    body = "".join('printf("%s %%zu", sizeof(%s));%s printf("\\n");\n' % (
        s, s, "".join(' printf(" %%zu", offsetof(%s, %s));' % (s, f) for f in fs)) for s, fs in fields.items())
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sqt.h"\nint main(void){\n%s return 0; }\n' % body)
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(pysqt.ROOT, "include"), "-o", str(exe), str(src)])
    got = {l.split()[0]: [int(x) for x in l.split()[1:]] for l in subprocess.check_output([str(exe)], text=True).splitlines()}
    mirror = {"sqt_scene_desc": pysqt.SceneDesc, "sqt_camera": pysqt.Camera, "sqt_render_params": pysqt.RenderParams, "sqt_node": pysqt.Node,
              "sqt_tri": pysqt.Tri, "sqt_material": pysqt.Material, "sqt_stats": pysqt.Stats}
    for s, fs in fields.items():
        T = mirror[s]
        assert got[s] == [C.sizeof(T)] + [getattr(T, f).offset for f in fs], s
    # the numbers written into INTEGRATION.md section 2
    assert got["sqt_scene_desc"] == [72, 0, 24, 32, 40, 48, 56, 64]
    assert got["sqt_camera"] == [48, 0, 12]
    assert got["sqt_render_params"] == [48, 0, 4, 8, 12, 16, 20, 24, 28, 32, 40, 44]
    assert got["sqt_node"][0] == 16 and got["sqt_tri"] == [48, 0, 12, 24, 36, 40, 44] and got["sqt_material"][0] == 32
    assert got["sqt_stats"][0] == 168
