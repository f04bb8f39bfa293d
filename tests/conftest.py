import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "squigly-trace_b200")
for p in (ROOT, PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def _have_gpu():
    try:
        import pysqt
        c = pysqt.Context(0)
        c.close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a B200 must fail loudly, not skip: only auto-skip when the user did not ask for gpu.
    if "gpu" in (config.getoption("-m") or ""):
        return
    skip = pytest.mark.skip(reason="needs a B200")
    if not _have_gpu():
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def built():
    """Build whatever is missing (the GPU box receives the prebuilt .so files with the snapshot)."""
    need = [os.path.join(PKG, "libsqt_b200.so"), os.path.join(PKG, "libsqt_host.so"),
            os.path.join(ROOT, "oracle", "liboracle.so"), os.path.join(ROOT, "tests", "emu", "libsqt_emu.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()
    return True


DATA = os.path.join(ROOT, "data")
OBJ = os.path.join(DATA, "scene.obj")
CAMERA = os.path.join(DATA, "camera")


@pytest.fixture(scope="session")
def oracle_scene(built):
    from oracle import oracle as O
    s = O.Scene.load(OBJ, DATA)
    s.make_bih()
    return s


@pytest.fixture(scope="session")
def host_scene(built):
    import pysqt
    return pysqt.HostScene.load(OBJ, DATA)


@pytest.fixture(scope="session")
def camera(built):
    import pysqt
    return pysqt.load_camera(CAMERA)


@pytest.fixture(scope="session")
def gpu_ctx(built):
    import pysqt
    c = pysqt.Context(0)
    yield c
    c.close()
