"""Shared helpers for the tests: the host build of the kernel logic (tests/emu), ray generators, scene builders."""
import ctypes as C
import os

import numpy as np

import pysqt
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
_EMU = None


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def emu_lib():
    global _EMU
    if _EMU is None:
        E = C.CDLL(os.path.join(HERE, "emu", "libsqt_emu.so"))
        E.emu_upload.restype = C.c_void_p
        E.emu_upload.argtypes = [C.c_void_p]
        E.emu_error.restype = C.c_char_p
        E.emu_error.argtypes = [C.c_void_p]
        E.emu_free.argtypes = [C.c_void_p]
        E.emu_height.restype = C.c_int
        E.emu_height.argtypes = [C.c_void_p]
        E.emu_set_leaf_cull.argtypes = [C.c_void_p, C.c_int]
        E.emu_set_spheres.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
        E.emu_intersect_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        E.emu_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _EMU = E
    return _EMU


class Emu:
    """Host build of the device code paths (same source as the kernels), for CPU-side parity tests."""

    def __init__(self, host_scene):
        self.scene = host_scene
        d = host_scene.desc()
        self.h = emu_lib().emu_upload(C.byref(d))
        err = emu_lib().emu_error(self.h)
        if err:
            raise RuntimeError(err.decode())

    def __del__(self):
        try:
            emu_lib().emu_free(self.h)
        except Exception:
            pass

    def intersect_batch(self, org, dirs):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = len(org)
        tri = np.zeros(n, np.int32); dist = np.zeros(n, np.float32); pt = np.zeros((n, 3), np.float32)
        cn = np.zeros(5, np.uint64)
        emu_lib().emu_intersect_batch(self.h, _p(org), _p(dirs), n, _p(tri), _p(dist), _p(pt), _p(cn))
        return tri, dist, pt, dict(branch_visits=int(cn[0]), child_box_tests=int(cn[1]), tri_tests=int(cn[2]), rays=int(cn[3]),
                                   leaves_culled=int(cn[4]))

    def intersect_batch_interleaved(self, org, dirs):
        """64 rays at a time stepped round-robin with their stacks interleaved like the pool kernel's (see emu.cpp)."""
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = len(org)
        tri = np.zeros(n, np.int32); dist = np.zeros(n, np.float32)
        E = emu_lib()
        E.emu_intersect_batch_interleaved.restype = C.c_int
        bad = E.emu_intersect_batch_interleaved(self.h, _p(org), _p(dirs), C.c_longlong(n), _p(tri), _p(dist))
        return tri, dist, bad

    def set_leaf_cull(self, on):
        emu_lib().emu_set_leaf_cull(self.h, 1 if on else 0)

    def set_spheres(self, spheres):
        a = np.zeros(len(spheres), pysqt.SPHERE_DT)
        for i, (cx, cy, cz, r, m) in enumerate(spheres):
            a[i]["center"] = (cx, cy, cz); a[i]["radius"] = r; a[i]["material"] = int(m)
        self._spheres = a
        emu_lib().emu_set_spheres(self.h, _p(a) if len(a) else None, len(a))

    def render(self, cam12, params, rank=0, world=1):
        acc = np.zeros((params.rows, params.cols, 3), np.float32)
        rgb = np.zeros((params.rows, params.cols, 3), np.uint8)
        st = np.zeros(5, np.uint64)
        cam = pysqt.camera_struct(cam12)
        emu_lib().emu_render(self.h, C.byref(cam), C.byref(params), rank, world, _p(acc), _p(rgb), _p(st))
        return dict(accum=acc, rgb8=rgb, rays=int(st[0]), samples=int(st[1]), primary_reused=int(st[2]),
                    branch_visits=int(st[3]), tri_tests=int(st[4]))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_same_hits(got, want, what=""):
    tri, dist, point = got[:3]
    otri, odist, opoint = want[:3]
    bad = np.nonzero(tri != otri)[0]
    assert len(bad) == 0, "%s: %d/%d hit indices differ, first at ray %d: got %d want %d" % (
        what, len(bad), len(tri), bad[0] if len(bad) else -1, tri[bad[0]] if len(bad) else 0, otri[bad[0]] if len(bad) else 0)
    assert np.array_equal(bits(dist), bits(odist)), what + ": dist bits differ"
    assert np.array_equal(bits(point), bits(opoint)), what + ": point bits differ"


def random_rays(n, seed, lo=-2.5, hi=2.5):
    rng = np.random.default_rng(seed)
    org = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    return org, d


def adversarial_rays(v9, seed=5, n_each=256):
    """Axis-aligned directions (zero components -> inf reciprocals), origins on box planes (0*inf = NaN in the
    slab test), rays through shared vertices and edge midpoints, denormal / huge direction components."""
    rng = np.random.default_rng(seed)
    v9 = np.asarray(v9, np.float32).reshape(-1, 3, 3)
    lo, hi = v9.reshape(-1, 3).min(0), v9.reshape(-1, 3).max(0)
    org, dirs = [], []
    # 1. axis-aligned rays from outside and inside
    for ax in range(3):
        for sgn in (1.0, -1.0):
            o = rng.uniform(lo, hi, (n_each, 3)).astype(np.float32)
            d = np.zeros((n_each, 3), np.float32); d[:, ax] = sgn
            org.append(o); dirs.append(d)
            o2 = o.copy(); o2[:, ax] = (lo[ax] - 1) if sgn > 0 else (hi[ax] + 1)
            org.append(o2); dirs.append(d)
    # 2. origin exactly on the root box planes with a zero direction component on that axis (NaN slabs)
    for ax in range(3):
        o = rng.uniform(lo, hi, (n_each, 3)).astype(np.float32)
        o[: n_each // 2, ax] = lo[ax]; o[n_each // 2:, ax] = hi[ax]
        d = rng.normal(size=(n_each, 3)).astype(np.float32); d[:, ax] = 0.0
        org.append(o); dirs.append(d)
        d2 = d.copy(); d2[:, ax] = -0.0
        org.append(o); dirs.append(d2)
    # 3. rays aimed exactly at vertices and edge midpoints (ties between neighbouring triangles)
    pick = rng.integers(0, len(v9), n_each)
    eye = np.array([0.0, 7.0, 0.75], np.float32)
    tv = v9[pick, rng.integers(0, 3, n_each)]
    org.append(np.repeat(eye[None], n_each, 0)); dirs.append((tv - eye).astype(np.float32))
    mid = ((v9[pick, 0] + v9[pick, 1]) * np.float32(0.5)).astype(np.float32)
    org.append(np.repeat(eye[None], n_each, 0)); dirs.append((mid - eye).astype(np.float32))
    # 4. tiny / huge / denormal direction components
    o = rng.uniform(lo, hi, (n_each, 3)).astype(np.float32)
    d = rng.normal(size=(n_each, 3)).astype(np.float32)
    d[: n_each // 4, 0] = np.float32(1e-42); d[n_each // 4: n_each // 2, 1] = np.float32(-1e-39)
    d[n_each // 2: 3 * n_each // 4] *= np.float32(1e18); d[3 * n_each // 4:] *= np.float32(1e-18)
    org.append(o); dirs.append(d)
    # 5. origins on triangle vertices (self hits rejected by t > eps)
    org.append(tv.astype(np.float32)); dirs.append(rng.normal(size=(n_each, 3)).astype(np.float32))
    return np.concatenate(org), np.concatenate(dirs)


def tame_boundary_rays(v9, n, seed=404):
    """Rays around the two limits of the `tame` condition of the child slabs: (a) origins on 1-norm shells at
    0.9 .. 1.1 of twice the root's half extent, |d|_1 between 1.9 and 2.1; (b) nearly tangential rays from inside the box;
    (c) direction lengths over six decades.  3n rays."""
    rng = np.random.default_rng(seed)
    pts = np.asarray(v9, np.float64).reshape(-1, 3)
    lo, hi = pts.min(0), pts.max(0)
    c, h = 0.5 * (lo + hi), 0.5 * (hi - lo)
    u = rng.normal(size=(n, 3)); u /= np.abs(u).sum(1, keepdims=True)
    org_a = c + u * (2.0 * h.sum()) * rng.uniform(0.9, 1.1, (n, 1))
    tgt = c + rng.uniform(-1, 1, (n, 3)) * h
    d_a = tgt - org_a
    d_a /= np.abs(d_a).sum(1, keepdims=True)
    d_a *= rng.uniform(1.9, 2.1, (n, 1))
    org_b = c + rng.uniform(-1, 1, (n, 3)) * h
    d_b = rng.normal(size=(n, 3)); d_b[:, int(np.argmin(h))] *= 0.02
    d_b /= np.linalg.norm(d_b, axis=1, keepdims=True)
    org_c = c + rng.uniform(-1.5, 1.5, (n, 3)) * h
    d_c = rng.normal(size=(n, 3)); d_c /= np.linalg.norm(d_c, axis=1, keepdims=True)
    d_c *= 10.0 ** rng.uniform(-3, 3, (n, 1))
    return (np.concatenate([org_a, org_b, org_c]).astype(np.float32), np.concatenate([d_a, d_b, d_c]).astype(np.float32))


def build_pair(v9, mat_idx, mats8):
    """(oracle scene with BIH, host scene) from the same triangle arrays."""
    osc = O.Scene.from_arrays(v9, mat_idx, mats8)
    osc.make_bih()
    hs = pysqt.HostScene.from_arrays(v9, mat_idx, mats8)
    return osc, hs


def default_camera():
    return pysqt.load_camera(os.path.join(pysqt.ROOT, "data", "camera"))
