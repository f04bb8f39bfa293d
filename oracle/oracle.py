"""ctypes binding of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of oracle.c.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, never by the
product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_load_obj.restype = C.c_void_p
        L.orc_load_obj.argtypes = [C.c_char_p, C.c_char_p]
        L.orc_scene_from_arrays.restype = C.c_void_p
        L.orc_scene_from_arrays.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_free_scene.argtypes = [C.c_void_p]
        L.orc_error.restype = C.c_char_p
        L.orc_error.argtypes = [C.c_void_p]
        for f in ("orc_n_tris", "orc_n_mats", "orc_n_nodes", "orc_height", "orc_longest_leaf", "orc_n_leaves", "orc_make_bih"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_set_spheres.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_get_tris.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_mats.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_export_bih.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_load_camera.restype = C.c_int
        L.orc_load_camera.argtypes = [C.c_char_p, C.c_void_p]
        L.orc_rot_matrix_rads.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_intersect_batch.restype = C.c_int
        L.orc_intersect_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_make_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_render.restype = C.c_int
        L.orc_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_render_window.restype = C.c_int
        L.orc_render_window.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        L.orc_tone_map.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]
        L.orc_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_random_r01.restype = C.c_float
        L.orc_random_r01.argtypes = [C.c_uint32]
        L.orc_draw_word.restype = C.c_uint32
        L.orc_draw_word.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.orc_sqt_sincos.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
        L.orc_sqt_acos.restype = C.c_float
        L.orc_sqt_acos.argtypes = [C.c_float]
        L.orc_sqt_atan.restype = C.c_float
        L.orc_sqt_atan.argtypes = [C.c_float]
        L.orc_intersects_bb.restype = C.c_int
        L.orc_intersects_bb.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_moller_trumbore.restype = C.c_int
        L.orc_moller_trumbore.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


class Params(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("xdiv", C.c_int32), ("ydiv", C.c_int32),
                ("seed_stride", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32), ("mode", C.c_int32),
                ("trig", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("split_samples", C.c_int32),
                ("seed", C.c_uint64)]


def make_params(width, height, spp, max_depth=3, seed=0, mode=0, trig=1, literal=False, rank=0, world=1,
                split_samples=False):
    """literal=True reproduces Lib.hs:69-85 for `-d width,height` verbatim: the array has
    `width` rows and `height` columns.  literal=False is the corrected mapping (rows=height, cols=width)."""
    if literal:
        return Params(width, height, width, height, width, spp, max_depth, mode, trig, rank, world, int(split_samples), seed)
    return Params(height, width, width, height, width, spp, max_depth, mode, trig, rank, world, int(split_samples), seed)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Scene:
    """Owns an orc_scene*."""

    def __init__(self, handle):
        self.h = handle
        err = lib().orc_error(self.h)
        if err:
            raise RuntimeError(err.decode())

    @classmethod
    def load(cls, obj_path, data_dir):
        return cls(lib().orc_load_obj(obj_path.encode(), data_dir.encode()))

    @classmethod
    def from_arrays(cls, v9, mat_idx, mats8):
        v9 = np.ascontiguousarray(v9, np.float32).reshape(-1, 9)
        mat_idx = np.ascontiguousarray(mat_idx, np.int32)
        mats8 = np.ascontiguousarray(mats8, np.float32).reshape(-1, 8)
        return cls(lib().orc_scene_from_arrays(_p(v9), _p(mat_idx), len(v9), _p(mats8), len(mats8)))

    def __del__(self):
        try:
            lib().orc_free_scene(self.h)
        except Exception:
            pass

    @property
    def n_tris(self):
        return lib().orc_n_tris(self.h)

    @property
    def n_nodes(self):
        return lib().orc_n_nodes(self.h)

    def tris(self):
        n = self.n_tris
        v9 = np.zeros((n, 9), np.float32)
        mi = np.zeros(n, np.int32)
        lib().orc_get_tris(self.h, _p(v9), _p(mi))
        return v9, mi

    def mats(self):
        m = np.zeros((lib().orc_n_mats(self.h), 8), np.float32)
        lib().orc_get_mats(self.h, _p(m))
        return m

    def set_spheres(self, spheres):
        """EXTENSION (no reference counterpart): rows (cx, cy, cz, radius, material index)."""
        a = np.ascontiguousarray(np.asarray(spheres, np.float32).reshape(-1, 5))
        lib().orc_set_spheres(self.h, _p(a) if len(a) else None, len(a))

    def make_bih(self):
        lib().orc_make_bih(self.h)
        return dict(nodes=self.n_nodes, height=lib().orc_height(self.h), longest_leaf=lib().orc_longest_leaf(self.h),
                    leaves=lib().orc_n_leaves(self.h))

    def export_bih(self):
        root = np.zeros(6, np.float32)
        nodes = np.zeros((self.n_nodes, 4), np.uint32)
        leaf = np.zeros(self.n_tris, np.int32)
        lib().orc_export_bih(self.h, _p(root), _p(nodes), _p(leaf))
        return root, nodes, leaf

    def intersect_batch(self, org, dir, naive=False, counters=False, nthreads=None):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dir = np.ascontiguousarray(dir, np.float32).reshape(-1, 3)
        n = len(org)
        tri = np.zeros(n, np.int32)
        dist = np.zeros(n, np.float32)
        point = np.zeros((n, 3), np.float32)
        cn = np.zeros(5, np.uint64) if counters else None
        rc = lib().orc_intersect_batch(self.h, _p(org), _p(dir), n, int(naive), _p(tri), _p(dist), _p(point), _p(cn),
                                       nthreads or os.cpu_count())
        if rc:
            raise RuntimeError("orc_intersect_batch failed (BIH not built?)")
        return (tri, dist, point, cn) if counters else (tri, dist, point)

    def render(self, cam12, params, want_rgb8=True, nthreads=None):
        cam12 = np.ascontiguousarray(cam12, np.float32)
        n = params.rows * params.cols
        accum = np.zeros((params.rows, params.cols, 3), np.float32)
        rgb8 = np.zeros((params.rows, params.cols, 3), np.uint8) if want_rgb8 else None
        stats = np.zeros(2, np.uint64)
        cn = np.zeros(5, np.uint64)
        rc = lib().orc_render(self.h, _p(cam12), C.byref(params), _p(accum), _p(rgb8), _p(stats), _p(cn),
                              nthreads or os.cpu_count())
        if rc:
            raise RuntimeError("orc_render failed (BIH not built?)")
        return dict(accum=accum, rgb8=rgb8, rays=int(stats[0]), samples=int(stats[1]),
                    counters=dict(branch_visits=int(cn[0]), child_box_tests=int(cn[1]), own_box_tests=int(cn[2]),
                                  tri_tests=int(cn[3]), rays=int(cn[4])))


def render_window(scene, cam12, params, pix_lo, pix_hi, nthreads=None):
    """Radiance sums of pixels [pix_lo, pix_hi) of the frame described by params (row-major pixel index)."""
    cam12 = np.ascontiguousarray(cam12, np.float32)
    acc = np.zeros((pix_hi - pix_lo, 3), np.float32)
    rc = lib().orc_render_window(scene.h, _p(cam12), C.byref(params), pix_lo, pix_hi, _p(acc), nthreads or os.cpu_count())
    if rc:
        raise RuntimeError("orc_render_window failed")
    return acc


def load_camera(path):
    cam = np.zeros(12, np.float32)
    rc = lib().orc_load_camera(path.encode(), _p(cam))
    if rc:
        raise RuntimeError("Failed to parse camera file %s" % path)
    return cam


def make_rays(params, cam12):
    cam12 = np.ascontiguousarray(cam12, np.float32)
    n = params.rows * params.cols
    org = np.zeros((n, 3), np.float32)
    dir = np.zeros((n, 3), np.float32)
    lib().orc_make_rays(C.byref(params), _p(cam12), _p(org), _p(dir))
    return org, dir


def tone_map(accum, spp, trig=1):
    accum = np.ascontiguousarray(accum, np.float32)
    out = np.zeros(accum.shape, np.uint8)
    lib().orc_tone_map(_p(accum), accum.size // 3, spp, trig, _p(out))
    return out


def philox(ctr, key):
    ctr = np.asarray(ctr, np.uint32)
    key = np.asarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out
