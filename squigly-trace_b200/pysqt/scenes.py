"""Seeded synthetic scenes for BASELINE.json configs 3-5 (numpy only; harness, not hot path).

Each generator returns (v9, mat_idx, mats8): triangles as 9 floats in scene space, a material index per
triangle, and materials as (reflective, surf rgb, emissive, emit rgb) rows -- the same shape of data
trisFromObj (Obj.hs:49) yields.
"""
import numpy as np


def _quad_grid(p0, du, dv, nu, nv):
    """nu x nv cells over the parallelogram p0 + s*du + t*dv -> (2*nu*nv, 9) triangles."""
    s = np.linspace(0.0, 1.0, nu + 1, dtype=np.float64)
    t = np.linspace(0.0, 1.0, nv + 1, dtype=np.float64)
    P = (np.asarray(p0, np.float64)[None, None, :] + s[:, None, None] * np.asarray(du, np.float64)[None, None, :]
         + t[None, :, None] * np.asarray(dv, np.float64)[None, None, :])
    a, b, c, d = P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:]
    t1 = np.concatenate([a, b, c], -1).reshape(-1, 9)
    t2 = np.concatenate([a, c, d], -1).reshape(-1, 9)
    return np.concatenate([t1, t2]).astype(np.float32)


def _box(lo, hi, n):
    lo = np.asarray(lo, np.float64); hi = np.asarray(hi, np.float64); d = hi - lo
    ex, ey, ez = np.array([d[0], 0, 0]), np.array([0, d[1], 0]), np.array([0, 0, d[2]])
    faces = [(lo, ex, ey), (lo + ez, ex, ey), (lo, ex, ez), (lo + ey, ex, ez), (lo, ey, ez), (lo + ex, ey, ez)]
    return np.concatenate([_quad_grid(p, u, v, n, n) for p, u, v in faces])


def cornell_box(target_tris=10000, seed=0x5171):
    """Config 3: open-front box x,z in [-2,2], y in [-2,2] (camera looks down -y from y=7 like data/camera),
    one emissive quad under the ceiling, walls with reflective in {0, 0.2, 1}, two floor-standing boxes."""
    n = max(2, int(round(np.sqrt(target_tris / (2.0 * (5 + 12 * 0.25))))))
    m = max(1, n // 2)
    parts, mats_idx = [], []

    def add(tris, mat):
        parts.append(tris); mats_idx.append(np.full(len(tris), mat, np.int32))
    add(_quad_grid([-2, -2, -2], [4, 0, 0], [0, 4, 0], n, n), 0)      # floor (z=-2), diffuse grey
    add(_quad_grid([-2, -2, 2], [4, 0, 0], [0, 4, 0], n, n), 0)       # ceiling
    add(_quad_grid([-2, -2, -2], [4, 0, 0], [0, 0, 4], n, n), 1)      # back wall (y=-2), 0.2 reflective
    add(_quad_grid([-2, -2, -2], [0, 4, 0], [0, 0, 4], n, n), 2)      # left wall (x=-2), red 0.2
    add(_quad_grid([2, -2, -2], [0, 4, 0], [0, 0, 4], n, n), 3)       # right wall (x=2), mirror
    add(_box([-1.3, -1.2, -2.0], [-0.3, -0.2, 0.4], m), 4)            # tall box, diffuse
    add(_box([0.3, -0.3, -2.0], [1.3, 0.7, -0.9], m), 5)              # short box, mirror-ish
    add(_quad_grid([-0.5, -0.5, 1.98], [1, 0, 0], [0, 1, 0], 1, 1), 6)  # light
    mats = np.array([
        [0.0, 0.73, 0.73, 0.73, 0, 0, 0, 0],
        [0.2, 0.60, 0.60, 0.60, 0, 0, 0, 0],
        [0.2, 0.63, 0.065, 0.05, 0, 0, 0, 0],
        [1.0, 0.90, 0.90, 0.90, 0, 0, 0, 0],
        [0.0, 0.14, 0.45, 0.091, 0, 0, 0, 0],
        [1.0, 0.80, 0.80, 0.80, 0, 0, 0, 0],
        [0.0, 0.0, 0.0, 0.0, 100, 1, 1, 1],
    ], np.float32)
    return np.concatenate(parts), np.concatenate(mats_idx), mats


def subdivided_mesh(target_tris=1_000_000, seed=0x5172, half=None):
    """Config 4: a jittered height-field mesh (deep BIH, traversal bound) standing like a back wall behind the
    origin, facing the camera of data/camera (at y = 7 looking down -y); one emissive quad above it, diffuse.
    The extent grows with the triangle count so that every triangle keeps |e1 x e2| well above the reference's
    ABSOLUTE epsilon 1e-4 (Geometry.hs:142) -- smaller triangles are simply invisible to mollerTrumbore."""
    rng = np.random.default_rng(seed)
    n = max(2, int(round(np.sqrt(target_tris / 2.0))))
    if half is None:
        half = max(2.0, 0.03 * n)                      # edge = 2*half/n >= 0.06
    xs = np.linspace(-half, half, n + 1); zs = np.linspace(-half, half, n + 1)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    edge = 2.0 * half / n
    Y = (-0.5 * half - 1.0 + 0.35 * np.sin(2.3 * X) * np.cos(1.7 * Z) + 0.15 * np.sin(9.1 * X + 1.0) * np.sin(7.7 * Z)
         + rng.uniform(-1e-3, 1e-3, X.shape) * edge)
    X = X + rng.uniform(-1e-3, 1e-3, X.shape) * edge
    Z = Z + rng.uniform(-1e-3, 1e-3, X.shape) * edge
    P = np.stack([X, Y, Z], -1)
    a, b, c, d = P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:]
    tris = np.concatenate([np.concatenate([a, b, c], -1).reshape(-1, 9), np.concatenate([a, c, d], -1).reshape(-1, 9)])
    q = 0.35 * half
    light = _quad_grid([-q, -0.5 * half + 1.5, 0.45 * half], [2 * q, 0, 0], [0, 2 * q, 0], 1, 1)
    v9 = np.concatenate([tris.astype(np.float32), light])
    mi = np.concatenate([np.zeros(len(tris), np.int32), np.ones(len(light), np.int32)])
    mats = np.array([[0.0, 0.7, 0.7, 0.7, 0, 0, 0, 0], [0.0, 0, 0, 0, 100, 1, 1, 1]], np.float32)
    return v9, mi, mats


def triangle_soup(n_tris=10_000_000, seed=0x5173, all_reflective=True, scale=1.0):
    """Config 5: centroids ~U([-1,1]^3), edge length ~U(0.002,0.02), random orientation, materials reflective=1, one
    in 64 triangles emissive -- all lengths times `scale`.  At scale=1 (BASELINE.json's literal numbers) |e1 x e2| of
    most triangles is below the reference's ABSOLUTE epsilon 1e-4 (Geometry.hs:118,142), so mollerTrumbore rejects
    them at the `a` guard and paths die at the primary ray: not a stress test.  The bench therefore runs the soup at
    scale=40 (the same factor as the 1M-triangle mesh of config 4): edges 0.08..0.8 in a [-40,40]^3 cube, mean free
    path about 1.3 units, so mirror paths really bounce 16 times through 480 MB of triangles."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, (n_tris, 3))
    L = rng.uniform(0.002, 0.02, (n_tris, 1))
    e = rng.normal(size=(n_tris, 3, 3))
    e /= np.linalg.norm(e, axis=-1, keepdims=True)
    v = (c[:, None, :] + e * L[:, None, :] * 0.5) * float(scale)
    v9 = v.reshape(n_tris, 9).astype(np.float32)
    mi = (rng.integers(0, 64, n_tris) == 0).astype(np.int32)
    r = 1.0 if all_reflective else 0.3
    mats = np.array([[r, 0.9, 0.9, 0.9, 0, 0, 0, 0], [r, 0.9, 0.9, 0.9, 20, 1, 0.9, 0.8]], np.float32)
    return v9, mi, mats


# BASELINE.json `configs`, as the bench and the tests run them (index = position in that list)
CONFIGS = {
    0: dict(name="config1: data/scene.obj 540x540 100spp depth3 (reference defaults, literal index convention)",
            scene="obj", width=540, height=540, spp=100, depth=3, literal=True),
    1: dict(name="config2: data/scene.obj 1920x1080 1024spp depth8", scene="obj", width=1920, height=1080, spp=1024, depth=8, literal=False),
    2: dict(name="config3: synthetic Cornell box ~10k tris 1920x1080 4096spp depth8", scene="cornell", n_tris=10000,
            width=1920, height=1080, spp=4096, depth=8, literal=False),
    3: dict(name="config4: synthetic 1M-triangle mesh 3840x2160 256spp depth8", scene="mesh", n_tris=1_000_000,
            width=3840, height=2160, spp=256, depth=8, literal=False),
    4: dict(name="config5: 10M-triangle soup (x40 scale), all reflective, 16 bounces, 3840x2160 64spp", scene="soup",
            n_tris=10_000_000, width=3840, height=2160, spp=64, depth=16, literal=False),
}


def config_arrays(cfg):
    """(v9, mat_idx, mats8) of a synthetic config; None for the .obj scene."""
    if cfg["scene"] == "cornell":
        return cornell_box(cfg["n_tris"])
    if cfg["scene"] == "mesh":
        return subdivided_mesh(cfg["n_tris"])
    if cfg["scene"] == "soup":
        return triangle_soup(cfg["n_tris"], scale=40.0)
    return None
