"""Kernel logic (csrc/sqt_core.cuh + sqt_paths.cuh compiled for the host, tests/emu) against the oracle.

This is the same source the CUDA kernels instantiate, run lane by lane on the CPU, so traversal/integrator logic
bugs are caught without a GPU.  The GPU parity tests proper are in test_gpu_parity.py.  CPU only."""
import numpy as np
import pytest

import pysqt
from oracle import oracle as O
from pysqt import scenes
from common import Emu, adversarial_rays, assert_same_hits, bits, build_pair, random_rays


@pytest.fixture(scope="module")
def emu(host_scene):
    return Emu(host_scene)


def test_primary_and_random_rays(emu, oracle_scene, camera):
    org, dirs = O.make_rays(O.make_params(160, 160, 1), camera)
    o2, d2 = random_rays(30000, 1)
    org = np.concatenate([org, o2]); dirs = np.concatenate([dirs, d2])
    want = oracle_scene.intersect_batch(org, dirs, counters=True)
    emu.set_leaf_cull(False)
    got = emu.intersect_batch(org, dirs)
    emu.set_leaf_cull(True)
    assert_same_hits(got, want, "emu, culling off")
    assert got[3]["branch_visits"] == int(want[3][0]) and got[3]["tri_tests"] == int(want[3][3]) and got[3]["child_box_tests"] == int(want[3][1])
    culled = emu.intersect_batch(org, dirs)
    assert_same_hits(culled, want, "emu, culling on")
    assert culled[3]["leaves_culled"] > 0 and culled[3]["tri_tests"] < got[3]["tri_tests"]


def test_adversarial_rays(emu, oracle_scene):
    v9, _ = oracle_scene.tris()
    org, dirs = adversarial_rays(v9, n_each=512)
    assert_same_hits(emu.intersect_batch(org, dirs), oracle_scene.intersect_batch(org, dirs), "adversarial")


def test_interleaved_stack_layout(emu, oracle_scene, camera):
    """The pool kernel keeps the stacks of a warp's 64 rays in one region, interleaved in 32-byte granules, with the
    stride launch_pool derives from the tree height.  Host run of exactly that layout, 64 live rays stepped round-robin:
    same hits as the oracle and nothing written behind the region."""
    org, dirs = random_rays(6000, seed=11)
    o2, d2 = O.make_rays(O.make_params(64, 48, 1), camera)
    org = np.concatenate([org, o2]); dirs = np.concatenate([dirs, d2])
    tri, dist, bad = emu.intersect_batch_interleaved(org, dirs)
    ref = oracle_scene.intersect_batch(org, dirs)
    assert bad == 0
    assert np.array_equal(tri, ref[0]) and np.array_equal(dist.view(np.uint32), ref[1].view(np.uint32))
    assert (tri >= 0).mean() > 0.3
    # a deeper tree (triangle soup: height 15+, many both-children-hit branches -> deep stacks)
    v9, mi, mats = scenes.triangle_soup(120000, seed=3)
    osc, hs = build_pair(v9 * 40.0, mi, mats)
    deep = Emu(hs)
    org, dirs = random_rays(3000, seed=12, lo=-20, hi=20)
    tri, dist, bad = deep.intersect_batch_interleaved(org, dirs)
    ref = osc.intersect_batch(org, dirs)
    assert bad == 0 and hs.stats()["height"] >= 15
    assert np.array_equal(tri, ref[0]) and np.array_equal(dist.view(np.uint32), ref[1].view(np.uint32))


@pytest.mark.parametrize("gen,n", [("cornell", 4000), ("soup", 30000), ("mesh", 20000)])
def test_synthetic_scenes(gen, n):
    v9, mi, mats = {"cornell": scenes.cornell_box, "soup": scenes.triangle_soup, "mesh": scenes.subdivided_mesh}[gen](n)
    osc, hs = build_pair(v9, mi, mats)
    e = Emu(hs)
    org, dirs = random_rays(20000, 2, lo=-1.5, hi=1.5)
    assert_same_hits(e.intersect_batch(org, dirs), osc.intersect_batch(org, dirs), gen)


@pytest.mark.parametrize("w,h,spp,depth,literal,flags", [
    (64, 48, 6, 3, False, 0), (48, 48, 4, 8, False, 0), (56, 40, 3, 3, True, 0),
    (48, 32, 4, 5, False, pysqt.SQT_F_NO_PRIMARY_REUSE | pysqt.SQT_F_NO_EARLY_TERMINATION), (32, 32, 2, 1, False, 0)])
def test_render_bit_exact(emu, oracle_scene, camera, w, h, spp, depth, literal, flags):
    got = emu.render(camera, pysqt.make_params(w, h, spp, max_depth=depth, seed=5, literal=literal, flags=flags))
    ref = oracle_scene.render(camera, O.make_params(w, h, spp, max_depth=depth, seed=5, trig=1, literal=literal))
    assert np.array_equal(bits(got["accum"]), bits(ref["accum"])) and np.array_equal(got["rgb8"], ref["rgb8"])
    assert got["samples"] == ref["samples"]
    if flags:
        assert got["rays"] == ref["rays"]
    else:
        assert got["rays"] <= ref["rays"]


def test_cast_mode(emu, oracle_scene, camera):
    got = emu.render(camera, pysqt.make_params(64, 48, 3, mode=1))
    ref = oracle_scene.render(camera, O.make_params(64, 48, 3, mode=1, trig=1))
    assert np.array_equal(bits(got["accum"]), bits(ref["accum"])) and np.array_equal(got["rgb8"], ref["rgb8"])


def test_rank_partition_sums_to_full_frame(emu, camera):
    p = pysqt.make_params(72, 40, 3, max_depth=4, seed=2)
    full = emu.render(camera, p)["accum"]
    parts = [emu.render(camera, p, rank=r, world=4)["accum"] for r in range(4)]
    assert np.array_equal(bits(sum(parts)), bits(full))
    ps = pysqt.make_params(72, 40, 4, max_depth=4, seed=2, flags=pysqt.SQT_F_SPLIT_SAMPLES)
    halves = [emu.render(camera, ps, rank=r, world=2)["accum"] for r in range(2)]
    assert np.allclose(halves[0] + halves[1], emu.render(camera, pysqt.make_params(72, 40, 4, max_depth=4, seed=2))["accum"], rtol=1e-6, atol=1e-6)


def test_upload_validation_errors(host_scene):
    import ctypes as C
    from common import emu_lib
    E = emu_lib()

    def err_of(desc):
        h = E.emu_upload(C.byref(desc)); e = E.emu_error(h).decode(); E.emu_free(h); return e
    d = host_scene.desc(); d.n_mats = 0
    assert "no materials" in err_of(d)
    nodes = host_scene.nodes.copy(); nodes["a"][0] = (nodes["a"][0] & 0xC0000000) | 0x3fffffff
    d = host_scene.desc(); d.nodes = nodes.ctypes.data
    assert "out of range" in err_of(d)
    nodes = host_scene.nodes.copy(); nodes["b"][0] = 0          # right child = root: not a tree
    d = host_scene.desc(); d.nodes = nodes.ctypes.data
    assert "twice" in err_of(d)
    tris = host_scene.tris.copy(); tris["material"][5] = 99
    d = host_scene.desc(); d.tris = tris.ctypes.data
    assert "material 99" in err_of(d)


SPHERES = [(0.8, 0.5, -1.2, 0.6, 5), (-0.9, 1.0, 0.3, 0.45, 1), (0.0, -0.5, 1.2, 0.3, 3), (0.0, 3.0, 0.5, 0.25, 2)]


def test_sphere_extension(host_scene, camera):
    """Analytic spheres (north-star extension without a reference counterpart): device logic vs the oracle's
    restatement of the same semantics -- mirror, diffuse and emissive spheres, one of them outside the BIH bounds."""
    e = Emu(host_scene)
    osc = O.Scene.load(pysqt.ROOT + "/data/scene.obj", pysqt.ROOT + "/data")
    osc.make_bih()
    e.set_spheres(SPHERES); osc.set_spheres(SPHERES)
    org, dirs = O.make_rays(O.make_params(160, 160, 1), camera)
    o2, d2 = random_rays(40000, 9)
    org = np.concatenate([org, o2]); dirs = np.concatenate([dirs, d2])
    want = osc.intersect_batch(org, dirs)
    assert_same_hits(e.intersect_batch(org, dirs), want, "spheres")
    assert (want[0] >= osc.n_tris).sum() > 1000
    naive = osc.intersect_batch(org[:4000], dirs[:4000], naive=True)
    assert np.array_equal(bits(naive[1]), bits(want[1][:4000]))
    for depth, mode in ((3, 0), (8, 0), (3, 1)):
        got = e.render(camera, pysqt.make_params(72, 48, 5, max_depth=depth, seed=3, mode=mode))
        ref = osc.render(camera, O.make_params(72, 48, 5, max_depth=depth, seed=3, trig=1, mode=mode))
        assert np.array_equal(bits(got["accum"]), bits(ref["accum"])) and np.array_equal(got["rgb8"], ref["rgb8"])
    e.set_spheres([])
    assert_same_hits(e.intersect_batch(org[:5000], dirs[:5000]), O.Scene.intersect_batch(_plain(osc), org[:5000], dirs[:5000]), "spheres removed")


def _plain(osc):
    osc.set_spheres([])
    return osc


def _big_overlapping_triangles(n, seed):
    """Large triangles whose extents overlap heavily: the +-0.001 planes of many nodes stick out of the clipped box they
    split (lmax > hi[ax] or rmin < lo[ax]), which is the case the interval stepping of desc_step must hand to the literal
    six-slab path (nodes flagged kSlow at upload)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, (n, 1, 3))
    e = rng.normal(size=(n, 3, 3)) * rng.uniform(0.05, 1.5, (n, 1, 1))
    v9 = (c + e).reshape(n, 9).astype(np.float32)
    # a few axis-aligned slivers that end exactly on the bounds of the whole set
    lo, hi = v9.reshape(-1, 3).min(0), v9.reshape(-1, 3).max(0)
    extra = np.array([[lo[0], lo[1], lo[2], hi[0], lo[1], lo[2], hi[0], hi[1], lo[2]],
                      [lo[0], hi[1], hi[2], hi[0], hi[1], hi[2], lo[0], lo[1], hi[2]]], np.float32)
    # long triangles with one vertex exactly ON a face of the bounds and their centroid far inside: wherever such a triangle
    # falls on the low side of a split along that axis, lmax = 0.001 + its vertex exceeds the (unpadded) root extent
    k = n // 8
    far = rng.uniform(lo, hi, (k, 3, 3)).astype(np.float32) * np.float32(0.3)
    face = rng.integers(0, 6, k)
    for i in range(k):
        ax = face[i] % 3
        far[i, 0, ax] = hi[ax] if face[i] < 3 else lo[ax]
    v9 = np.concatenate([v9, extra, far.reshape(k, 9)])
    mats = np.array([[0.3, .5, .5, .5, 0, 0, 0, 0]], np.float32)
    return v9, np.zeros(len(v9), np.int32), mats


def test_nodes_whose_planes_leave_their_box_take_the_literal_path():
    v9, mi, mats = _big_overlapping_triangles(3000, 17)
    osc, hs = build_pair(v9, mi, mats)
    e = Emu(hs)
    import ctypes as C
    from common import emu_lib
    emu_lib().emu_slow_nodes.restype = C.c_int
    emu_lib().emu_slow_nodes.argtypes = [C.c_void_p]
    assert emu_lib().emu_slow_nodes(e.h) > 20, "the scene is meant to exercise the slow path"
    org, dirs = random_rays(20000, 23, lo=-3, hi=3)
    o2, d2 = adversarial_rays(v9, seed=3, n_each=64)
    org = np.concatenate([org, o2]); dirs = np.concatenate([dirs, d2])
    want = osc.intersect_batch(org, dirs, counters=True)
    e.set_leaf_cull(False)
    got = e.intersect_batch(org, dirs)
    assert_same_hits(got, want, "slow nodes")
    assert got[3]["branch_visits"] == int(want[3][0]) and got[3]["tri_tests"] == int(want[3][3])
    e.set_leaf_cull(True)
    assert_same_hits(e.intersect_batch(org, dirs), want, "slow nodes, culling on")
    tri, dist, bad = e.intersect_batch_interleaved(org, dirs)
    assert bad == 0 and np.array_equal(tri, want[0]) and np.array_equal(dist.view(np.uint32), want[1].view(np.uint32))


def test_non_finite_planes_disable_interval_stepping():
    """A vertex at infinity makes box planes infinite (and slab values NaN): every ray must then take the literal path with
    the Haskell min/max (operand order matters with NaN), like the oracle."""
    v9, mi, mats = scenes.cornell_box(600)
    v9 = v9.copy()
    v9[5, 0] = np.inf
    v9[17, 4] = -np.inf
    osc, hs = build_pair(v9, mi, mats)
    e = Emu(hs)
    org, dirs = random_rays(8000, 29, lo=-2.5, hi=2.5)
    assert_same_hits(e.intersect_batch(org, dirs), osc.intersect_batch(org, dirs), "infinite planes")


def many_spheres(n, seed, duplicates=True, log_r=(-2.5, -0.6)):
    """n spheres inside the reference scene's bounds; every 16th one is an exact copy of an earlier one (same centre and
    radius -> identical dist: the earlier index must win), radii log-uniform from tiny to large."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1.9, 1.9, (n, 3)).astype(np.float32)
    r = (10.0 ** rng.uniform(log_r[0], log_r[1], n)).astype(np.float32)
    m = rng.integers(0, 6, n)
    if duplicates:
        for k in range(16, n, 16):
            j = int(rng.integers(0, k))
            c[k] = c[j]; r[k] = r[j]
    return [(float(c[k, 0]), float(c[k, 1]), float(c[k, 2]), float(r[k]), int(m[k])) for k in range(n)]


def test_sphere_hierarchy_gives_the_definitions_result(host_scene, camera):
    """The spheres are found through a bounding-volume hierarchy (sphere_step); the result must be the definition's --
    closest of [BIH hit, sphere 0, sphere 1, ..], earliest on ties -- for camera rays, incoherent rays, rays from sphere
    surfaces, duplicated spheres, and the adversarial rays (zero / tiny / huge components take the literal loop)."""
    e = Emu(host_scene)
    osc = O.Scene.load(pysqt.ROOT + "/data/scene.obj", pysqt.ROOT + "/data")
    osc.make_bih()
    sph = many_spheres(3000, 5)
    e.set_spheres(sph); osc.set_spheres(sph)
    org, dirs = O.make_rays(O.make_params(120, 90, 1), camera)
    o2, d2 = random_rays(30000, 19)
    v9, _ = osc.tris()
    o3, d3 = adversarial_rays(v9, seed=4, n_each=64)
    c = np.array([s[:3] for s in sph[:2000]], np.float32); rr = np.array([s[3] for s in sph[:2000]], np.float32)
    rng = np.random.default_rng(2)
    u = rng.normal(size=(2000, 3)).astype(np.float32); u /= np.linalg.norm(u, axis=1, keepdims=True)
    o4 = (c + u * rr[:, None]).astype(np.float32); d4 = rng.normal(size=(2000, 3)).astype(np.float32)      # origins ON sphere surfaces
    org = np.concatenate([org, o2, o3, o4]); dirs = np.concatenate([dirs, d2, d3, d4])
    want = osc.intersect_batch(org, dirs)
    got = e.intersect_batch(org, dirs)
    assert_same_hits(got, want, "sphere hierarchy")
    assert (want[0] >= osc.n_tris).sum() > 8000
    import ctypes as C
    from common import emu_lib
    emu_lib().emu_set_sphere_bvh.argtypes = [C.c_void_p, C.c_int]
    emu_lib().emu_set_sphere_bvh(e.h, 0)                    # the literal loop over all spheres, same answer
    assert_same_hits(e.intersect_batch(org[:6000], dirs[:6000]), (want[0][:6000], want[1][:6000], want[2][:6000]), "sphere loop")
    emu_lib().emu_set_sphere_bvh(e.h, 1)
    e.set_spheres(sph)
    got = e.render(camera, pysqt.make_params(64, 48, 4, max_depth=6, seed=3))
    ref = osc.render(camera, O.make_params(64, 48, 4, max_depth=6, seed=3, trig=1))
    assert np.array_equal(bits(got["accum"]), bits(ref["accum"])) and np.array_equal(got["rgb8"], ref["rgb8"])


def _filter_stats(tri12, org, dirs):
    import ctypes as C
    from common import emu_lib, _p
    tri12 = np.ascontiguousarray(tri12, np.float32); org = np.ascontiguousarray(org, np.float32); dirs = np.ascontiguousarray(dirs, np.float32)
    out = np.zeros(5, np.uint64)
    E = emu_lib()
    E.emu_filter_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
    E.emu_filter_stats(_p(tri12), _p(org), _p(dirs), len(org), _p(out))
    return [int(x) for x in out]


def test_triangle_filter_is_conservative():
    """The pool kernel's division-free a/u filter (moller_trumbore_au) never rejects a pair the full test takes past `u`,
    agrees with it on the `a` guard, and lets through hardly anything else -- on ordinary pairs, on pairs whose u sits
    on 0 or 1 to the last bit, on extreme scales and on non-finite input."""
    rng = np.random.default_rng(77)
    n = 400_000
    v0 = rng.uniform(-1, 1, (n, 3)); e1 = rng.uniform(-1, 1, (n, 3)) * 0.5; e2 = rng.uniform(-1, 1, (n, 3)) * 0.5
    # aim at a point u*e1 + v*e2 with u spread around [0, 1], a third of them exactly on u = 0 or u = 1
    u = rng.uniform(-0.5, 1.5, n); v = rng.uniform(-0.2, 1.0, n)
    kind = rng.integers(0, 3, n)
    u = np.where(kind == 1, 0.0, np.where(kind == 2, 1.0, u))
    target = v0 + u[:, None] * e1 + v[:, None] * e2
    org = target + rng.normal(size=(n, 3)) * rng.choice([0.01, 1.0, 30.0], (n, 1))
    d = target - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d *= rng.choice([1.0, 1.0, 1e-3, 1e3], (n, 1))
    tri = np.zeros((n, 12), np.float32)
    tri[:, 0:3] = v0; tri[:, 3:6] = e1; tri[:, 6:9] = e2
    pairs, exact, filt, viol, amis = _filter_stats(tri, org, d)
    assert pairs == n and viol == 0 and amis == 0
    assert exact > n // 4
    o = kind == 0                                       # the ordinary pairs: the filter is as selective as the guard itself
    _, ex_o, f_o, _, _ = _filter_stats(tri[o], org[o], d[o])
    assert ex_o <= f_o <= ex_o + max(20, ex_o // 5000), (ex_o, f_o)
    # extreme scales: everything multiplied by 2^k (the geometry is scale invariant up to the eps guards)
    for k in (-60, -30, -12, 12, 30, 60):
        sc = np.float32(2.0) ** k
        p2, ex2, f2, viol2, amis2 = _filter_stats(tri * sc, (org * sc).astype(np.float32), d.astype(np.float32))
        assert viol2 == 0 and amis2 == 0, (k, viol2, amis2)
    # non-finite and zero input anywhere
    m = 60_000
    sub = slice(0, m)
    bad = np.array([np.inf, -np.inf, np.nan, 0.0, -0.0, 1e-45, -1e-45, 3e38], np.float32)
    tri_b = tri[sub].copy(); org_b = org[sub].astype(np.float32).copy(); d_b = d[sub].astype(np.float32).copy()
    where = rng.integers(0, 15, m)
    val = bad[rng.integers(0, len(bad), m)]
    for i in range(m):
        w = where[i]
        if w < 9: tri_b[i, w] = val[i]
        elif w < 12: org_b[i, w - 9] = val[i]
        else: d_b[i, w - 12] = val[i]
    p3, ex3, f3, viol3, amis3 = _filter_stats(tri_b, org_b, d_b)
    assert viol3 == 0 and amis3 == 0, (viol3, amis3)


def test_child_slabs_are_exact_around_the_tame_boundary():
    """Child slabs on a scene where they matter (a height-field mesh: every subtree keeps the root's extent on the
    thin axis): rays whose |d|_1 straddles 2 and whose origin straddles twice the root's half extent -- the two limits of `tame`,
    where a ray switches between using and ignoring the slabs --, rays that graze the surface, and rays with direction lengths from
    1e-3 to 1e3.  Results equal the oracle's bit for bit; with the culling off they do as well, and the culling really removes visits."""
    import ctypes as C
    from common import emu_lib
    v9, mi, mats = scenes.subdivided_mesh(20000)
    osc, hs = build_pair(v9, mi, mats)
    e = Emu(hs)
    emu_lib().emu_tight_children.restype = C.c_int
    emu_lib().emu_tight_children.argtypes = [C.c_void_p]
    assert emu_lib().emu_tight_children(e.h) > 1000, "the mesh is meant to have many usable slabs"
    from common import tame_boundary_rays
    n = 6000
    org, dirs = tame_boundary_rays(v9, n)
    want = osc.intersect_batch(org, dirs)
    got = e.intersect_batch(org, dirs)
    assert_same_hits(got, want, "child slabs on")
    assert (want[0] >= 0).sum() > n // 2
    e.set_leaf_cull(False)
    got_off = e.intersect_batch(org, dirs)
    assert_same_hits(got_off, want, "culling off")
    assert got_off[3]["leaves_culled"] == 0
    assert got[3]["branch_visits"] < 0.75 * got_off[3]["branch_visits"], (got[3]["branch_visits"], got_off[3]["branch_visits"])
    assert got[3]["tri_tests"] < 0.5 * got_off[3]["tri_tests"]


def test_slab_bound_packing_is_conservative():
    """pack_slab_lo hides the slab's axis code in the two low mantissa bits of its lower bound: the packed value must never be
    above the input (a slab may only grow), must stay within a few ulps of it, and must give the code back."""
    import ctypes as C
    from common import emu_lib, _p
    rng = np.random.default_rng(5)
    n = 200_000
    bits_ = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    lo = bits_.view(np.float32).copy()
    lo[:8] = [0.0, -0.0, 1e-45, -1e-45, 1e-30, -1e-30, 3.4e38, -3.4e38]
    lo[8:5000] = rng.uniform(-50, 50, 4992).astype(np.float32)
    finite = np.isfinite(lo) & (np.abs(lo) < 3.0e38)
    lo = lo[finite]
    code = rng.integers(0, 4, len(lo)).astype(np.uint32)
    out = np.zeros(len(lo), np.float32)
    E = emu_lib()
    E.emu_pack_slab_lo.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
    E.emu_pack_slab_lo(_p(lo), _p(code), len(lo), _p(out))
    assert np.all(out <= lo)
    assert np.array_equal(out.view(np.uint32) & 3, code)
    big = np.abs(lo) >= 1e-30
    ulp = np.spacing(np.abs(lo[big]))
    assert np.all(lo[big] - out[big] <= 8 * ulp)
    assert np.all(out[~big] >= -1.0001e-30)              # zeros and denormals become -1e-30


def test_culling_fuzz_short():
    """A dozen scenes of tests/fuzz_slabs.py (the full campaign of round 2: 18 122 scenes, 110 M rays, no mismatch)."""
    import fuzz_slabs
    assert fuzz_slabs.campaign(seed=20261018, seconds=120.0, max_scenes=12, quiet=True) == 12
