"""Fuzz campaign for the exact culling (leaf culling + child slabs): random scenes of five kinds (clusters with
anisotropic extent, thin bumpy sheets, needles that cross the whole box, coordinates around 1000, a 0.02-sized scene) x rays aimed at
the geometry, the tame-boundary set and the adversarial set; the host build of the device code (tests/emu, culling on, plain and
interleaved-pool layout) must return the oracle's hits bit for bit.
usage: python tests/fuzz_slabs.py SEED SECONDS     (round 2: seeds 1-4 x 600 s and 21-26 x 1200 s = 18 122 scenes, 110 M rays, no mismatch)
tests/test_emu_parity.py::test_culling_fuzz_short runs a dozen scenes of it."""
import os
import sys, time
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(_R, 'tests'), os.path.join(_R, 'squigly-trace_b200'), _R):
    if _p not in sys.path: sys.path.insert(0, _p)
import numpy as np, pysqt
from pysqt import scenes
from common import build_pair, Emu, random_rays, adversarial_rays, tame_boundary_rays, assert_same_hits
rng = None
def rand_scene(kind, n):
    if kind == 0:   # clustered small triangles, anisotropic extent
        c = rng.uniform(-1, 1, (n, 1, 3)) * rng.choice([0.01, 1.0, 30.0], 3)
        v = c + rng.normal(size=(n, 3, 3)) * rng.choice([1e-3, 0.05, 0.5])
    elif kind == 1: # thin sheet with bumps (one axis never split)
        u = rng.uniform(-2, 2, (n, 1, 2)); d = rng.normal(size=(n, 3, 2)) * 0.05
        xy = u + d; z = 0.02 * np.sin(5 * xy[..., :1]) + rng.normal(size=(n, 3, 1)) * 1e-3
        v = np.concatenate([xy, z], -1)
        v = v[..., rng.permutation(3)]
    elif kind == 2: # long needles crossing the whole box
        a = rng.uniform(-1, 1, (n, 1, 3)); b = a + rng.normal(size=(n, 1, 3)) * 2.0; c2 = a + rng.normal(size=(n, 1, 3)) * 0.02
        v = np.concatenate([a, b, c2], 1)
    elif kind == 3: # far from the origin, large coordinates
        v = 1000.0 + rng.uniform(-5, 5, (n, 1, 3)) + rng.normal(size=(n, 3, 3)) * 0.3
    else:           # tiny scene
        v = (rng.uniform(-1, 1, (n, 1, 3)) + rng.normal(size=(n, 3, 3)) * 0.1) * 0.02
    v9 = v.reshape(n, 9).astype(np.float32)
    mats = np.array([[0.3, .5, .5, .5, 0, 0, 0, 0]], np.float32)
    return v9, np.zeros(n, np.int32), mats
def campaign(seed, seconds, max_scenes=None, quiet=False):
    global rng
    rng = np.random.default_rng(seed)
    t0 = time.time(); cases = 0
    while (time.time() - t0 < seconds) and (max_scenes is None or cases < max_scenes):
        kind = int(rng.integers(0, 5)); n = int(rng.choice([50, 400, 3000, 12000]))
        v9, mi, mats = rand_scene(kind, n)
        osc, hs = build_pair(v9, mi, mats)
        e = Emu(hs)
        pts = v9.reshape(-1, 3); lo, hi = pts.min(0), pts.max(0); c = 0.5 * (lo + hi); h = 0.5 * (hi - lo) + 1e-6
        m = 4000
        org = (c + rng.uniform(-2.2, 2.2, (m, 3)) * h).astype(np.float32)
        tgt = pts[rng.integers(0, len(pts), m)] + rng.normal(size=(m, 3)) * h * 0.01
        d = (tgt - org); d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-30
        d = (d * rng.choice([1.0, 1.0, 0.5, 1.15], (m, 1))).astype(np.float32)
        o2, d2 = tame_boundary_rays(v9, 1000, seed=int(rng.integers(1 << 30)))
        o3, d3 = adversarial_rays(v9, seed=int(rng.integers(1 << 30)), n_each=64)
        O_ = np.concatenate([org, o2, o3]); D_ = np.concatenate([d, d2, d3])
        want = osc.intersect_batch(O_, D_)
        got = e.intersect_batch(O_, D_)
        assert_same_hits(got, want, "kind %d n %d" % (kind, n))
        ti, di, bad = e.intersect_batch_interleaved(O_[:2048], D_[:2048])
        assert bad == 0 and np.array_equal(ti, want[0][:2048]) and np.array_equal(di.view(np.uint32), want[1][:2048].view(np.uint32))
        cases += 1
        if cases % 10 == 0 and not quiet: print(cases, "scenes ok, hits", int((want[0] >= 0).sum()), "culled", got[3]["leaves_culled"], flush=True)
    return cases


if __name__ == '__main__':
    n = campaign(int(sys.argv[1]) if len(sys.argv) > 1 else 1, float(sys.argv[2]) if len(sys.argv) > 2 else 300.0)
    print('DONE', n, 'scenes, no mismatch')
