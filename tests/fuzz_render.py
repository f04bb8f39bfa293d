"""Fuzz campaign for the integrator on top of the exact culling: random scenes (three of the kinds of tests/fuzz_slabs.py, scaled
into the default camera's view) with random materials (diffuse / half / fully reflective, some emissive), depth 1 / 3 / 8, small frames;
the host build of the device code (tests/emu: same unit steps, integrator and accumulation order as the kernels) must produce the
oracle's accumulation buffer and RGB8 bit for bit.
usage: python tests/fuzz_render.py SEED SECONDS     (round 2: seeds 11-13 x 300 s = 22 795 frames, 17 172 of them lit, no mismatch)"""
import os
import sys
import time
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(_R, 'tests'), os.path.join(_R, 'squigly-trace_b200'), _R):
    if _p not in sys.path: sys.path.insert(0, _p)
import numpy as np, pysqt
import fuzz_slabs
from common import build_pair, Emu, bits
from oracle import oracle as O
seed = int(sys.argv[1]); secs = float(sys.argv[2])
fuzz_slabs.rng = np.random.default_rng(seed)
rng = fuzz_slabs.rng
cam = pysqt.load_camera(pysqt.ROOT + "/data/camera")
t0 = time.time(); cases = 0; lit = 0
while time.time() - t0 < secs:
    kind = int(rng.integers(0, 3)); n = int(rng.choice([200, 2000, 8000]))
    v9, mi, mats = fuzz_slabs.rand_scene(kind, n)
    v9 = (v9.reshape(-1, 3) * np.float32(1.5 / max(1e-6, np.abs(v9).max()))).reshape(n, 9)      # fit the camera's view
    nm = 6
    mats = np.zeros((nm, 8), np.float32)
    mats[:, 0] = rng.choice([0.0, 0.3, 1.0], nm); mats[:, 1:4] = rng.uniform(0, 1, (nm, 3))
    mats[:, 4] = rng.choice([0.0, 0.0, 2.0, 10.0], nm); mats[:, 5:8] = rng.uniform(0, 1, (nm, 3))
    mats[0, 4] = 5.0
    mi = rng.integers(0, nm, n).astype(np.int32)
    osc, hs = build_pair(v9, mi, mats)
    e = Emu(hs)
    depth = int(rng.choice([1, 3, 8])); spp = int(rng.choice([2, 5]))
    p = pysqt.make_params(48, 36, spp, max_depth=depth, seed=int(rng.integers(1000)))
    op = O.make_params(48, 36, spp, max_depth=depth, seed=int(p.seed), trig=1)
    out = e.render(cam, p); ref = osc.render(cam, op)
    assert np.array_equal(bits(out["accum"]), bits(ref["accum"])) and np.array_equal(out["rgb8"], ref["rgb8"]), (kind, n, depth, spp)
    lit += int((ref["accum"] != 0).any())
    cases += 1
print("DONE", cases, "renders bit-exact,", lit, "with light in the frame")
