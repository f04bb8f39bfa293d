"""Measures the BASELINE.json configs that are not the bench line (1, 3, 4, 5) at a reduced spp, with a closest-hit
parity spot check against the oracle on each scene.  usage: python tools/run_configs.py [1 3 4 5]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "squigly-trace_b200"), os.path.join(ROOT, "tests")]
import pysqt
from pysqt import scenes
from oracle import oracle as O

DATA = os.path.join(ROOT, "data")
which = [int(x) for x in sys.argv[1:]] or [1, 3, 4, 5]
cam = pysqt.load_camera(os.path.join(DATA, "camera"))
ctx = pysqt.Context(0)
out = []
for c in which:
    t0 = time.time()
    if c == 1:
        hs = pysqt.HostScene.load(os.path.join(DATA, "scene.obj"), DATA); osc = O.Scene.load(os.path.join(DATA, "scene.obj"), DATA)
        name, W, H, spp, full_spp, depth, literal = "config 1: data/scene.obj 540x540 100spp depth3 (reference defaults)", 540, 540, 100, 100, 3, True
    else:
        gen, n, name, W, H, spp, full_spp, depth = {
            3: (scenes.cornell_box, 10000, "config 3: Cornell ~10k tris 1920x1080", 1920, 1080, 128, 4096, 8),
            4: (scenes.subdivided_mesh, 1_000_000, "config 4: 1M-tri mesh 3840x2160", 3840, 2160, 16, 256, 8),
            5: (scenes.triangle_soup, 10_000_000, "config 5: 10M-tri soup, all reflective, 3840x2160", 3840, 2160, 4, 64, 16)}[c]
        v9, mi, mats = gen(n)
        hs = pysqt.HostScene.from_arrays(v9, mi, mats); osc = O.Scene.from_arrays(v9, mi, mats)
        literal = False
    t_build = time.time() - t0
    t0 = time.time(); ctx.upload(hs); t_up = time.time() - t0
    p = pysqt.make_params(W, H, spp, max_depth=depth, seed=0, literal=literal)
    ctx.render_resident(cam, p)
    st = ctx.render_resident(cam, p)
    # parity spot check: 20k random rays + 20k camera rays against the oracle
    osc.make_bih()
    rng = np.random.default_rng(c)
    org = rng.uniform(-1.5, 1.5, (20000, 3)).astype(np.float32); d = rng.normal(size=(20000, 3)).astype(np.float32)
    po = O.make_params(200, 100, 1); o2, d2 = O.make_rays(po, cam)
    org = np.concatenate([org, o2]); d = np.concatenate([d, d2])
    g = ctx.intersect_batch(org, d); w = osc.intersect_batch(org, d)
    exact = bool(np.array_equal(g[0], w[0]) and np.array_equal(g[1].view(np.uint32), w[1].view(np.uint32)) and np.array_equal(g[2].view(np.uint32), w[2].view(np.uint32)))
    r = dict(config=name, tris=hs.n_tris, bih=hs.stats(), spp_run=spp, spp_config=full_spp, depth=depth, device_ms=st["device_ms"],
             rays=st["rays_traced"], samples=st["samples"], mrays_s=st["rays_traced"] / st["device_ms"] / 1e3,
             msamples_s=st["samples"] / st["device_ms"] / 1e3, host_build_s=round(t_build, 2), upload_s=round(t_up, 3),
             parity_40k_rays_bit_exact=exact, hit_fraction=float((g[0] >= 0).mean()))
    print(json.dumps(r), flush=True)
    out.append(r)
