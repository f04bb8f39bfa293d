"""Throughput of sqt_intersect_batch (batched Scene.intersect) on data/scene.obj: random incoherent rays and the
a coherent 1920x1080 pinhole fan.  Device time from the library's events (sqt_stats.device_ms).
usage: python tools/bench_intersect.py [n_rays]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "squigly-trace_b200")]
import pysqt

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
data = os.path.join(ROOT, "data")
hs = pysqt.HostScene.load(os.path.join(data, "scene.obj"), data)
ctx = pysqt.Context(0)
ctx.upload(hs)
rng = np.random.default_rng(1)
org = rng.uniform(-2.5, 2.5, (n, 3)).astype(np.float32)
d = rng.normal(size=(n, 3)).astype(np.float32)
# coherent bundle: a 1920x1080 pinhole fan from outside the scene towards its centre
lo, hi = hs.root[:3].astype(np.float64), hs.root[3:].astype(np.float64)
c = 0.5 * (lo + hi); eye = c + np.array([0.3, -1.0, 0.25]) * float((hi - lo).max())
fwd = (c - eye) / np.linalg.norm(c - eye); right = np.cross(fwd, [0, 0, 1.0]); right /= np.linalg.norm(right); up = np.cross(right, fwd)
gx, gy = np.meshgrid(np.linspace(-0.6, 0.6, 1920), np.linspace(-0.34, 0.34, 1080))
pd = (fwd[None, :] + gx.reshape(-1, 1) * right[None, :] + gy.reshape(-1, 1) * up[None, :]).astype(np.float32)
po = np.tile(eye.astype(np.float32), (len(pd), 1))
reps = max(1, n // len(po))
po = np.tile(po, (reps, 1)); pd = np.tile(pd, (reps, 1))
for name, o_, d_ in (("random", org, d), ("pinhole 1080p", po, pd)):
    for i in range(3):
        tri, dist, point, st = ctx.intersect_batch(o_, d_, want_stats=True)
    print("%-14s %9d rays: device %.2f ms -> %.1f Mrays/s (hit rate %.2f)" % (name, len(o_), st["device_ms"], len(o_) / st["device_ms"] / 1e3, float((tri >= 0).mean())))
ctx.close()
