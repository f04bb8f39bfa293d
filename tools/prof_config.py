"""One BASELINE config rendered at a reduced spp (the launch ncu captures): python tools/prof_config.py CONFIG_INDEX SPP [N_TRIS]
CONFIG_INDEX = position in BASELINE.json `configs` (0..4).  Prints device time, Mrays/s, rays per sample."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.environ.get("SQT_PKG") or os.path.join(ROOT, "squigly-trace_b200"))   # SQT_PKG: a frozen build (A/B)
import pysqt
from pysqt import scenes

ci, spp = int(sys.argv[1]), int(sys.argv[2])
cfg = dict(scenes.CONFIGS[ci])
if len(sys.argv) > 3:
    cfg["n_tris"] = int(sys.argv[3])
data = os.path.join(ROOT, "data")
t0 = time.time()
arr = scenes.config_arrays(cfg)
hs = pysqt.HostScene.load(os.path.join(data, "scene.obj"), data) if arr is None else pysqt.HostScene.from_arrays(*arr)
t1 = time.time()
cam = pysqt.load_camera(os.path.join(data, "camera"))
ctx = pysqt.Context(0)
ctx.upload(hs)
t2 = time.time()
p = pysqt.make_params(cfg["width"], cfg["height"], spp, max_depth=cfg["depth"], seed=0, literal=cfg["literal"])
print("%s: %d tris, bih %s, host build %.1f s, upload %.2f s" % (cfg["name"], hs.n_tris, hs.stats(), t1 - t0, t2 - t1), flush=True)
for i in range(2):
    st = ctx.render_resident(cam, p)
    print("render %d spp: device %.2f ms paths %.2f ms primary %.3f ms rays %d samples %d -> %.1f Mrays/s %.1f Msamples/s %.2f rays/sample" % (
        spp, st["device_ms"], st["paths_ms"], st["primary_ms"], st["rays_traced"], st["samples"], st["rays_traced"] / st["device_ms"] / 1e3,
        st["samples"] / st["device_ms"] / 1e3, st["rays_traced"] / max(1, st["samples"])), flush=True)
ctx.close()
