"""Build one BASELINE config's scene once and render it under several SQT_POOL_TUNE settings (one context per setting).
usage: python tools/prof_config_multi.py CONFIG_INDEX SPP "b,t,c" "b,t,c" ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "squigly-trace_b200"))
import pysqt
from pysqt import scenes

ci, spp = int(sys.argv[1]), int(sys.argv[2])
cfg = dict(scenes.CONFIGS[ci])
data = os.path.join(ROOT, "data")
arr = scenes.config_arrays(cfg)
hs = pysqt.HostScene.load(os.path.join(data, "scene.obj"), data) if arr is None else pysqt.HostScene.from_arrays(*arr)
cam = pysqt.load_camera(os.path.join(data, "camera"))
p = pysqt.make_params(cfg["width"], cfg["height"], spp, max_depth=cfg["depth"], seed=0, literal=cfg["literal"])
for tune in sys.argv[3:]:
    os.environ["SQT_POOL_TUNE"] = tune
    ctx = pysqt.Context(0)
    ctx.upload(hs)
    ctx.render_resident(cam, p)
    st = min((ctx.render_resident(cam, p) for _ in range(2)), key=lambda s: s["device_ms"])
    print("config %d tune %-9s: device %.2f ms -> %.1f Mrays/s" % (ci + 1, tune, st["device_ms"], st["rays_traced"] / st["device_ms"] / 1e3), flush=True)
    ctx.close()
