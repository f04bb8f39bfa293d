"""Regenerates the golden fixtures under tests/golden/.

  python tests/golden/make_golden.py            (in the build container; /root/reference must be mounted for
                                                 the example.png fixture, everything else needs only this repo)

scene_obj_golden.json : produced by the ORACLE (oracle/oracle.c) on data/scene.obj -- BIH statistics, SHA-256 of
    the flattened tree / leaf order / primary hit indices / a small render.  The reference has no golden vectors
    of its own (test/Spec.hs:1-2) and cannot be built here, so these pin the oracle against regressions and pin
    the host C++ builder and the CUDA path to the oracle; they do not pin the oracle to GHC ("parity unpinned").
example_67.npy : /root/reference/render/example.png (540x540, the reference's only output artefact) box-filtered
    to 67x67x3 float32 -- the one weak external pin: silhouette and colour layout of the oracle's render.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    data = os.path.join(ROOT, "data")
    s = O.Scene.load(os.path.join(data, "scene.obj"), data)
    st = s.make_bih()
    cam = O.load_camera(os.path.join(data, "camera"))
    root, nodes, leaf = s.export_bih()
    v9, mi = s.tris()
    org, dirs = O.make_rays(O.make_params(540, 540, 1), cam)
    tri, dist, point, cn = s.intersect_batch(org, dirs, counters=True)
    r = s.render(cam, O.make_params(64, 64, 4, max_depth=3, seed=1, trig=1))
    g = {
        "n_tris": int(s.n_tris), "bih": st, "root_bounds": [float(x) for x in root],
        "sha_tris": sha(v9), "sha_mat_idx": sha(mi), "sha_mats": sha(s.mats()),
        "sha_nodes": sha(nodes), "sha_leaf_order": sha(leaf), "sha_camera": sha(cam),
        "primary_540": {"hits": int((tri >= 0).sum()), "sha_tri": sha(tri), "sha_dist": sha(dist), "sha_point": sha(point),
                        "branch_visits": int(cn[0]), "child_box_tests": int(cn[1]), "own_box_tests": int(cn[2]), "tri_tests": int(cn[3])},
        "render_64x64_4spp_d3_seed1_sqttrig": {"sha_accum": sha(r["accum"]), "sha_rgb8": sha(r["rgb8"]), "rays": r["rays"],
                                               "samples": r["samples"]},
        "philox_kat": {"zeros": [int(x) for x in O.philox([0, 0, 0, 0], [0, 0])],
                       "ones": [int(x) for x in O.philox([0xffffffff] * 4, [0xffffffff] * 2)],
                       "pi": [int(x) for x in O.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])]},
    }
    json.dump(g, open(os.path.join(HERE, "scene_obj_golden.json"), "w"), indent=1, sort_keys=True)
    ref_png = "/root/reference/render/example.png"
    if os.path.exists(ref_png):
        from PIL import Image
        im = np.asarray(Image.open(ref_png).convert("RGB")).astype(np.float32)[:536, :536]
        np.save(os.path.join(HERE, "example_67.npy"), im.reshape(67, 8, 67, 8, 3).mean((1, 3)).astype(np.float32))
    print("golden fixtures written")


if __name__ == "__main__":
    main()
