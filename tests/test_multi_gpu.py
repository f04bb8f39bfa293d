"""Multi-GPU path on real devices (skipped on boxes with fewer than 2 GPUs; the CPU-side equivalent is test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu


def test_two_ranks_nccl_reduce_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(HERE, "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "MULTI_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_single_process_group_matches_single_gpu():
    """sqt_comm_init_all + sqt_upload_scene_group + sqt_render_group: the single-process path INTEGRATION.md gives the Haskell
    host (Main.hs:43 -> render).  The frame must equal the 1-GPU frame bit for bit on every available power-of-two group
    size, and a group in which one rank cannot start (no scene uploaded on it) must return an error on ALL ranks instead of
    leaving the others waiting in ncclReduce."""
    import numpy as np
    import torch
    import pysqt
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs 2 GPUs")
    data = os.path.join(os.path.dirname(HERE), "data")
    hs = pysqt.HostScene.load(os.path.join(data, "scene.obj"), data)
    cam = pysqt.load_camera(os.path.join(data, "camera"))
    p = pysqt.make_params(320, 200, 16, max_depth=6, seed=4)
    solo = pysqt.Context(0); solo.upload(hs)
    ref = solo.render(cam, p)
    solo.close()
    n = 2
    while n <= n_dev:
        g = pysqt.Group(list(range(n)))
        g.upload(hs)
        out = g.render(cam, p)
        assert np.array_equal(out["accum"].view(np.uint32), ref["accum"].view(np.uint32)), "group of %d" % n
        assert np.array_equal(out["rgb8"], ref["rgb8"])
        g.close()
        n *= 2
    # one rank without a scene: everybody gets an error, nobody hangs
    g = pysqt.Group([0, 1])
    g.ctxs[0].upload(hs)
    with pytest.raises(pysqt.SqtError):
        g.render(cam, p)
    g.upload(hs)                                  # and the group is still usable afterwards
    out = g.render(cam, p)
    assert np.array_equal(out["accum"].view(np.uint32), ref["accum"].view(np.uint32))
    g.close()
