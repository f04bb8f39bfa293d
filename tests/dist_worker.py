"""Worker for test_dist_gloo.py: one rank of a world_size-N gloo group.  Each rank renders its pixel-group share
with the host build of the kernel logic (same work_to_pixel / work_items code the CUDA kernels run), the shares are
summed with a reduce to rank 0 exactly like ncclReduce does on the GPUs, and rank 0 checks the sum against the
single-rank frame bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "squigly-trace_b200"), HERE]
import pysqt
from common import Emu, bits


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    data = os.path.join(ROOT, "data")
    hs = pysqt.HostScene.load(os.path.join(data, "scene.obj"), data)
    cam = pysqt.load_camera(os.path.join(data, "camera"))
    e = Emu(hs)
    ok = True
    for flags in (0, pysqt.SQT_F_SPLIT_SAMPLES):
        p = pysqt.make_params(80, 48, 4, max_depth=4, seed=6, flags=flags)
        mine = e.render(cam, p, rank=rank, world=world)
        t = torch.from_numpy(mine["accum"].copy())
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
        samples = torch.tensor([mine["samples"]], dtype=torch.int64)
        dist.all_reduce(samples)
        if rank == 0:
            full = e.render(cam, p, rank=0, world=1)
            assert int(samples.item()) == full["samples"] == 80 * 48 * 4
            if flags == 0:      # disjoint pixel groups: x + 0 is exact, the frame equals the 1-rank frame bit for bit
                ok &= bool(np.array_equal(bits(t.numpy()), bits(full["accum"])))
                own = (mine["accum"].reshape(-1, 3) != 0).any(1)
                groups = np.arange(80 * 48) // 32
                ok &= bool(np.all(groups[own] % world == 0))
            else:               # sample ranges: same samples, different summation order
                ok &= bool(np.allclose(t.numpy(), full["accum"], rtol=1e-6, atol=1e-6))
    if rank == 0:
        print("DIST_OK" if ok else "DIST_MISMATCH")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
