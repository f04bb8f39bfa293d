"""Worker for test_multi_gpu.py: one rank per GPU (torchrun).  Checks on real devices that the pixel-group partition
+ ncclReduce gives the 1-GPU frame bit for bit, and that the sample-range partition gives it up to summation order."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "squigly-trace_b200"), HERE]
import pysqt


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    saved = os.dup(1); os.dup2(2, 1)                      # NCCL banner off stdout
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    data = os.path.join(ROOT, "data")
    hs = pysqt.HostScene.load(os.path.join(data, "scene.obj"), data)
    cam = pysqt.load_camera(os.path.join(data, "camera"))
    solo = pysqt.Context(local); solo.upload(hs)          # no communicator: renders the whole frame alone
    grp = pysqt.Context(local); grp.upload(hs)
    box = [pysqt.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    grp.comm_init(rank, world, box[0])
    ok = True
    for flags in (0, pysqt.SQT_F_SPLIT_SAMPLES):
        p = pysqt.make_params(320, 200, 16, max_depth=6, seed=4, flags=flags)
        out = grp.render(cam, p)
        ref = solo.render(cam, pysqt.make_params(320, 200, 16, max_depth=6, seed=4))
        tot = torch.tensor([out["stats"]["samples"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        if rank == 0:
            ok &= int(tot.item()) == 320 * 200 * 16
            if flags == 0:
                ok &= bool(np.array_equal(out["accum"].view(np.uint32), ref["accum"].view(np.uint32)) and np.array_equal(out["rgb8"], ref["rgb8"]))
            else:
                ok &= bool(np.allclose(out["accum"], ref["accum"], rtol=1e-5, atol=1e-5))
    dist.barrier()
    os.dup2(saved, 1)
    if rank == 0:
        print("MULTI_GPU_OK" if ok else "MULTI_GPU_MISMATCH", flush=True)
    grp.close(); solo.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
