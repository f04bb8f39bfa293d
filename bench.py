#!/usr/bin/env python
"""bench.py -- headline benchmark of the squigly-trace B200 backend (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): data/scene.obj + scene.sq + camera at 1920x1080, 1024 spp, max 8 bounces.
A step = one full render of that frame through the hot path.  Metric = Mrays/s, where a ray is one closest-hit
query (Lib.hs:131) ACTUALLY executed -- primary hits are traced once per pixel and reused by its samples, paths end
at surfaces with surfColor = 0; both are exact (bit-identical image), and only executed rays are counted.

  value : rays of all ranks / max-over-ranks CUDA-event time of the K timed steps (scene resident in HBM)
  e2e   : same through the host-buffer C ABI (sqt_upload_scene + sqt_render): H2D of the scene, D2H of the RGB8 frame
  roofline : FP32 (non-fused issue rate; the bit-exact path may not contract to FMA) of the dominant kernel k_paths
  cpu_baseline : the oracle (C port of the reference algorithm) on the host cores, bounded sample of the same frame

`--impl reference` times that oracle alone (the Haskell reference cannot be built: no GHC in the image).
Multi-GPU: launched by torchrun, one rank per GPU; pixel groups are partitioned over ranks and the accumulation
buffers summed with ncclReduce inside the library (bit-identical to 1 GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "squigly-trace_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

WIDTH, HEIGHT, SPP, DEPTH, SEED = 1920, 1080, 1024, 8, 0
DATA = os.path.join(ROOT, "data")
WORKLOAD = "data/scene.obj 1920x1080 1024spp depth8"
NCU_DRAM_BYTES_PER_LAUNCH = 205336064 + 591681024        # profiles/r01_k_paths_v7_pool.txt


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.3 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return None
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows if r[2].replace(".", "").isdigit())}


def algorithmic_fp32_ops(st):
    """SURVEY 8(d), guard-aware: 3 reciprocals per ray; 24 ops per child-box test (6 sub, 6 mul, 10 min/max, 2 cmp);
    per triangle test 14 ops up to the `a` guard, +9 and the divide up to the `u` guard, +16 up to the `v` guard,
    +6 up to the `t` guard, +14 and the sqrt for an accepted hit (Geometry.hs:117-142 with edges precomputed).
    Shading (RNG, trig, bounce) is NOT counted."""
    return (3 * st["rays_traced"] + 24 * st["child_box_tests"] + 14 * st["tri_tests"] + 10 * st["mt_pass_a"]
            + 16 * st["mt_pass_u"] + 6 * st["mt_pass_v"] + 15 * st["mt_accept"])


def algorithmic_fp32_ops_upper(st):
    """SURVEY 8(d) upper bound: every triangle test charged in full (59 add/mul + div + sqrt)."""
    return 3 * st["rays_traced"] + 24 * st["child_box_tests"] + 61 * st["tri_tests"]


def cpu_baseline(spp_probe=1, target_s=15.0):
    """Oracle (oracle/oracle.c, all host threads) on a bounded sample: the full 1920x1080 frame, depth 8, few spp."""
    from oracle import oracle as O
    sc = O.Scene.load(os.path.join(DATA, "scene.obj"), DATA)
    sc.make_bih()
    cam = O.load_camera(os.path.join(DATA, "camera"))
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    r = sc.render(cam, O.make_params(WIDTH, HEIGHT, spp_probe, max_depth=DEPTH, seed=SEED, trig=0), want_rgb8=False, nthreads=cores)
    t1 = time.perf_counter() - t0
    spp = int(max(1, min(64, round(target_s / max(t1, 1e-3) * spp_probe))))
    if spp > spp_probe:
        t0 = time.perf_counter()
        r = sc.render(cam, O.make_params(WIDTH, HEIGHT, spp, max_depth=DEPTH, seed=SEED, trig=0), want_rgb8=False, nthreads=cores)
        t1 = time.perf_counter() - t0
    else:
        spp = spp_probe
    return {"value": r["rays"] / t1 / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": "full 1920x1080 frame, depth 8, %d spp (%d rays, %.1f s); C port of the reference algorithm, "
                      "libm trig, no primary-hit reuse (Lib.hs:81-87)" % (spp, r["rays"], t1),
            "msamples_per_s": r["samples"] / t1 / 1e6, "seconds": t1, "rays": r["rays"], "spp": spp}


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU algorithm on the host cores (oracle port; GHC is not in the image)."""
    if rank != 0:
        return
    from oracle import oracle as O
    sc = O.Scene.load(os.path.join(DATA, "scene.obj"), DATA)
    sc.make_bih()
    cam = O.load_camera(os.path.join(DATA, "camera"))
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    sc.render(cam, O.make_params(WIDTH, HEIGHT, 1, max_depth=DEPTH, seed=SEED, trig=0), want_rgb8=False, nthreads=cores)
    probe = time.perf_counter() - t0
    budget = 150.0 / max(1, args.steps + args.warmup)            # whole run within a few minutes
    spp = int(max(1, min(64, budget / max(probe, 1e-3))))
    p = O.make_params(WIDTH, HEIGHT, spp, max_depth=DEPTH, seed=SEED, trig=0)
    for _ in range(args.warmup):
        sc.render(cam, p, want_rgb8=True, nthreads=cores)
    rays = 0
    samples = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = sc.render(cam, p, want_rgb8=True, nthreads=cores)
        rays += r["rays"]; samples += r["samples"]
    dt = time.perf_counter() - t0
    v = rays / dt / 1e6
    sample = "each step = full 1920x1080 frame, depth 8, %d spp (bounded sample of the 1024 spp job)" % spp
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "reference scene data/scene.obj",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "msamples_per_s": samples / dt / 1e6, "gpu_launches": 0,
        "note": "C port of the reference algorithm (oracle/oracle.c), all host threads; the Haskell binary cannot be built here",
    }))


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import pysqt

    # NCCL prints its version banner with printf on stdout when NCCL_DEBUG is set in the environment: keep fd 1
    # pointed at stderr until the one JSON line is ready.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    hs = pysqt.HostScene.load(os.path.join(DATA, "scene.obj"), DATA)
    cam = pysqt.load_camera(os.path.join(DATA, "camera"))
    ctx = pysqt.Context(local_rank)
    ctx.upload(hs)
    if world > 1:
        box = [pysqt.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(rank, world, box[0])
    p = pysqt.make_params(WIDTH, HEIGHT, SPP, max_depth=DEPTH, seed=SEED)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident steps -------------------------------------------------------------
    for _ in range(args.warmup):
        ctx.render_resident(cam, p)
        flush.zero_()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    wall0 = time.time()
    dev_ms = paths_ms = 0.0
    rays = samples = launches = 0
    per_step = []
    for _ in range(args.steps):
        st = ctx.render_resident(cam, p)            # CUDA events on the library's stream bracket every kernel
        dev_ms += st["device_ms"]; paths_ms += st["paths_ms"]
        rays += st["rays_traced"]; samples += st["samples"]; launches += st["kernel_launches"]
        per_step.append(st["device_ms"])
        flush.zero_()                                # L2 flush between timed iterations (outside the events)
    barrier()
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if sampler else None
    T = reduce_max(dev_ms)
    total_rays = reduce_sum(rays)
    total_samples = reduce_sum(samples)
    total_launches = reduce_sum(launches)
    value = total_rays / (T * 1e-3) / 1e6

    # ---- end to end through the host-buffer ABI ---------------------------------------------
    scene_bytes = hs.nodes.nbytes + hs.tris.nbytes + hs.mats.nbytes + 24
    for _ in range(1):
        ctx.upload(hs); ctx.render(cam, p, want_accum=False)
    barrier()
    e_rays = 0
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        ctx.upload(hs)                               # H2D: nodes + triangles + materials
        out = ctx.render(cam, p, want_accum=False)   # D2H: RGB8 frame on rank 0
        e_rays += out["stats"]["rays_traced"]; d2h = out["stats"]["d2h_bytes"]
    barrier()
    e_T = reduce_max(time.perf_counter() - t0)
    e_value = reduce_sum(e_rays) / e_T / 1e6

    # ---- roofline of the dominant kernel (k_paths) -------------------------------------------
    # Numerator = the REFERENCE ALGORITHM's work for exactly this ray set: counted by the instrumented
    # kernels with the leaf culling off, where branch visits and triangle tests equal the oracle's counters one for
    # one (tests/test_gpu_parity.py).  The work the default kernels execute (culling on) is reported next to it.
    roof = None
    fp32_peak = ctx.fp32_peak_gops()
    l2_peak = ctx.l2_bandwidth_gbs()
    pc = pysqt.make_params(WIDTH, HEIGHT, SPP, max_depth=DEPTH, seed=SEED, flags=pysqt.SQT_F_COUNT_WORK)
    ctx.set_leaf_cull(False)
    ref_cnt = ctx.render_resident(cam, pc)
    ctx.set_leaf_cull(True)
    exe_cnt = ctx.render_resident(cam, pc)
    barrier()
    if rank == 0:
        ops = algorithmic_fp32_ops(ref_cnt)         # this rank's share; k_paths time is this rank's too
        k_ms = paths_ms / args.steps
        achieved = ops / (k_ms * 1e-3) / 1e12
        mem_bytes = 16 * ref_cnt["branch_visits"] + 36 * ref_cnt["tri_tests"]
        keys = ("rays_traced", "branch_visits", "child_box_tests", "tri_tests", "mt_pass_a", "mt_pass_u", "mt_pass_v", "mt_accept", "leaves_culled")
        roof = {"bound": "fp32", "kernel": "k_paths_pool", "achieved": achieved, "peak": fp32_peak / 1e3, "unit": "TFLOP/s",
                "frac": achieved / (fp32_peak / 1e3), "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one 64-spp k_paths_pool launch (the bench issues 16 such "
                                  "launches per frame), ncu --set full, profiles/r01_k_paths_v7_pool.txt: sample-buffer writes and first-touch "
                                  "fills; scene, pool-slot stacks and path state are L1/L2 resident",
                "peak_source": "measured live: non-fused FADD/FMUL issue rate of this GPU (sqt_measure_fp32_peak); FMA contraction is "
                               "forbidden on the bit-exact path, so this is the FP32 ceiling (MEASURED_PEAKS.json has no FP32 figure)",
                "kernel_ms": k_ms, "kernel_ms_note": "all k_paths + k_accumulate launches of one frame (one pair per 64-spp round)",
                "kernel_share_of_step": paths_ms / dev_ms,
                "algorithmic": dict({k: ref_cnt[k] for k in keys}, fp32_ops=ops, fp32_ops_upper=algorithmic_fp32_ops_upper(ref_cnt),
                                    node_tri_bytes=mem_bytes, note="reference algorithm, guard-aware (SURVEY 8d)"),
                "executed": dict({k: exe_cnt[k] for k in keys}, fp32_ops=algorithmic_fp32_ops(exe_cnt),
                                 note="what the default kernels do: conservative leaf culling skips triangle tests"),
                "frac_executed": algorithmic_fp32_ops(exe_cnt) / (k_ms * 1e-3) / 1e12 / (fp32_peak / 1e3),
                "l1l2": {"achieved_gbs": mem_bytes / (k_ms * 1e-3) / 1e9, "l2_peak_gbs_measured": l2_peak,
                         "hbm_peak_gbs_measured": _hbm_peak(),
                         "note": "16 B per branch visit + 36 B per triangle test; the scene (0.36 MB) is L1/L2 resident, HBM traffic ~0"}}

    cpu = cpu_baseline() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps({
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": T / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "reference scene data/scene.obj (6238 triangles), camera data/camera",
            "config": {"workload": WORKLOAD, "width": WIDTH, "height": HEIGHT, "spp": SPP, "max_depth": DEPTH,
                       "index_convention": "corrected (rows=1080, cols=1920)", "partition": "pixel groups of 32, round-robin over ranks",
                       "l2_flush": "256 MiB device write between steps"},
            "msamples_per_s": total_samples / (T * 1e-3) / 1e6,
            "rays_per_step": total_rays / args.steps, "per_step_ms": per_step,
            "e2e": {"value": e_value, "unit": "Mrays/s", "h2d_bytes_per_step": scene_bytes + 104, "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e_T / args.steps * 1e3},
            "gpu_launches": int(total_launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "wall_ms_per_step": (wall1 - wall0) / args.steps * 1e3,
        }))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def _hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # not under torchrun: relaunch as one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr",
               "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup)] + (["--no-cpu-baseline"] if args.no_cpu_baseline else [])
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
