#!/usr/bin/env python
"""bench.py -- Mpaths/s (and Mrays/s) of the hw5 path tracer hot path on practice5_dragon_100k.

One "step" = one full Scene::Render pass of the scene (W x H x SAMPLES camera paths, RAY_DEPTH 6)
spp-sharded over the N ranks (each rank renders SAMPLES/N samples per pixel with disjoint Philox
streams; the float accumulation buffers are summed onto rank 0 with an NCCL reduce), so scaling is
STRONG: the total work per step is fixed.

  value        device-resident throughput: scene already in HBM, timed region = render kernels
               (+ the NCCL reduce for N > 1), CUDA events, max over ranks.
  e2e          same metric through the C-ABI with host buffers: every step re-uploads the
               flattened scene (H2D), renders, resolves to 8-bit and copies the image back (D2H).
  roofline     dominant kernel (k_traverse = BVH part of Scene::RayIntersection); see DESIGN.md "Roofline".
  cpu_baseline the UNMODIFIED reference (oracle/_ref/librefprobe.so -> Scene::Sample loop) on the
               host cores, on a bounded sample of the same scene.

`--impl reference` times that same reference arm as its own JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = "practice5_dragon_100k"
METRIC = "Mpaths/s"


def scene_file(name):
    p = os.path.join(ROOT, "scenes", name + ".txt")
    if not os.path.exists(p):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_scenes.py")])
    return p


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {}
        for n in dir(nv):
            if n.startswith("nvmlClocksThrottleReason") or n.startswith("nvmlClocksEventReason"):
                v = getattr(nv, n)
                if isinstance(v, int) and v:
                    names.setdefault(v, n.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", ""))
        while not self.stop_flag.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit and bit & (bit - 1) == 0:
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def result(self):
        self.stop_flag.set()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        reasons = sorted(r for r in self.reasons if r not in ("GpuIdle", "None", "ApplicationsClocksSetting"))
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples)}


def reference_arm(args, rank):
    """The reference's own CPU implementation of the path on the host cores (bounded sample)."""
    if rank != 0:
        return 0
    import orclib
    if not orclib.have_ref():
        # the reference did not compile here -> the oracle port (documented in DESIGN.md)
        backend, kind = orclib.oracle(), "port"
    else:
        backend, kind = orclib.ref(), "reference"
    s = orclib.Scene(backend, scene_file(args.scene))
    full = (s.width, s.height, s.samples)
    w, h, spp = args.ref_width, args.ref_height, args.ref_spp
    s.override(w, h, spp)
    cores = os.cpu_count() or 1

    def one():
        t0 = time.perf_counter()
        if kind == "reference":
            s.ref_render_linear(nthreads=cores)
        else:
            s.render_sum(1, 0, spp, nthreads=cores)
        return time.perf_counter() - t0

    for _ in range(args.warmup if args.impl == "reference" else 0):
        one()
    steps = args.steps if args.impl == "reference" else 1
    times = [one() for _ in range(steps)]
    paths = w * h * spp
    value = paths / (sum(times) / len(times)) / 1e6
    sample = "%s at %dx%d, %d spp, depth %d (%d paths/step) of the %dx%d, %d spp frame" % (
        args.scene, w, h, spp, s.ray_depth, paths, full[0], full[1], full[2])
    base = {"value": value, "unit": METRIC, "cores": cores, "kind": kind, "sample": sample}
    s.close()
    if args.impl != "reference":
        return base
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.scene, "width": full[0], "height": full[1], "spp": full[2], "sample": sample},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default=WORKLOAD)
    ap.add_argument("--spp", type=int, default=-1, help="override SAMPLES of the scene")
    ap.add_argument("--width", type=int, default=-1)
    ap.add_argument("--height", type=int, default=-1)
    ap.add_argument("--batch-paths", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-width", type=int, default=64)
    ap.add_argument("--ref-height", type=int, default=64)
    ap.add_argument("--ref-spp", type=int, default=8)
    ap.add_argument("--traversal", type=int, default=0)
    ap.add_argument("--pipeline", action="store_true",
                    help="keep two frames in flight on two streams (measured: +0.8 %% before the cone nodes, +0.1 %% after; off by default)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return reference_arm(args, rank)

    import torch
    import torch.distributed as dist
    import raytracing_course_b200 as rtc

    if not torch.cuda.is_available() or rtc.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    scene = rtc.Scene(path=scene_file(args.scene), device=local_rank)
    scene.override(args.width, args.height, args.spp, -1)
    scene.set_traversal(args.traversal)
    if args.batch_paths:
        scene.set_batch_paths(args.batch_paths)
    W, H, spp, depth = scene.width, scene.height, scene.samples, scene.ray_depth
    # spp sharding: rank r renders samples [lo, hi)
    lo, hi = rtc.shard_samples(spp, rank, world)
    npix = W * H
    # --pipeline: frames two deep, frame i+1 starts on the other stream (own accumulation buffer) while the last
    # kernels of frame i drain, so the tail of a persistent k_traverse launch overlaps the next frame's first
    # kernels.  Every frame is still rendered, reduced and complete inside the timed region.  Default: one frame
    # at a time.
    accums = [torch.zeros(npix * 3, dtype=torch.float32, device=dev) for _ in range(2)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    rgb = torch.zeros(npix * 3, dtype=torch.uint8, device=dev)
    host_rgb = torch.zeros(npix * 3, dtype=torch.uint8).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def render_step(step, e2e, pipelined=False):
        h2d = 0
        if not pipelined:
            accum = accums[0]
            if e2e:
                h2d = scene.upload()                      # host scene arrays -> HBM
            accum.zero_()
            scene.render_accumulate(accum.data_ptr(), seed=1000 + step, sample_begin=lo, sample_count=hi - lo)
            if world > 1:
                dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
            if e2e and rank == 0:
                scene.render_resolve(accum.data_ptr(), spp, rgb.data_ptr())
                host_rgb.copy_(rgb, non_blocking=False)   # D2H of the 8-bit frame
            return h2d
        slot = step & 1
        accum, st = accums[slot], streams[slot]
        with torch.cuda.stream(st):
            accum.zero_()
            scene.render_accumulate(accum.data_ptr(), seed=1000 + step, sample_begin=lo, sample_count=hi - lo,
                                    stream=st.cuda_stream)
            if world > 1:
                dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
        return 0

    def timed(nsteps, e2e, first_step, pipelined=False):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        a.record()
        for st in streams:
            st.wait_stream(cur)
        h2d = 0
        for i in range(nsteps):
            h2d = render_step(first_step + i, e2e, pipelined)
        for st in streams:
            cur.wait_stream(st)
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), h2d

    # ---- warm-up (also sizes the wavefront buffers), then an untimed instrumented pass
    for i in range(max(args.warmup, 3)):
        render_step(i, False, args.pipeline)
    barrier()
    scene.reset_counters()
    scene.set_profiling(kernel_events=False, count_visits=True)
    render_step(0, False)
    stats = scene.counters()

    # ---- timed region: device-resident (no per-kernel events: they cost ~1 % of a step)
    scene.set_profiling(False, False)
    scene.reset_counters()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total, _ = timed(args.steps, False, 100, pipelined=args.pipeline)
    clocks = sampler.result()
    cnt = scene.counters()
    # ---- the same steps again with CUDA events around every kernel (roofline durations)
    scene.set_profiling(kernel_events=True, count_visits=False)
    render_step(0, False)   # the single-stream pass may need larger wavefront buffers: allocate them untimed
    scene.profile()
    ms_profiled, _ = timed(args.steps, False, 100)
    prof = scene.profile()
    scene.set_profiling(False, False)

    # ---- timed region: end to end through host buffers
    render_step(0, True)
    e2e_ms, h2d_bytes = timed(args.steps, True, 200)

    paths_rank = npix * (hi - lo) * args.steps
    t_paths = torch.tensor([float(paths_rank), float(cnt["rays"]), float(cnt["launches"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_paths, op=dist.ReduceOp.SUM)
    total_paths, total_rays, total_launches = [float(x) for x in t_paths.tolist()]
    ms_per_step = ms_total / args.steps
    value = total_paths / (ms_total * 1e-3) / 1e6
    mrays = total_rays / (ms_total * 1e-3) / 1e6
    e2e_value = total_paths / (e2e_ms * 1e-3) / 1e6

    if rank == 0:
        # roofline of the dominant kernel, per launch, rank 0 (algorithmic bytes: DESIGN.md "Roofline")
        rays0 = max(cnt["rays"], 1)
        trav0 = max(cnt["traversed_rays"], 1)
        visits_per_ray = stats["index_node_visits"] / max(stats["traversed_rays"], 1)
        tests_per_ray = stats["prim_tests"] / max(stats["traversed_rays"], 1)
        planes = scene.nprims - scene.nbvh
        node_bytes = scene.stats().get("index_node_bytes", 64)   # 96: child boxes + child cones + refs
        alg = {  # (bytes per unit, units processed in the timed region, kernel)
            "traverse": (4 + 32 + 4 + 4 + node_bytes * visits_per_ray + 48 * tests_per_ray, trav0,
                         "k_traverse" if args.traversal == 0 else "k_extend_reftree"),
            "pre": (32 + 8 + node_bytes + 32 * planes + 4 * (trav0 / rays0), rays0, "k_pre"),
            "shade": (64 + 4 + 48 + 32 + 64 + 12, rays0, "k_shade"),
        }
        top = max(alg, key=lambda k: prof[k]["ms"])
        bytes_per_unit, units, kname = alg[top]
        launches = max(prof[top]["launches"], 1)
        per_launch_ms = prof[top]["ms"] / launches
        per_launch_bytes = bytes_per_unit * units / launches
        peak, peak_src = measured_peaks()
        achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        kernel_ms = {k: v["ms"] / args.steps for k, v in prof.items()}
        traffic = None  # DRAM bytes per launch of that kernel from the committed ncu pass (profiles/)
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("workload") == args.scene and kname in tj.get("kernels", {}):
                traffic = tj["kernels"][kname]["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "bytes_per_unit": bytes_per_unit, "unit_name": "ray entering the BVH" if top == "traverse" else "ray",
                    "units_per_launch": units / launches,
                    "index_node_visits_per_traversed_ray": visits_per_ray, "prim_tests_per_traversed_ray": tests_per_ray,
                    "traversed_fraction_of_rays": trav0 / rays0,
                    "avg_launch_ms": per_launch_ms, "launches": launches,
                    "share_of_step": prof[top]["ms"] / ms_profiled if ms_profiled > 0 else None,
                    "profiled_ms_per_step": ms_profiled / args.steps,
                    "kernel_ms_per_step": kernel_ms,
                    "munits_per_s_in_kernel": units / (prof[top]["ms"] * 1e-3) / 1e6 if prof[top]["ms"] > 0 else None}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = reference_arm(args, 0)
        line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.scene, "width": W, "height": H, "spp": spp, "ray_depth": depth,
                           "paths_per_step": npix * spp, "triangles": scene.nbvh, "parallelism": "spp-shard x%d" % world,
                           "l2": "wavefront state (%.0f MB/step) exceeds the 126 MB L2" % (min(npix * (hi - lo), 1 << 23) * 148 / 1e6),
                           "traversal": "index" if args.traversal == 0 else "reftree",
                           "frames_in_flight": 2 if args.pipeline else 1},
                "mrays_per_s": mrays, "rays_per_path": total_rays / max(total_paths, 1),
                "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": int(h2d_bytes),
                        "d2h_bytes_per_step": int(npix * 3), "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(total_launches), "fallback_rays": cnt["fallback_rays"],
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    scene.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
