#!/usr/bin/env python
"""Benchmark of the smart-tree hot path on B200 (contract: see the task statement / DESIGN.md).

  python bench.py --gpus N --steps K --warmup W            # B200 arm (N>1: launched under torchrun)
  python bench.py --impl reference --steps K --warmup W    # reference arm = CPU oracle port on host cores

One "step" = one pass of Pipeline.process_cloud over one synthetic tree cloud per rank:
CentreCloud -> block tiling -> voxelise -> sparse UNet -> class filter -> skeletonise ->
prune/repair/smooth, then (N>1) the NCCL gather of the packed skeletons.
Workload = BASELINE.json configs[1]: noble-elevator-58, 1M-point tree, 1 cm voxels, one tree per GPU
(weak scaling, seeds 0..N-1).  metric = points/s (whole job).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WEIGHTS = os.path.join(ROOT, "smart-tree_b200", "model", "weights", "noble-elevator-58_model_weights.pt")
METRIC = "points/sec through sparse-UNet+skeleton"
UNIT = "points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--voxel", type=float, default=0.01)
    ap.add_argument("--cpu-sample-points", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--conv-impl", default=os.environ.get("ST_CONV_IMPL", "auto"))
    return ap.parse_args()


def workload_config(args, world):
    return {"workload": "noble-elevator-58 UNet inference + skeleton, 1M-point synthetic tube tree per GPU, 1cm voxels "
                        "(BASELINE.json configs[1])",
            "points_per_gpu": args.points, "voxel_size": args.voxel, "block_size": 4, "buffer_size": 0.4, "K": 16,
            "weights": "noble-elevator-58", "trees": world, "tree_seed": 0, "parallelism": f"tree-sharded x{world}, NCCL all-gather of packed skeletons",
            "l2": "256 MiB scratch write between timed iterations (inputs < 126 MB L2)"}


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Samples before this call (nvidia-smi start-up, warm-up steps) are not part of the reported clocks."""
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.rows = self.rows[getattr(self, "first", 0):]
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nme, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------ reference arm (CPU oracle)
def cpu_oracle_points_per_s(n_points, voxel, steps=1, warmup=0):
    """Times the oracle port of Pipeline.process_cloud (numpy/scipy/torch-CPU, all host threads the
    libraries use) on a bounded sample of the workload: a synthetic tree of `n_points`."""
    import numpy as np
    import torch

    from oracle import pipeline_ref as P
    from oracle import unet_ref as U
    from smart_tree_b200 import synth
    sd = torch.load(WEIGHTS, map_location="cpu", weights_only=True)
    params = U.to_numpy_params(sd)
    tr = synth.make_tree(0, n_points)
    stages = {}

    def one():
        t0 = time.perf_counter()
        lab = P.infer(params, P.centre_cloud(tr.xyz), tr.rgb, voxel, 4, 0.4)
        t1 = time.perf_counter()
        P.process_cloud(None, None, None, labelled=lab)
        t2 = time.perf_counter()
        stages["unet_s"], stages["skeleton_s"] = t1 - t0, t2 - t1
        return t2 - t0

    for _ in range(warmup):
        one()
    times = [one() for _ in range(max(steps, 1))]
    return n_points / (sum(times) / len(times)), sum(times) / len(times), stages


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    v, sec, stages = cpu_oracle_points_per_s(args.cpu_sample_points, args.voxel, args.steps, min(args.warmup, 1))
    sample = f"synthetic tube tree seed 0, {args.cpu_sample_points} points, {args.voxel} m voxels (same pipeline, bounded sample)"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, **stages},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from smart_tree_b200 import _lib, ops, synth
    from smart_tree_b200 import dist as stdist
    from smart_tree_b200.data_types.cloud import Cloud
    from smart_tree_b200.dataset.augmentations import AugmentationPipeline, CentreCloud
    from smart_tree_b200.model.model_inference import ModelInference
    from smart_tree_b200.pipeline import Pipeline
    from smart_tree_b200.skeleton.skeletonize import Skeletonizer

    rank, world, local = stdist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (B200 arm) needs a CUDA device; there is no CPU fallback. Use --impl reference for the CPU oracle.")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.check(_lib.load().st_device_check(local), "st_device_check")

    mi = ModelInference(None, WEIGHTS, args.voxel, 4, 0.4, device=dev)
    if args.conv_impl != "auto":
        mi.model.conv_impl = args.conv_impl
    pipe = Pipeline(AugmentationPipeline([CentreCloud()]), mi, Skeletonizer(16, 0.02, 32, device=dev), repair_skeletons=True,
                    smooth_skeletons=True, smooth_kernel_size=11, prune_skeletons=True, min_skeleton_radius=0.01,
                    min_skeleton_length=0.02, device=dev)
    # weak scaling: every GPU gets the SAME workload tree (seed 0), so per-GPU work is exactly fixed as N grows
    tr = synth.make_tree(0, args.points)
    h_xyz = torch.from_numpy(tr.xyz).pin_memory()
    h_rgb = torch.from_numpy(tr.rgb).pin_memory()
    d_cloud = Cloud(xyz=h_xyz.to(dev), rgb=h_rgb.to(dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(resident: bool):
        flush.fill_(1)
        if resident:
            cloud = d_cloud
        else:
            cloud = Cloud(xyz=h_xyz.to(dev, non_blocking=True), rgb=h_rgb.to(dev, non_blocking=True))
        sk = pipe.process_cloud(cloud=cloud)
        if world > 1:
            return stdist.gather_skeletons([sk], [rank], device=dev), sk
        return None, sk

    def timed(resident: bool, steps: int):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        per_step = []
        for _ in range(steps):
            ts = time.perf_counter()
            _, sk = step(resident)
            per_step.append((time.perf_counter() - ts) * 1e3)     # the step ends with its result on the host
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        ms = max(ms, 0.0)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), sk, per_step

    # the sampler starts BEFORE the warm-up: nvidia-smi's start-up (NVML initialisation) stalled one timed step in five
    # by ~10 ms when it was launched right at the start of the timed region; only samples taken after mark() count
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step(True)
    import gc
    gc.collect()
    clocks.mark()
    launches0 = ops.LAUNCHES
    dev_ms, wall_ms, sk, steps_res = timed(True, args.steps)
    launches = ops.LAUNCHES - launches0
    stage_t = dict(pipe.timings)
    for _ in range(min(args.warmup, 2)):          # the host-buffer path warms up too (fresh device buffers every step)
        step(False)
    e2e_ms, e2e_wall_ms, sk2, steps_e2e = timed(False, args.steps)
    clk = clocks.stop() if rank == 0 else None
    # the device timeline includes host gaps (the step has host sync points), so device-event time == step time
    ms_per_step = max(dev_ms, wall_ms) / args.steps
    total_points = args.points * world
    value = total_points / (ms_per_step / 1e3)
    e2e_value = total_points / (max(e2e_ms, e2e_wall_ms) / args.steps / 1e3)
    d2h = sum(b.xyz.numel() * 4 + b.radii.numel() * 4 + 16 for s in sk2.skeletons for b in s.branches.values())

    # ---- roofline of the dominant kernel family (3x3x3 gather conv): one instrumented, warm forward
    bb = mi.last_batch
    roof = None
    if bb is not None and bb.feats.shape[0]:
        levels = mi.model.build_levels(bb.coords)
        feats = bb.feats[:, :3].contiguous()
        for _ in range(2):
            mi.model.forward(feats, bb.coords, levels=levels)
        per = {}
        reps = 5
        for _ in range(reps):
            flush.fill_(1)
            ops.conv_profile(True)
            mi.model.forward(feats, bb.coords, levels=levels)
            torch.cuda.synchronize(dev)
            rec = ops.conv_profile(False)
            for cin, cout, taps, n_out, extra, a, b, impl in rec:
                if taps != 27:
                    continue
                key = f"conv{taps}_{cin}x{cout}_{impl}"
                d = per.setdefault(key, {"ms": 0.0, "bytes": 0, "launches": 0})
                d["ms"] += a.elapsed_time(b)
                d["bytes"] += 4 * n_out * (cin + cout + extra)
                d["launches"] += 1
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        table = {k: {"gbps": v["bytes"] / v["ms"] / 1e6, "us_per_launch": v["ms"] / v["launches"] * 1e3,
                     "bytes_per_launch": v["bytes"] // v["launches"], "launches_per_forward": v["launches"] // reps,
                     "ms_per_forward": v["ms"] / reps} for k, v in per.items()}
        top = max(table, key=lambda k: table[k]["ms_per_forward"])
        tot_ms = sum(v["ms"] for v in per.values()) / reps
        tot_b = sum(v["bytes"] for v in per.values()) / reps
        traffic, traffic_src = None, None
        try:          # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/README.md)
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if top in tj:
                traffic, traffic_src = tj[top]["dram_bytes_per_launch"], "profiles/ncu_conv_r1c.csv (dram__bytes_read.sum + dram__bytes_write.sum)"
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": top, "achieved": table[top]["gbps"], "peak": peak, "unit": "GB/s",
                "frac": table[top]["gbps"] / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_kind,
                "all_gather_convs": {"achieved": tot_b / tot_ms / 1e6, "frac": tot_b / tot_ms / 1e6 / peak, "ms_per_forward": tot_ms},
                "per_kernel": table, "level_voxels": [lv.n for lv in levels]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, sec, stages = cpu_oracle_points_per_s(args.cpu_sample_points, args.voxel)
            cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": f"synthetic tube tree seed 0, {args.cpu_sample_points} points, {args.voxel} m voxels, 1 pass ({sec:.1f} s)",
                   **stages}
        except Exception as exc:  # the oracle is only a reported baseline
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(args, world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h_xyz.numel() * 4 + h_rgb.numel() * 4),
                        "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
                "stages_ms": {k: v * 1e3 for k, v in stage_t.items()},
                "step_ms": {"resident": [round(v, 2) for v in steps_res], "host_buffers": [round(v, 2) for v in steps_e2e]},
                "result": {"skeletons": len(sk.skeletons), "branches": sum(len(s.branches) for s in sk.skeletons),
                           "voxels": int(bb.feats.shape[0]) if bb is not None else 0}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION; stdout must carry the JSON line only
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
