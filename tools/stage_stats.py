"""Per-stage wall times of one Pipeline.process_cloud step on the bench workload, plus the
sample_tree phase counters and the SSSP sweep count.  Run on the GPU box:
    python tools/stage_stats.py [--points 1000000] [--gt-medial]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from smart_tree_b200 import _lib, _timing, ops, synth
from smart_tree_b200.data_types.cloud import Cloud
from smart_tree_b200.dataset.augmentations import AugmentationPipeline, CentreCloud
from smart_tree_b200.model.model_inference import ModelInference
from smart_tree_b200.pipeline import Pipeline
from smart_tree_b200.skeleton.skeletonize import Skeletonizer

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=1_000_000)
ap.add_argument("--voxel", type=float, default=0.01)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
dev = torch.device("cuda:0")
W = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "smart-tree_b200", "model", "weights", "noble-elevator-58_model_weights.pt")
pipe = Pipeline(AugmentationPipeline([CentreCloud()]), ModelInference(None, W, args.voxel, 4, 0.4, device=dev),
                Skeletonizer(16, 0.02, 32, device=dev), repair_skeletons=True, smooth_skeletons=True, smooth_kernel_size=11,
                prune_skeletons=True, min_skeleton_radius=0.01, min_skeleton_length=0.02, device=dev)
tr = synth.make_tree(0, args.points)
cloud = Cloud(xyz=torch.from_numpy(tr.xyz).to(dev), rgb=torch.from_numpy(tr.rgb).to(dev))
for _ in range(2):
    pipe.process_cloud(cloud=cloud)
_timing.enable(True)
t0 = time.perf_counter()
for _ in range(args.reps):
    sk = pipe.process_cloud(cloud=cloud)
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / args.reps * 1e3
rec = {k: v / args.reps for k, v in _timing.RECORDS.items()}
spread = {k: [round(min(v), 2), round(sorted(v)[len(v) // 2], 2), round(max(v), 2)] for k, v in _timing.SAMPLES.items() if max(v) > 1.0}
_timing.enable(False)
t0 = time.perf_counter()
for _ in range(args.reps):
    pipe.process_cloud(cloud=cloud)
torch.cuda.synchronize()
untimed = (time.perf_counter() - t0) / args.reps * 1e3
stats = (C.c_ulonglong * 8)()
lib = _lib.load()
lib.st_debug_sample_stats.argtypes = [C.c_void_p]
lib.st_debug_sample_stats(stats)
iters = (C.c_int * 8192)()
lib.st_debug_sample_iters.argtypes = [C.c_void_p]
lib.st_debug_sample_iters(iters)
import numpy as _np
_np.save("gpurun_out/sample_iters.npy", _np.array(iters[:]).reshape(1024, 8))
# depth of the predecessor tree (hops) by pointer jumping on the host
import numpy as np
last = pipe.skeletonizer.last
try:
    pred = last["pred"].cpu().numpy().astype(np.int64)
    off = last["comp_off"].cpu().numpy().astype(np.int64)
    basev = np.repeat(off[:-1], off[1:] - off[:-1])
    jump = np.where(pred >= 0, pred + basev, -1)
    depth = (jump >= 0).astype(np.int64)
    for _ in range(20):
        has = jump >= 0
        j = np.maximum(jump, 0)
        depth = depth + np.where(has, depth[j], 0)
        jump = np.where(has, jump[j], jump)
    max_depth = int(depth.max())
except Exception as exc:
    max_depth = repr(exc)
sst = (C.c_uint * 5)()
lib.st_debug_sssp_stats.argtypes = [C.c_void_p, C.c_void_p]
lib.st_debug_sssp_stats(C.c_void_p(ops.LAST_SSSP_CTL.data_ptr()), sst)
bst = (C.c_ulonglong * 16)()
lib.st_debug_sample_batch_stats.argtypes = [C.c_void_p]
lib.st_debug_sample_batch_stats(bst)
bnames = ["rounds", "long_iterations", "batches", "members_offered", "accepted", "skipped", "cut_long_member", "cut_start_touched",
          "cut_route_or_parent_touched", "cut_passed_over_entry_unclaimed", "cycles_batches", "cycles_long", "cycles_A_window_scan",
          "cycles_BC_select_routes", "cycles_D_claim", "cluster_size"]
batch_stats = {n: int(bst[i]) for i, n in enumerate(bnames)}
names = ["find", "trace", "claim", "resolve", "finish"]
cyc = {n: stats[i] for i, n in enumerate(names)}
last = pipe.skeletonizer.last
print(json.dumps({"ms_per_step_with_timers": tot, "ms_per_step": untimed, "sections_ms": rec, "min_med_max_ms": spread,
                  "sample_tree_cycles": cyc, "sample_tree_batches": batch_stats, "sample_tree_iterations": stats[5], "sample_tree_path_vertices": stats[6],
                  "cluster_size": stats[7], "sssp_chunks_evals_improvements_x_lanemode": list(sst), "sssp_tree_max_depth_hops": max_depth, "branches": sum(len(s.branches) for s in sk.skeletons), "components": last["n_components"],
                  "skeleton_vertices": int(last["order"].shape[0]), "edges": int(last["edges"].shape[0]),
                  "voxels": int(pipe.model_inference.last_batch.feats.shape[0])}, indent=1))
