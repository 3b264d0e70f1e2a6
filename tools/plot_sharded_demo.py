"""Block- and component-sharded plot path over NCCL (SURVEY 8e, config C5 at reduced size).  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/plot_sharded_demo.py [--trees 8] [--points 250000]
Every rank holds the plot; blocks and components are dealt round-robin; the labelled voxels and the packed
skeletons are all-gathered.  Rank 0 also runs the plot on its own GPU alone and checks the gathered result
bit for bit, then prints one JSON line."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from smart_tree_b200 import dist as stdist, synth
from smart_tree_b200.data_types.cloud import Cloud
from smart_tree_b200.dataset.augmentations import AugmentationPipeline, CentreCloud
from smart_tree_b200.model.model_inference import ModelInference
from smart_tree_b200.pipeline import Pipeline
from smart_tree_b200.skeleton.skeletonize import Skeletonizer

ap = argparse.ArgumentParser()
ap.add_argument("--trees", type=int, default=8)
ap.add_argument("--points", type=int, default=250_000)
ap.add_argument("--voxel", type=float, default=0.01)
ap.add_argument("--block", type=float, default=0.64)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
rank, world, local = stdist.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
W = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "smart-tree_b200", "model", "weights", "noble-elevator-58_model_weights.pt")
pipe = Pipeline(AugmentationPipeline([CentreCloud()]), ModelInference(None, W, args.voxel, args.block, 0.4, device=dev),
                Skeletonizer(16, 0.02, 32, device=dev), repair_skeletons=True, smooth_skeletons=True, smooth_kernel_size=11,
                prune_skeletons=True, min_skeleton_radius=0.01, min_skeleton_length=0.02, device=dev)
xyz = np.concatenate([synth.make_tree(s, args.points).xyz + np.float32([5.0 * (s % 4), 0, 5.0 * (s // 4)]) for s in range(args.trees)])
cloud = Cloud(xyz=torch.from_numpy(xyz).to(dev), rgb=torch.zeros(len(xyz), 3, device=dev))


def sharded():
    sk = pipe.process_plot_sharded(cloud, rank, world)
    return stdist.gather_skeletons([sk], [0], device=dev), sk


for _ in range(2):
    sharded()
ts = []
for _ in range(args.reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    merged, mine = sharded()
    torch.cuda.synchronize(dev)
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ts.append(float(t) * 1e3)
if rank == 0:
    for _ in range(2):
        single = pipe.process_cloud(cloud=cloud)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    single = pipe.process_cloud(cloud=cloud)
    torch.cuda.synchronize(dev)
    t_single = (time.perf_counter() - t0) * 1e3
    got = {k[1]: v for k, v in merged.items()}
    ok = sorted(got) == list(range(len(single.skeletons)))
    for i, ref in enumerate(single.skeletons):
        g = got.get(i)
        ok = ok and g is not None and sorted(g.branches) == sorted(ref.branches)
        if ok:
            for bid, b in ref.branches.items():
                ok = ok and g.branches[bid].parent_id == b.parent_id and torch.equal(g.branches[bid].xyz, b.xyz)
    print(json.dumps({"workload": f"plot of {args.trees} synthetic trees x {args.points} points, {args.voxel} m voxels, {args.block} m blocks",
                      "points": int(len(xyz)), "n_gpus": world, "ms_sharded": ts, "ms_single_gpu": t_single,
                      "points_per_s_sharded": len(xyz) / (min(ts) / 1e3), "points_per_s_single": len(xyz) / (t_single / 1e3),
                      "skeletons": len(single.skeletons), "branches": sum(len(s.branches) for s in single.skeletons),
                      "blocks": int(pipe.model_inference.last_batch.block_centres.shape[0]) if pipe.model_inference.last_batch is not None else None,
                      "bit_identical_to_single_gpu": bool(ok)}), flush=True)
if world > 1:
    dist.destroy_process_group()
