"""Per-op timing of SmartTreeEngine.build_levels on the bench workload (synchronising timers).  GPU box only."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from smart_tree_b200 import ops, synth
from smart_tree_b200.data_types.cloud import Cloud
from smart_tree_b200.dataset.augmentations import CentreCloud
from smart_tree_b200.dataset.dataset import SingleTreeInference

dev = torch.device("cuda:0")
tr = synth.make_tree(0, 1_000_000)
cloud = CentreCloud()(Cloud(xyz=torch.from_numpy(tr.xyz).to(dev), rgb=torch.from_numpy(tr.rgb).to(dev)))
bb = SingleTreeInference(cloud, 0.01, 4, 0.4).voxelize_all()
coords = bb.coords.contiguous()
acc = {}


def timed(name, fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t) * 1e3
    return r


REPS = 10
for rep in range(REPS + 2):
    if rep == 2:
        acc.clear()
    perm = timed("morton_perm", lambda: ops.morton_perm(coords))
    c = timed("gather coords", lambda: coords[perm.long()].contiguous())
    for lvl in range(4):
        table = timed(f"L{lvl} hash", lambda: ops.CoordTable(c))
        timed(f"L{lvl} subm_map", lambda: ops.subm_map(c, table))
        if lvl < 3:
            oc = timed(f"L{lvl} strided_coords", lambda: ops.strided_coords(c, morton=True))
            t2 = timed(f"L{lvl} hash(next)", lambda: ops.CoordTable(oc))
            timed(f"L{lvl} strided_maps", lambda: ops.strided_maps(c, oc, t2))
            c = oc
tot = 0
for k, v in acc.items():
    print(f"{k:22s} {v / REPS:7.3f} ms")
    tot += v / REPS
print("sum", tot)
