"""Sweep of the SSSP schedule knobs (threshold advance rule, polls per barrier chunk, threshold step) on the bench
workload: stage times of skel.sssp / skel.tree_dist per setting.  Any setting gives the same distances (tests)."""
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from smart_tree_b200 import _timing, synth
from smart_tree_b200.data_types.cloud import Cloud
from smart_tree_b200.dataset.augmentations import AugmentationPipeline, CentreCloud
from smart_tree_b200.model.model_inference import ModelInference
from smart_tree_b200.pipeline import Pipeline
from smart_tree_b200.skeleton.skeletonize import Skeletonizer

dev = torch.device("cuda:0")
W = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "smart-tree_b200", "model", "weights", "noble-elevator-58_model_weights.pt")
pipe = Pipeline(AugmentationPipeline([CentreCloud()]), ModelInference(None, W, 0.01, 4, 0.4, device=dev),
                Skeletonizer(16, 0.02, 32, device=dev), repair_skeletons=True, smooth_skeletons=True, smooth_kernel_size=11,
                prune_skeletons=True, min_skeleton_radius=0.01, min_skeleton_length=0.02, device=dev)
tr = synth.make_tree(0, 1_000_000)
cloud = Cloud(xyz=torch.from_numpy(tr.xyz).to(dev), rgb=torch.from_numpy(tr.rgb).to(dev))
for _ in range(2):
    pipe.process_cloud(cloud=cloud)
ref = None
rows = []
settings = [dict(adv=0, passes=32, delta=0.5)]
for adv, passes, delta in itertools.product([1, 0], [64, 128, 256], [0.015, 0.03, 0.06, 0.125]):
    settings.append(dict(adv=adv, passes=passes, delta=delta))
for s in settings:
    os.environ["ST_SSSP_ADVANCE"] = str(s["adv"])
    os.environ["ST_SSSP_PASSES"] = str(s["passes"])
    os.environ["ST_SSSP_DELTA"] = str(s["delta"])
    _timing.enable(True)
    _timing.RECORDS.clear(); _timing.SAMPLES.clear()
    for _ in range(4):
        pipe.process_cloud(cloud=cloud)
    torch.cuda.synchronize()
    med = {k: float(np.median(v)) for k, v in _timing.SAMPLES.items() if k in ("skel.sssp", "skel.tree_dist", "skel.sample_tree")}
    _timing.enable(False)
    d = pipe.skeletonizer.last["dist"].clone()
    if ref is None:
        ref = d
    same = bool(torch.equal(ref, d))
    rows.append({**s, **med, "same_distances": same})
    print(json.dumps(rows[-1]), flush=True)
best = min(rows, key=lambda r: r["skel.sssp"])
print("BEST", json.dumps(best))
