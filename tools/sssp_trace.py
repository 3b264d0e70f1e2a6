"""Where does the SSSP wave spend its time?  k_sssp_blob records (ST_SSSP_BLOB_FLAGS=4) the time of the last accepted
improvement of every vertex; along the deepest shortest path this gives the arrival time of the FINAL value hop by hop:
hops inside a CTA range vs hops across a range boundary."""
import ctypes as C
import json
import os
import sys

os.environ["ST_SSSP_BLOB_FLAGS"] = "4"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from smart_tree_b200 import _lib, ops, synth
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
for _ in range(3):
    pipe.process_cloud(cloud=cloud)
torch.cuda.synchronize()
last = pipe.skeletonizer.last
m = int(last["order"].shape[0])
lib = _lib.load()
tf = (C.c_uint * m)()
lib.st_debug_sssp_tfinal.argtypes = [C.c_void_p, C.c_int64]
lib.st_debug_sssp_tfinal(tf, m)
tf = np.frombuffer(tf, dtype=np.uint32).astype(np.int64) * 8 / 1000.0      # us
lc = pipe.labelled_cloud.filter_by_class([0]).filter(last["keep"])
med = lc.medial_pts[last["order"]].contiguous()
perm, rank = ops.spatial_order(med, torch.zeros(m, dtype=torch.int32, device=dev))
rank = rank.cpu().numpy()
pred = last["pred"].cpu().numpy()
dist = last["dist"].cpu().numpy()
t = tf[rank]                                 # per caller vertex
deep = int(np.argmax(np.where(dist < 1e30, dist, -1)))
path = []
v = deep
while v >= 0:
    path.append(v)
    v = pred[v]
path = np.array(path[::-1])
tp = t[path] - t[path].min()
VB = 1920
cta = rank[path] // VB
cross = cta[1:] != cta[:-1]
dt = np.diff(tp)
rows = [{"hop": int(i), "cta": int(cta[i]), "dist_m": round(float(dist[path[i]]), 3), "t_us": round(float(tp[i]), 1)} for i in range(0, len(path), 8)]
print(json.dumps({"hops": len(path) - 1, "total_us": float(tp.max()), "crossings": int(cross.sum()),
                  "dt_us_cross_mean": float(dt[cross].mean()), "dt_us_cross_median": float(np.median(dt[cross])), "dt_us_local_mean": float(dt[~cross].mean()),
                  "dt_us_local_median": float(np.median(dt[~cross])), "dt_sum_cross": float(dt[cross].sum()), "dt_sum_local": float(dt[~cross].sum()),
                  "negative_dt": int((dt < 0).sum())}))
for r in rows:
    print(json.dumps(r))

el = (C.c_ulonglong * 4096)()
lib.st_debug_sssp_epoch_log.argtypes = [C.c_void_p]
lib.st_debug_sssp_epoch_log(el)
el = np.frombuffer(el, dtype=np.uint64).reshape(1024, 4).astype(np.int64)
t_prev = None
for e in range(60):
    if el[e, 0] == 0:
        break
    print(json.dumps({"epoch": e, "advancer_cta": int(el[e, 1]), "its_rounds": int(el[e, 2]),
                      "epoch_us": None if t_prev is None else round((el[e, 0] - t_prev) / 1000.0, 1),
                      "advancer_awake_us_in_epoch": round((el[e, 0] - el[e, 3]) / 1000.0, 1)}))
    t_prev = el[e, 0]
