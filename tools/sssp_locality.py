"""How local is the SSSP graph under a vertex numbering?  For the bench workload (C2): fraction of arcs whose endpoints fall
into different CTA ranges of k_sssp_blob (1024 * G consecutive vertices), fraction of vertices with such an arc, and the
number of range crossings along the deepest shortest path -- for the caller's numbering and for ops.spatial_order with
several cell sizes."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from smart_tree_b200 import ops, synth
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
tr = synth.make_tree(0, int(os.environ.get("POINTS", 1_000_000)))
cloud = Cloud(xyz=torch.from_numpy(tr.xyz).to(dev), rgb=torch.from_numpy(tr.rgb).to(dev))
pipe.process_cloud(cloud=cloud)
last = pipe.skeletonizer.last
order = last["order"]                       # new id -> filtered-cloud vertex id
n_all = int(last["keep"].sum())
m = order.shape[0]
new_id = torch.full((n_all,), -1, dtype=torch.int64, device=dev)
new_id[order] = torch.arange(m, device=dev)
e = last["edges"].long()
eu, ev = new_id[e[:, 0]], new_id[e[:, 1]]
ok = (eu >= 0) & (ev >= 0)
eu, ev = eu[ok], ev[ok]
lc = pipe.labelled_cloud
# medial points of the skeleton vertices, as the skeletonizer computes them
cl = lc.filter_by_class([0]) if hasattr(lc, "filter_by_class") else lc
cl = cl.filter(last["keep"])
med = cl.medial_pts[order].contiguous()
pred = last["pred"].long()                  # component-local == global here (one component)
dist = last["dist"]
comp_of = torch.zeros(m, dtype=torch.int32, device=dev)
deep = int(torch.argmax(torch.where(dist < 1e30, dist, torch.full_like(dist, -1.0))))
path = []
v = deep
predc = pred.cpu().tolist()
while v >= 0 and len(path) < 100000:
    path.append(v)
    v = predc[v]
path = torch.tensor(path, device=dev)
print(json.dumps({"vertices": m, "arcs": int(eu.shape[0]), "deepest_path_hops": len(path) - 1,
                  "edge_len_mean_m": float((med[eu] - med[ev]).norm(dim=1).mean()), "path_len_m": float(dist[deep])}))
def report(name, rank):
    for VB in (1024, 2048, 4096):
        cu, cv = rank[eu] // VB, rank[ev] // VB
        cross = cu != cv
        bv = torch.zeros(m, dtype=torch.bool, device=dev)
        bv[eu[cross]] = True
        pr = rank[path] // VB
        crossings = int((pr[1:] != pr[:-1]).sum())
        print(json.dumps({"order": name, "VB": VB, "cross_arc_frac": round(float(cross.float().mean()), 4),
                          "boundary_vertex_frac": round(float(bv.float().mean()), 4), "deepest_path_crossings": crossings,
                          "ctas": -(-m // VB)}), flush=True)
report("caller", torch.arange(m, device=dev))
for cell in (0.01, 0.02, 0.04, 0.08, 0.16):
    perm, rank = ops.spatial_order(med, comp_of, cell=cell)
    report(f"z-order cell {cell}", rank.long())
# distance order (what a perfect wavefront numbering would give)
rank = torch.empty(m, dtype=torch.long, device=dev)
rank[torch.argsort(dist)] = torch.arange(m, device=dev)
report("by distance", rank)
