"""Stage-by-stage comparison of the CUDA skeletoniser with the oracle at benchmark size (BASELINE config C2 by default):
outlier mask -> kNN edges / weights -> components -> SSSP predecessors -> tree distances -> sample_tree paths.  The
oracle gets the CUDA network's own labelled cloud.  Reports the FIRST mismatch of every stage instead of asserting, so
one run localises a divergence.   python tools/parity_stages.py [--points N] [--voxel V] [--seed S] [--weights NAME]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import skeleton_ref as S
from smart_tree_b200 import synth
from smart_tree_b200.data_types.cloud import Cloud
from smart_tree_b200.dataset.augmentations import AugmentationPipeline, CentreCloud
from smart_tree_b200.model.model_inference import ModelInference
from smart_tree_b200.pipeline import Pipeline
from smart_tree_b200.skeleton.skeletonize import Skeletonizer

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=1_000_000)
ap.add_argument("--voxel", type=float, default=0.01)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--weights", default="noble-elevator-58")
ap.add_argument("--out", default="gpurun_out/parity_stages.json")
args = ap.parse_args()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda:0")
W = os.path.join(ROOT, "smart-tree_b200", "model", "weights", f"{args.weights}_model_weights.pt")
pipe = Pipeline(AugmentationPipeline([CentreCloud()]), ModelInference(None, W, args.voxel, 4, 0.4, device=dev),
                Skeletonizer(16, 0.02, 32, device=dev), repair_skeletons=False, device=dev)
tr = synth.make_tree(args.seed, args.points)
cloud = Cloud(xyz=torch.from_numpy(tr.xyz).to(dev), rgb=torch.from_numpy(tr.rgb).to(dev))
rep = {"args": vars(args)}


def first_diff(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return {"shape_gpu": list(a.shape), "shape_ref": list(b.shape)}
    d = np.nonzero((a != b).reshape(len(a), -1).any(1))[0] if a.size else np.zeros(0, int)
    if len(d) == 0:
        return None
    i = int(d[0])
    return {"mismatches": int(len(d)), "first": i, "gpu": np.asarray(a[i]).tolist(), "ref": np.asarray(b[i]).tolist()}


for mode in ("batched", "sequential"):
    if mode == "sequential":
        os.environ["ST_SAMPLE_SEQUENTIAL"] = "1"
    else:
        os.environ.pop("ST_SAMPLE_SEQUENTIAL", None)
    pipe.process_cloud(cloud=cloud)
    lc = pipe.labelled_cloud
    sel = (lc.class_l.reshape(-1) == 0)
    xyz = lc.xyz[sel].cpu().numpy()
    mv = lc.medial_vector[sel].cpu().numpy()
    last = pipe.skeletonizer.last
    r = {}
    if mode == "batched":
        t0 = time.perf_counter()
        medial = xyz + mv
        radius = np.sqrt((mv[:, 0] * mv[:, 0] + mv[:, 1] * mv[:, 1]) + mv[:, 2] * mv[:, 2])
        keep = S.outlier_removal(medial, radius, 8)
        r["outlier_keep"] = first_diff(last["keep"].cpu().numpy(), keep)
        x2, m2, r2 = xyz[keep], medial[keep], radius[keep]
        edges, w = S.nn_graph(m2, np.maximum(r2, np.float32(0.02)), 16)
        ge, gw = last["edges"].cpu().numpy(), last["edge_weights"].cpu().numpy()
        r["edges"] = first_diff(ge, edges)
        r["edge_weights"] = first_diff(gw, w) if r["edges"] is None else "skipped"
        comps = S.connected_components(len(x2), edges, 32)
        off = last["comp_off"].cpu().numpy()
        order = last["order"].cpu().numpy()
        r["n_components"] = [int(len(off) - 1), len(comps)]
        ref_sk = S.skeletonize(xyz, mv, 16, 0.02, 32)
        rep["oracle_seconds"] = round(time.perf_counter() - t0, 1)
    for c, sk in enumerate(ref_sk[:4]):
        rc = {}
        lo, hi = int(off[c]), int(off[c + 1])
        rc["vertex_ids"] = first_diff(order[lo:hi], sk.vertex_ids)
        rc["pred"] = first_diff(last["pred"][lo:hi].cpu().numpy(), sk.preds)
        rc["tree_dist"] = first_diff(last["tree_dist"][lo:hi].cpu().numpy(), sk.distances)
        nb, npth = int(last["comp_n_branches"][c]), int(last["comp_n_path"][c])
        rc["n_branches"] = [nb, len(sk.branches)]
        gl = last["branch_len"][lo:lo + nb].cpu().numpy()
        gp = last["branch_parent"][lo:lo + nb].cpu().numpy()
        rl = np.array([len(b.path) for b in sk.branches])
        rp = np.array([b.parent_id for b in sk.branches])
        k = min(nb, len(sk.branches))
        rc["branch_len"] = first_diff(gl[:k], rl[:k])
        rc["branch_parent"] = first_diff(gp[:k], rp[:k])
        paths = last["path"][lo:lo + npth].cpu().numpy()
        rpaths = np.concatenate([b.path for b in sk.branches]) if sk.branches else np.zeros(0, np.int64)
        k = min(len(paths), len(rpaths))
        rc["paths"] = first_diff(paths[:k], rpaths[:k])
        if rc["branch_len"] is not None or rc["branch_parent"] is not None:
            i = min([d["first"] for d in (rc["branch_len"], rc["branch_parent"]) if d and "first" in d] or [0])
            rc["around_first_bad_branch"] = {"index": i, "gpu_len": gl[max(i - 2, 0):i + 3].tolist(), "ref_len": rl[max(i - 2, 0):i + 3].tolist(),
                                             "gpu_parent": gp[max(i - 2, 0):i + 3].tolist(), "ref_parent": rp[max(i - 2, 0):i + 3].tolist()}
        r[f"component_{c}"] = rc
    rep[mode] = r
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump(rep, open(args.out, "w"), indent=1)
print(json.dumps(rep, indent=1))
