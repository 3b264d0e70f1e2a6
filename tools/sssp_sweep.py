"""Sweep of the SSSP schedule knobs on the bench workload (C2): the CTA-local kernel (k_sssp_blob: vertices per CTA,
local polls per round, optional per-CTA threshold step) against the register-resident near-far kernel (k_sssp).  Prints the
median stage time of skel.sssp per setting and whether the distances are identical (they must be: any schedule reaches
the same fp32 fixed point)."""
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
tr = synth.make_tree(0, int(os.environ.get("POINTS", 1_000_000)))
cloud = Cloud(xyz=torch.from_numpy(tr.xyz).to(dev), rgb=torch.from_numpy(tr.rgb).to(dev))
KNOBS = ("ST_CC_SAMPLE", "ST_SSSP_PAIRS", "ST_CC_NO_SAMPLE", "ST_CC_PRELINK", "ST_SSSP_BLOB_FLAGS", "ST_SSSP_LOCAL", "ST_SSSP_PASSES", "ST_SSSP_DELTA", "ST_SSSP_NO_LOCAL", "ST_SSSP_LOCAL_G", "ST_SSSP_NLOCAL", "ST_SSSP_BLOB_DELTA", "ST_SSSP_SPATIAL")
# default: the 2-D sweep of threshold step x polls per barrier behind profiles/sssp_schedule_sweep_r2.txt, then the variants
settings = [{}]
for d in ("0.06", "0.09", "0.125", "0.18", "0.25"):
    for ps in ("16", "32", "48", "64", "96"):
        settings.append({"ST_SSSP_DELTA": d, "ST_SSSP_PASSES": ps})
settings += [{"ST_SSSP_PAIRS": "0"}, {"ST_SSSP_SPATIAL": "1"}, {"ST_SSSP_LOCAL": "1"}, {"ST_CC_NO_SAMPLE": "1"}, {"ST_CC_SAMPLE": "16"}]
ref = None
rows = []
for s in settings:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(s)
    pipe.skeletonizer = Skeletonizer(16, 0.02, 32, device=dev)      # (reads ST_SSSP_SPATIAL / ST_SSSP_LOCAL)
    pipe.process_cloud(cloud=cloud)
    _timing.enable(True)
    _timing.RECORDS.clear(); _timing.SAMPLES.clear()
    for _ in range(5):
        pipe.process_cloud(cloud=cloud)
    torch.cuda.synchronize()
    med = {k: round(float(np.median(v)), 3) for k, v in _timing.SAMPLES.items() if k in ("skel.sssp", "skel.components", "skel.emit")}
    _timing.enable(False)
    d = pipe.skeletonizer.last["dist"].clone()
    p = pipe.skeletonizer.last["pred"].clone()
    if ref is None:
        ref = (d, p)
    import ctypes as C
    from smart_tree_b200 import _lib, ops
    st = (C.c_ulonglong * 11)()
    lib = _lib.load()
    lib.st_debug_sssp_blob_stats.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    if "ST_SSSP_LOCAL" in s:
        lib.st_debug_sssp_blob_stats(C.c_void_p(ops.LAST_SSSP_CTL.data_ptr()), int(d.shape[0]), st)
    med["relax_accepted_remote_epochs_rounds"] = list(st)[:5]
    ncta = max(1, -(-int(d.shape[0]) // (960 * int(s.get("ST_SSSP_LOCAL_G", 2)))))
    med["avg_us_per_cta_local_global_barrier_idle_total"] = [round(x / ncta / 1965.0, 1) for x in list(st)[5:10]]
    med["busiest_cta_busy_us"] = round(st[10] / 1965.0, 1)
    rows.append({**s, **med, "same_distances_and_preds": bool(torch.equal(ref[0], d) and torch.equal(ref[1], p))})
    print(json.dumps(rows[-1]), flush=True)
print("BEST", json.dumps(min(rows, key=lambda r: r["skel.sssp"])))
