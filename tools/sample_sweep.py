"""Sweep of the sample_tree round knobs (window size, scan steps) on the bench workload; same result for any value."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from smart_tree_b200 import _lib, _timing, synth
from smart_tree_b200.data_types.cloud import Cloud
from smart_tree_b200.dataset.augmentations import AugmentationPipeline, CentreCloud
from smart_tree_b200.model.model_inference import ModelInference
from smart_tree_b200.pipeline import Pipeline
from smart_tree_b200.skeleton.skeletonize import Skeletonizer
from smart_tree_b200.util.digest import skeleton_digest

dev = torch.device("cuda:0")
W = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "smart-tree_b200", "model", "weights", "noble-elevator-58_model_weights.pt")
pipe = Pipeline(AugmentationPipeline([CentreCloud()]), ModelInference(None, W, 0.01, 4, 0.4, device=dev),
                Skeletonizer(16, 0.02, 32, device=dev), repair_skeletons=True, smooth_skeletons=True, smooth_kernel_size=11,
                prune_skeletons=True, min_skeleton_radius=0.01, min_skeleton_length=0.02, device=dev)
tr = synth.make_tree(0, 1_000_000)
cloud = Cloud(xyz=torch.from_numpy(tr.xyz).to(dev), rgb=torch.from_numpy(tr.rgb).to(dev))
lib = _lib.load()
lib.st_debug_sample_batch_stats.argtypes = [C.c_void_p]
ref = None
for win, steps, cl, sep in ((128, 4, 16, 1.0), (128, 4, 16, 2.0), (128, 4, 16, 1.5), (128, 4, 16, 0.8), (64, 4, 16, 1.0), (256, 4, 16, 1.0), (512, 4, 16, 1.0),
                           (128, 4, 8, 1.0), (128, 4, 4, 1.0)):
    if True:
        os.environ["ST_SAMPLE_SEP"] = str(sep)
        os.environ["ST_SAMPLE_WIN"] = str(win)
        os.environ["ST_SAMPLE_SCAN_STEPS"] = str(steps)
        os.environ["ST_SAMPLE_CLUSTER"] = str(cl)
        sk = pipe.process_cloud(cloud=cloud)
        _timing.enable(True)
        _timing.RECORDS.clear(); _timing.SAMPLES.clear()
        for _ in range(5):
            sk = pipe.process_cloud(cloud=cloud)
        torch.cuda.synchronize()
        ms = float(np.median(_timing.SAMPLES["skel.sample_tree"]))
        _timing.enable(False)
        bst = (C.c_ulonglong * 16)()
        lib.st_debug_sample_batch_stats(bst)
        phs = (C.c_ulonglong * 8)()
        lib.st_debug_sample_round_phases.argtypes = [C.c_void_p]
        lib.st_debug_sample_round_phases(phs)
        print(json.dumps(dict(zip(["kc_B_select", "kc_C_routes", "kc_C_fill_windows", "kc_barrier1", "kc_E_inputs", "kc_verdict", "kc_F_commit", "kc_barrier2"],
                                  [int(v) // 1000 for v in phs]))))
        dg = skeleton_digest(sk.skeletons)["topology"]
        ref = ref or dg
        print(json.dumps({"win": win, "sep": sep, "cluster": cl, "sample_tree_ms": round(ms, 3), "rounds": int(bst[0]), "batches": int(bst[2]),
                          "offered": int(bst[3]), "accepted": int(bst[4]), "cut_gap": int(bst[9]), "kcycles_A": int(bst[12]) // 1000,
                          "kcycles_BC": int(bst[13]) // 1000, "kcycles_D": int(bst[14]) // 1000, "kcycles_batches": int(bst[10]) // 1000,
                          "same_topology": dg == ref}), flush=True)
