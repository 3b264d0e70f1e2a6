"""One Pipeline.process_cloud step of the bench workload (C2) -- the command profiled under ncu:
    ncu --set full --clock-control none -k regex:'k_conv|k_sssp|k_sample_tree_c|k_knn|k_heads' -c 40 -o gpurun_out/ncu_step python tools/one_step.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from smart_tree_b200 import synth
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
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    sk = pipe.process_cloud(cloud=cloud)
torch.cuda.synchronize()
print("branches", sum(len(s.branches) for s in sk.skeletons))
