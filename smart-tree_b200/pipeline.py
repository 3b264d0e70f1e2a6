"""Pipeline with the reference's constructor and process_cloud
(/root/reference/smart_tree/pipeline.py:13-106): load -> preprocess -> infer -> class filter ->
skeletonise -> prune / repair / smooth.  Viewing and open3d export are out of scope (GUI);
`save_outputs` writes the reference's npz skeleton schema instead."""
from __future__ import annotations

import time
from pathlib import Path

import numpy as np
import torch

from .data_types.cloud import Cloud
from .data_types.tree import DisjointTreeSkeleton
from .util.file import load_cloud, save_skeleton
from ._timing import section


class Pipeline:
    def __init__(self, preprocessing, model_inference, skeletonizer, repair_skeletons=False, smooth_skeletons=False,
                 smooth_kernel_size=0, prune_skeletons=False, min_skeleton_radius=0.0, min_skeleton_length=1000,
                 view_model_output=False, view_skeletons=False, save_outputs=False, save_path="/", branch_classes=[0],
                 cmap=[[1, 0, 0], [0, 1, 0]], device=torch.device("cuda:0")):
        self.preprocessing = preprocessing
        self.model_inference = model_inference
        self.skeletonizer = skeletonizer
        self.repair_skeletons = repair_skeletons
        self.smooth_skeletons = smooth_skeletons
        self.smooth_kernel_size = smooth_kernel_size
        self.prune_skeletons = prune_skeletons
        self.min_skeleton_radius = min_skeleton_radius
        self.min_skeleton_length = min_skeleton_length
        self.view_model_output = view_model_output
        self.view_skeletons = view_skeletons
        self.save_outputs = save_outputs
        self.save_path = save_path
        self.branch_classes = list(branch_classes)
        self.cmap = np.asarray(cmap)
        self.device = torch.device(device)
        self.timings = {}
        self.labelled_cloud = None

    def process_cloud(self, path: Path = None, cloud: Cloud = None) -> DisjointTreeSkeleton:
        cloud = load_cloud(Path(path), pin_memory=self.device.type == "cuda") if path is not None else cloud      # pinned: the copy below is asynchronous
        t = {}
        sync = (lambda: torch.cuda.synchronize(self.device)) if self.device.type == "cuda" else (lambda: None)
        t0 = time.perf_counter()
        with section("pipe.preprocess"):
            cloud = cloud.to_device(self.device, non_blocking=True)
            cloud = self.preprocessing(cloud)
        sync(); t["preprocess"] = time.perf_counter() - t0; t0 = time.perf_counter()
        lc: Cloud = self.model_inference.forward(cloud).to_device(self.device)      # pipeline.py:63
        sync(); t["inference"] = time.perf_counter() - t0; t0 = time.perf_counter()
        self.labelled_cloud = lc
        with section("pipe.filter_class"):
            branch_cloud = lc.filter_by_class(self.branch_classes)                  # pipeline.py:68
        skeleton = self.skeletonizer.forward(branch_cloud, post=self._fused_post())  # pipeline.py:71 (+ fused 76-88)
        sync(); t["skeleton"] = time.perf_counter() - t0; t0 = time.perf_counter()
        self.post_process(skeleton)
        sync(); t["post_process"] = time.perf_counter() - t0
        self.timings = t
        if self.view_model_output or self.view_skeletons:
            print("smart_tree_b200: viewers are out of scope (open3d GUI); ignoring view_* flags")
        if self.save_outputs:
            for s in skeleton.skeletons:
                save_skeleton(s, f"{self.save_path}/skeleton_{s._id}.npz")
        return skeleton

    def process_plot_sharded(self, cloud: Cloud, rank: int, world: int, exchange=None):
        """Multi-GPU plot (BASELINE config C5): every rank holds the plot, labels blocks
        rank, rank+world, ... and all-gathers the labelled voxels (dist.gather_labelled, restored to block order so the
        merged cloud is identical to the single-GPU one); every rank then builds the same neighbour graph and
        components and skeletonises components rank, rank+world, ... of the size-ordered list.  Returns this rank's
        DisjointTreeSkeleton (skeleton ids = global component indices); gather them with dist.gather_skeletons.
        `exchange(parts) -> list of all ranks' parts` replaces the collective (tests run the ranks in one process)."""
        from . import dist as stdist
        t = {}
        sync = (lambda: torch.cuda.synchronize(self.device)) if self.device.type == "cuda" else (lambda: None)
        t0 = time.perf_counter()
        cloud = self.preprocessing(cloud.to_device(self.device))
        sync(); t["preprocess"] = time.perf_counter() - t0; t0 = time.perf_counter()
        lc = self.model_inference.forward(cloud, shard=(rank, world)).to_device(self.device)
        sync(); t["inference_own_blocks"] = time.perf_counter() - t0; t0 = time.perf_counter()
        part = stdist.labelled_part(lc, self.model_inference.last_voxel_block)
        parts = exchange(part) if exchange is not None else stdist.gather_labelled(part, device=self.device)
        lc = stdist.merge_labelled(parts, self.device)
        sync(); t["gather_labelled_voxels"] = time.perf_counter() - t0; t0 = time.perf_counter()
        self.labelled_cloud = lc
        branch_cloud = lc.filter_by_class(self.branch_classes)
        skeleton = self.skeletonizer.forward(branch_cloud, post=self._fused_post(), shard=(rank, world))
        sync(); t["skeleton_redundant_graph_plus_own_components"] = time.perf_counter() - t0
        self.timings = t
        done = getattr(skeleton, "post_applied", None) or {}
        if not done:
            # object-level post-processing: only the globally first skeleton is pruned (quirk C-18)
            if self.prune_skeletons and skeleton.skeletons and skeleton.skeletons[0]._id == 0:
                skeleton.prune(min_length=self.min_skeleton_length, min_radius=self.min_skeleton_radius)
            if self.repair_skeletons:
                skeleton.repair(device=self.device)
            if self.smooth_skeletons:
                skeleton.smooth(self.smooth_kernel_size)
        elif self.smooth_skeletons and "smooth" not in done:
            skeleton.smooth(self.smooth_kernel_size)
        return skeleton

    def _fused_post(self):
        """The part of post_process (pipeline.py:76-88) the skeletoniser can run on the device in its branch
        assembly launch.  Steps are order-dependent (prune -> repair -> smooth), so fusing stops at the first
        step the kernel cannot do (an even smoothing kernel)."""
        post = {}
        if self.prune_skeletons:
            post["prune"] = (float(self.min_skeleton_radius), float(self.min_skeleton_length))
        if self.repair_skeletons:
            post["repair"] = True
        if self.smooth_skeletons and int(self.smooth_kernel_size) % 2 == 1:
            post["smooth"] = int(self.smooth_kernel_size)
        return post or None

    def post_process(self, skeleton: DisjointTreeSkeleton):
        done = getattr(skeleton, "post_applied", None) or {}
        if done:
            if self.smooth_skeletons and "smooth" not in done:
                with section("post.smooth"):
                    skeleton.smooth(self.smooth_kernel_size)
            return
        if self.prune_skeletons:
            with section("post.prune"):
                skeleton.prune(min_length=self.min_skeleton_length, min_radius=self.min_skeleton_radius)
        if self.repair_skeletons:
            with section("post.repair"):
                skeleton.repair(device=self.device)
        if self.smooth_skeletons:
            with section("post.smooth"):
                skeleton.smooth(self.smooth_kernel_size)
