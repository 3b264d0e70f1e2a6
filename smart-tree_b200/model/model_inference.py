"""ModelInference with the reference's constructor and forward
(/root/reference/smart_tree/model/model_inference.py:11-100)."""
from pathlib import Path

import torch

from ..data_types.cloud import Cloud
from ..dataset.dataset import load_dataloader
from ..engine import SmartTreeEngine
from .._timing import section


def load_model(model_path, weights_path, device=torch.device("cuda:0"), bn_eps=1e-4):
    """The reference unpickles the whole module (`model_path`) and then loads `weights_path`; the
    pickle needs spconv/cumm/omegaconf importable, so only the plain state_dict is read here and
    the architecture is derived from it.  `model_path` is accepted for signature compatibility."""
    sd = torch.load(f"{weights_path}", map_location="cpu", weights_only=True)
    return SmartTreeEngine(sd, device=device, eps=bn_eps)


class ModelInference:
    def __init__(self, model_path: Path, weights_path: Path, voxel_size: float, block_size: float, buffer_size: float,
                 num_workers=8, batch_size=4, device=torch.device("cuda:0"), verbose=False, bn_eps=1e-4):
        self.device = torch.device(device)
        self.verbose = verbose
        self.voxel_size = voxel_size
        self.block_size = block_size
        self.buffer_size = buffer_size
        self.num_workers = num_workers
        self.batch_size = batch_size
        self.model = load_model(model_path, weights_path, self.device, bn_eps)
        self.last_batch = None

    @torch.no_grad()
    def forward(self, cloud: Cloud, return_masked=True) -> Cloud:
        """Tiles the cloud into blocks, voxelises, runs the network on all blocks as one batch
        (eval-mode BatchNorm makes the result independent of how blocks are batched) and returns the
        labelled voxel cloud (on the device; the reference returns it on the CPU and the pipeline
        moves it straight back, pipeline.py:63)."""
        if cloud.xyz.device.type != "cuda":
            cloud = cloud.to_device(self.device)
        with section("infer.blocks"):
            ds = load_dataloader(cloud, self.voxel_size, self.block_size, self.buffer_size, self.num_workers, self.batch_size)
        clouds, all_preds = [], []
        for bb in ds.voxelize_chunks():
            self.last_batch = bb
            if bb.feats.shape[0] == 0:
                continue
            with section("infer.levels"):
                levels = self.model.build_levels(bb.coords)
            with section("infer.unet"):
                preds = self.model.forward(bb.feats[:, :3], bb.coords, levels=levels, fused_outputs=True)
            with section("infer.tail"):
                lc = Cloud(xyz=bb.feats[:, :3], rgb=bb.feats[:, 3:6], medial_vector=preds["medial_vector"],
                           class_l=preds["class_idx"].long().unsqueeze(1))
                self.last_preds = preds
                clouds.append(lc.filter(bb.mask) if return_masked else lc)
        if not clouds:
            z = torch.zeros(0, 3, device=cloud.xyz.device)
            return Cloud(xyz=z, rgb=z.clone(), medial_vector=z.clone(), class_l=torch.zeros(0, 1, dtype=torch.int64, device=z.device))
        if len(clouds) == 1:
            return clouds[0]
        return Cloud(xyz=torch.cat([c.xyz for c in clouds]), rgb=torch.cat([c.rgb for c in clouds]),
                     medial_vector=torch.cat([c.medial_vector for c in clouds]), class_l=torch.cat([c.class_l for c in clouds]))

    @staticmethod
    def from_cfg(cfg):
        return ModelInference(model_path=cfg.model_path, weights_path=cfg.weights_path, voxel_size=cfg.voxel_size,
                              block_size=cfg.block_size, buffer_size=cfg.buffer_size, num_workers=cfg.num_workers,
                              batch_size=cfg.batch_size)
