"""ModelInference with the reference's constructor and forward
(/root/reference/smart_tree/model/model_inference.py:11-100)."""
from pathlib import Path

import torch

from .. import ops

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
    def forward(self, cloud: Cloud, return_masked=True, shard=None) -> Cloud:
        """Tiles the cloud into blocks, voxelises, runs the network on all blocks as one batch
        (eval-mode BatchNorm makes the result independent of how blocks are batched) and returns the
        labelled voxel cloud (on the device; the reference returns it on the CPU and the pipeline
        moves it straight back, pipeline.py:63)."""
        if cloud.xyz.device.type != "cuda":
            cloud = cloud.to_device(self.device)
        with section("infer.blocks"):
            ds = load_dataloader(cloud, self.voxel_size, self.block_size, self.buffer_size, self.num_workers, self.batch_size)
            if shard is not None:          # (rank, world): this rank's blocks only; self.last_voxel_block = their global ids
                ds.keep_shard(*shard)
        clouds, all_preds, vblocks = [], [], []
        chunk0 = 0
        for bb in ds.voxelize_chunks():
            self.last_batch = bb
            if bb.feats.shape[0] == 0:
                chunk0 += ds.MAX_BLOCKS_PER_LAUNCH
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
                if shard is not None:
                    vb = ds.block_global[(bb.coords[:, 0].long() + chunk0)]
                    vblocks.append(vb[bb.mask] if return_masked else vb)
            chunk0 += ds.MAX_BLOCKS_PER_LAUNCH
        self.last_voxel_block = torch.cat(vblocks) if vblocks else torch.zeros(0, dtype=torch.int64, device=cloud.xyz.device)
        if not clouds:
            z = torch.zeros(0, 3, device=cloud.xyz.device)
            return Cloud(xyz=z, rgb=z.clone(), medial_vector=z.clone(), class_l=torch.zeros(0, 1, dtype=torch.int64, device=z.device))
        if len(clouds) == 1:
            return clouds[0]
        return Cloud(xyz=torch.cat([c.xyz for c in clouds]), rgb=torch.cat([c.rgb for c in clouds]),
                     medial_vector=torch.cat([c.medial_vector for c in clouds]), class_l=torch.cat([c.class_l for c in clouds]))

    @torch.no_grad()
    def forward_points(self, cloud: Cloud) -> Cloud:
        """Devoxelised inference (not in the reference, which keeps one labelled point per voxel and discards
        pc_voxel_id, dataset.py:214): EVERY input point gets the medial vector and class of its voxel, taken from
        the block whose inner cube contains the point.  class_l = -1 (zero vector) where there is no prediction
        (blocks of <= 20 points).  Same order and length as `cloud`."""
        if cloud.xyz.device.type != "cuda":
            cloud = cloud.to_device(self.device)
        xyz = cloud.xyz.contiguous().float()
        n, dev = xyz.shape[0], xyz.device
        ds = load_dataloader(cloud, self.voxel_size, self.block_size, self.buffer_size, self.num_workers, self.batch_size)
        medial = torch.zeros((n, 3), dtype=torch.float32, device=dev)
        cls = torch.full((n,), -1, dtype=torch.int32, device=dev)
        for bb in ds.voxelize_chunks():
            if bb.feats.shape[0] == 0:
                continue
            preds = self.model.forward(bb.feats[:, :3], bb.coords, fused_outputs=True)
            pm, pc, _ = ops.devoxelize(xyz, bb.point_index.contiguous(), bb.point_block.contiguous(), bb.pc_voxel_id.contiguous(),
                                       bb.block_centres.contiguous().float(), self.block_size,
                                       preds["medial_vector"].contiguous(), preds["class_idx"].int().contiguous())
            hit = pc >= 0
            medial = torch.where(hit.unsqueeze(1), pm, medial)
            cls = torch.where(hit, pc, cls)
        return Cloud(xyz=cloud.xyz, rgb=cloud.rgb, medial_vector=medial, class_l=cls.long().unsqueeze(1))

    @staticmethod
    def from_cfg(cfg):
        return ModelInference(model_path=cfg.model_path, weights_path=cfg.weights_path, voxel_size=cfg.voxel_size,
                              block_size=cfg.block_size, buffer_size=cfg.buffer_size, num_workers=cfg.num_workers,
                              batch_size=cfg.batch_size)
