"""Block tiling + voxelisation for inference
(/root/reference/smart_tree/dataset/dataset.py:144-242).  The reference loops over blocks on the
host (one O(N) mask, one device->host copy and one CPU voxeliser call per block); here every block
is formed, ranged and voxelised in one batched pass on the device (st_voxelize)."""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .. import ops
from ..data_types.cloud import Cloud
from ..util.maths import cube_filter


@dataclass
class BlockBatch:
    feats: torch.Tensor          # [M,6] xyz,rgb of each voxel's first point
    coords: torch.Tensor         # [M,4] int32 (block, z, y, x)
    mask: torch.Tensor           # [M] bool: representative point inside the inner block cube
    block_centres: torch.Tensor  # [B,3]
    pc_voxel_id: torch.Tensor    # [T] voxel of every (block, point) pair, -1 = dropped by the voxeliser
    point_index: torch.Tensor    # [T] index of that pair's point in the input cloud
    point_block: torch.Tensor    # [T]


def _round_half_away(q):
    t = torch.trunc(q)
    return t + ((q - t) >= 0.5).to(q.dtype)


class SingleTreeInference:
    def __init__(self, cloud: Cloud, voxel_size: float, block_size: float = 4, buffer_size: float = 0.4, min_points=20,
                 file_name=None, device=torch.device("cuda:0")):
        self.cloud = cloud
        self.voxel_size = voxel_size
        self.block_size = block_size
        self.buffer_size = buffer_size
        self.min_points = min_points
        self.device = device
        self.file_name = file_name
        self.compute_blocks()

    def compute_blocks(self):
        """dataset.py:166-190 in three kernel launches (st_block_list / _count / _emit)."""
        xyz = self.cloud.xyz.contiguous().float()
        ids, pidx, pblk, lo, hi = ops.block_tiling(xyz, self.block_size, self.buffer_size, self.min_points)
        self.block_ids = ids
        self.block_centres = ids * self.block_size + (self.block_size / 2)
        self.point_index, self.point_block = pidx, pblk
        self.block_lo, self.block_hi = lo, hi

    def keep_shard(self, rank: int, world: int):
        """Multi-GPU plots: keep blocks rank, rank+world, ... of the (deterministic) block list.
        Blocks are independent forward passes (eval-mode BatchNorm), so every rank labels its own blocks and the
        labelled voxels are exchanged afterwards (dist.gather_labelled).  `block_global` = index in the full list."""
        nb = int(self.block_centres.shape[0])
        dev = self.point_block.device
        self.block_global = torch.arange(rank, nb, world, device=dev)
        if world == 1:
            return self
        sel = (self.point_block % world) == rank
        self.point_index = self.point_index[sel].contiguous()
        self.point_block = torch.div(self.point_block[sel], world, rounding_mode="floor").to(self.point_block.dtype).contiguous()
        self.block_ids, self.block_centres = self.block_ids[rank::world], self.block_centres[rank::world]
        self.block_lo, self.block_hi = self.block_lo[rank::world].contiguous(), self.block_hi[rank::world].contiguous()
        return self

    MAX_BLOCKS_PER_LAUNCH = 16384      # the batch index has 15 bits in the packed voxel key

    def voxelize_chunks(self):
        """Yields one BlockBatch per <= 16384 blocks (block indices in `coords` are local to the chunk).  A tree
        is a single chunk; plots tiled into tens of thousands of 64^3 blocks are cut into several forwards, which
        changes nothing numerically (eval-mode BatchNorm: blocks do not interact)."""
        nb = self.block_centres.shape[0]
        if nb <= self.MAX_BLOCKS_PER_LAUNCH:
            yield self._voxelize()
            return
        dev = self.cloud.xyz.device
        for b0 in range(0, nb, self.MAX_BLOCKS_PER_LAUNCH):
            b1 = min(b0 + self.MAX_BLOCKS_PER_LAUNCH, nb)
            lo_hi = torch.searchsorted(self.point_block, torch.tensor([b0, b1], dtype=self.point_block.dtype, device=dev)).tolist()
            sub = SingleTreeInference.__new__(SingleTreeInference)
            sub.__dict__.update(self.__dict__)
            sub.point_index = self.point_index[lo_hi[0]:lo_hi[1]]
            sub.point_block = (self.point_block[lo_hi[0]:lo_hi[1]] - b0).contiguous()
            sub.block_centres, sub.block_lo, sub.block_hi = self.block_centres[b0:b1], self.block_lo[b0:b1], self.block_hi[b0:b1]
            yield sub._voxelize()

    def voxelize_all(self) -> BlockBatch:
        """Every block through the PointToVoxel restatement in one launch (dataset.py:192-226)."""
        assert self.block_centres.shape[0] <= self.MAX_BLOCKS_PER_LAUNCH, "use voxelize_chunks() for this many blocks"
        return self._voxelize()

    def _voxelize(self) -> BlockBatch:
        xyz, rgb = self.cloud.xyz, self.cloud.rgb
        dev = xyz.device
        if rgb is None:
            rgb = torch.zeros_like(xyz)
        pts = torch.cat((xyz, rgb), 1)[self.point_index].contiguous().float()
        if pts.shape[0] == 0:
            z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
            return BlockBatch(z(0, 6), z(0, 4, dt=torch.int32), z(0, dt=torch.bool), self.block_centres, z(0, dt=torch.int32),
                              self.point_index, self.point_block)
        lo, hi = self.block_lo, self.block_hi                                         # each block cloud's own bounding box
        vs = torch.tensor(self.voxel_size, dtype=torch.float32, device=dev)
        grid = _round_half_away((hi - lo) / vs).int().contiguous()                     # spconv calc_meta_data
        pc, rep, coords = ops.voxelize(pts, self.point_block.contiguous(), lo.contiguous(), grid, float(self.voxel_size))
        feats = ops.gather_rows(pts, rep)
        mask = cube_filter(feats[:, :3], self.block_centres[coords[:, 0].long()], self.block_size)   # dataset.py:224
        return BlockBatch(feats, coords, mask, self.block_centres, pc, self.point_index, self.point_block)

    def __len__(self):
        return int(self.block_centres.shape[0])

    def __getitem__(self, idx):
        """Per-block view with the reference's return signature: feats [M,6], coords [M,4] (batch
        column 0), mask [M], file name."""
        bb = getattr(self, "_all", None) or self.voxelize_all()
        self._all = bb
        sel = bb.coords[:, 0] == idx
        coords = bb.coords[sel].clone()
        coords[:, 0] = 0
        return bb.feats[sel], coords, bb.mask[sel], self.file_name


def load_dataloader(cloud: Cloud, voxel_size: float, block_size: float, buffer_size: float, num_workers: float, batch_size: float):
    """Reference signature (dataset.py:232-242).  Returns the block dataset itself: batching and
    worker processes are not needed when all blocks are voxelised in one device pass."""
    return SingleTreeInference(cloud, voxel_size, block_size, buffer_size)
