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
        xyz = self.cloud.xyz
        dev = xyz.device
        n = xyz.shape[0]
        q = torch.div(xyz, self.block_size, rounding_mode="floor")                    # dataset.py:167-169
        ids, counts = torch.unique(q, return_counts=True, dim=0)                      # sorted rows
        ids = ids[counts > self.min_points]                                           # dataset.py:175 (quirk C-13)
        self.block_ids = ids
        self.block_centres = ids * self.block_size + (self.block_size / 2)
        nb = ids.shape[0]
        if nb == 0 or n == 0:
            self.point_index = torch.zeros(0, dtype=torch.int64, device=dev)
            self.point_block = torch.zeros(0, dtype=torch.int32, device=dev)
            return
        # dense lookup block-id triple -> block index
        lo = ids.min(0)[0]
        ext = (ids.max(0)[0] - lo + 1).long()
        lin = lambda t: ((t[:, 0] - lo[0]).long() * ext[1] + (t[:, 1] - lo[1]).long()) * ext[2] + (t[:, 2] - lo[2]).long()
        table = torch.full((int(ext.prod().item()),), -1, dtype=torch.int64, device=dev)
        table[lin(ids)] = torch.arange(nb, device=dev)
        cube = self.block_size + self.buffer_size * 2                                  # dataset.py:184
        reach = int(-(-self.buffer_size // self.block_size)) if self.buffer_size > 0 else 0
        pidx, pblk = [], []
        rng = range(-reach, reach + 1)
        ar = torch.arange(n, device=dev)
        for dx in rng:
            for dy in rng:
                for dz in rng:
                    cand = q + torch.tensor([dx, dy, dz], dtype=q.dtype, device=dev)
                    inside = ((cand >= lo) & (cand < lo + ext.to(q.dtype))).all(1)
                    b = torch.where(inside, table[lin(torch.where(inside[:, None], cand, lo.expand_as(cand)))], torch.full_like(ar, -1))
                    ok = b >= 0
                    bi = b.clamp(min=0)
                    ok &= cube_filter(xyz, self.block_centres[bi], cube)
                    pidx.append(ar[ok]); pblk.append(b[ok])
        pidx, pblk = torch.cat(pidx), torch.cat(pblk)
        order = torch.argsort(pblk * n + pidx)               # block-major, original point order inside a block
        self.point_index = pidx[order]
        self.point_block = pblk[order].int()

    def voxelize_all(self) -> BlockBatch:
        """Every block through the PointToVoxel restatement in one launch (dataset.py:192-226)."""
        xyz, rgb = self.cloud.xyz, self.cloud.rgb
        dev = xyz.device
        nb = self.block_centres.shape[0]
        if rgb is None:
            rgb = torch.zeros_like(xyz)
        pts = torch.cat((xyz, rgb), 1)[self.point_index].contiguous().float()
        if pts.shape[0] == 0:
            z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
            return BlockBatch(z(0, 6), z(0, 4, dt=torch.int32), z(0, dt=torch.bool), self.block_centres, z(0, dt=torch.int32),
                              self.point_index, self.point_block)
        if nb <= 64:          # points are block-major: per-block bounding box = min/max of a contiguous slice
            cnt = torch.bincount(self.point_block.long(), minlength=nb).cumsum(0).tolist()
            seg = [(0 if b == 0 else cnt[b - 1], cnt[b]) for b in range(nb)]
            inf = torch.full((3,), float("inf"), device=dev)
            lo = torch.stack([pts[a:b, :3].min(0)[0] if b > a else inf for a, b in seg])
            hi = torch.stack([pts[a:b, :3].max(0)[0] if b > a else -inf for a, b in seg])
        else:
            pb = self.point_block.long()
            big = torch.full((nb, 3), float("inf"), device=dev)
            lo = big.scatter_reduce(0, pb[:, None].expand(-1, 3), pts[:, :3], "amin")
            hi = (-big).scatter_reduce(0, pb[:, None].expand(-1, 3), pts[:, :3], "amax")
        vs = torch.tensor(self.voxel_size, dtype=torch.float32, device=dev)
        grid = _round_half_away((hi - lo) / vs).int().contiguous()                     # spconv calc_meta_data
        pc, rep, coords = ops.voxelize(pts, self.point_block.contiguous(), lo.contiguous(), grid, float(vs.item()))
        feats = pts[rep.long()]
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
