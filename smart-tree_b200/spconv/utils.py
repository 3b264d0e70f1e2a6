"""spconv.pytorch.utils.PointToVoxel stand-in (max_num_points_per_voxel == 1 only), call site
/root/reference/smart_tree/dataset/dataset.py:199-216.  Deterministic on the device: the first
point in input order represents a voxel, voxels are numbered by first appearance -- the CPU
semantics of spconv, not its hash-order GPU variant."""
import torch

from .. import ops


class PointToVoxel:
    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_voxels, max_num_points_per_voxel, device=torch.device("cuda")):
        if max_num_points_per_voxel != 1:
            raise NotImplementedError("libst_b200 voxeliser keeps one point per voxel")
        if len(set(float(v) for v in vsize_xyz)) != 1:
            raise NotImplementedError("isotropic voxels only")
        self.vsize = float(vsize_xyz[0])
        self.range = [float(v) for v in coors_range_xyz]
        self.num_point_features = num_point_features
        self.max_num_voxels = max_num_voxels
        self.device = device

    def generate_voxel_with_id(self, pc: torch.Tensor):
        if not pc.is_cuda:
            pc = pc.to(self.device if torch.device(self.device).type == "cuda" else "cuda")
        pc = pc.contiguous().float()
        dev = pc.device
        lo = torch.tensor([self.range[:3]], dtype=torch.float32, device=dev)
        hi = torch.tensor([self.range[3:]], dtype=torch.float32, device=dev)
        q = (hi - lo) / torch.tensor(self.vsize, dtype=torch.float32, device=dev)
        t = torch.trunc(q)
        grid = (t + ((q - t) >= 0.5).float()).int().contiguous()
        pcid, rep, coords = ops.voxelize(pc, None, lo.contiguous(), grid, self.vsize)
        m = min(int(rep.shape[0]), int(self.max_num_voxels))
        voxels = pc[rep[:m].long()].unsqueeze(1)
        if m < rep.shape[0]:
            pcid = torch.where(pcid >= m, torch.full_like(pcid, -1), pcid)
        return voxels, coords[:m, 1:].contiguous(), torch.ones(m, dtype=torch.int32, device=dev), pcid.long()

    def __call__(self, pc):
        v, c, n, _ = self.generate_voxel_with_id(pc)
        return v, c, n
