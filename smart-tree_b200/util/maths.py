"""/root/reference/smart_tree/util/maths.py:135-155 -- half-open axis-aligned cube test."""
import torch


def torch_bb_filter(points, min_x, max_x, min_y, max_y, min_z, max_z):
    return ((points[:, 0] >= min_x) & (points[:, 0] < max_x) & (points[:, 1] >= min_y) & (points[:, 1] < max_y)
            & (points[:, 2] >= min_z) & (points[:, 2] < max_z))


def cube_filter(points: torch.Tensor, center: torch.Tensor, cube_size):
    """center may be [3] (one cube) or [N,3] (one cube per point).  Bounds are computed in the
    dtype of `center` exactly like the reference: lo = center - cube_size/2, hi = center + cube_size/2."""
    center = center.to(points.device)
    lo = center - (cube_size / 2)
    hi = center + (cube_size / 2)
    if center.dim() == 1:
        return torch_bb_filter(points, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])
    return ((points >= lo) & (points < hi)).all(dim=1)
