"""/root/reference/smart_tree/skeleton/filter.py:6-11 on st_outlier_mask."""
import torch

from .. import ops


def outlier_removal(points: torch.Tensor, radii: torch.Tensor, nb_points=4):
    """radii is [N,1] as in the reference call (skeletonize.py:34).  Keeps a point iff its
    nb_points nearest neighbours (self included) all lie strictly inside its own radius."""
    radii = radii.reshape(-1).contiguous().float()
    r = float(radii.max().item()) if len(radii) else 0.0
    return ops.outlier_mask(points.contiguous().float(), radii, r, int(nb_points))
