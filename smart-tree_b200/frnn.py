"""`frnn` stand-in: frnn_grid_points with the call signature used at
/root/reference/smart_tree/skeleton/graph.py:15-24 (batch size 1), on st_knn."""
import torch

from . import ops


def frnn_grid_points(points1, points2, lengths1=None, lengths2=None, K=-1, r=-1.0, grid=None, return_nn=False,
                     return_sorted=True, radius_cell_ratio=2.0):
    if points1.dim() != 3 or points1.shape[0] != 1 or points2.shape[0] != 1:
        raise NotImplementedError("batch size 1 only (all the reference uses)")
    if not return_sorted:
        raise NotImplementedError("return_sorted=False")
    r = float(r.item() if torch.is_tensor(r) else r)
    idx, d2 = ops.knn(points1[0].contiguous().float(), points2[0].contiguous().float(), int(K), r)
    nn_pts = points2[0][idx.clamp(min=0).long()].unsqueeze(0) if return_nn else None
    return d2.unsqueeze(0), idx.long().unsqueeze(0), nn_pts, None
