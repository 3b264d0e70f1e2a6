"""kNN / graph construction with the reference's signatures
(/root/reference/smart_tree/skeleton/graph.py:12-60), running on st_knn / st_edges_from_knn."""
from __future__ import annotations

import torch

from .. import ops
from ..data_types.graph import Graph


def knn(src: torch.Tensor, dest: torch.Tensor, K=50, r=1.0, grid=None):
    """Returns idxs [N,K] int64 (-1 padded), dists [N,K] = sqrt(d2) (NaN where padded, as the
    reference's `dists.sqrt()` of FRNN's -1 padding), grid (always None: the grid is rebuilt)."""
    idx, d2 = ops.knn(src.contiguous().float(), dest.contiguous().float(), int(K), float(r))
    return idx.long(), d2.sqrt(), None


def nn(src, dest, r=1.0, grid=None):
    idx, dist, grid = knn(src, dest, K=1, r=r, grid=grid)
    return idx.squeeze(1), dist.squeeze(1), grid


def make_edges(dists, idxs):
    n, K = dists.shape
    parent = torch.arange(n, device=dists.device).unsqueeze(1).expand(n, K)
    valid = idxs.reshape(-1) > 0                              # sic (graph.py:59)
    return torch.stack([parent, idxs], dim=2).reshape(-1, 2)[valid], dists.reshape(-1)[valid]


def nn_graph(points: torch.Tensor, radii: torch.Tensor, K=40) -> Graph:
    """graph.py:36-40.  The per-point radius cut-off is applied inside the kNN kernel
    (query_radius), which yields exactly the reference's `idxs[dists > radii] = -1`."""
    points = points.contiguous().float()
    radii = radii.contiguous().float()
    r = float(radii.max().item()) if len(radii) else 0.0
    idx, d2 = ops.knn(points, points, int(K), r, query_radius=radii)
    edges, w = ops.edges_from_knn(idx, d2)
    return Graph(points, edges, w)
