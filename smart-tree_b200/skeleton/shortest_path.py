"""/root/reference/smart_tree/skeleton/shortest_path.py:12-21,46-74 on st_csr_build / st_sssp /
st_tree_distances."""
from __future__ import annotations

import torch

from .. import ops


def shortest_paths(root: int, edges: torch.Tensor, edge_weights: torch.Tensor, renumber=True, n=None):
    """Undirected weighted SSSP from `root`.  Returns (vertex ids, predecessor (-1 at the root),
    distance) in vertex-id order, as the reference assumes of cugraph.sssp(renumber=False)."""
    n = int(edges.max().item()) + 1 if n is None else n
    row_ptr, col, w = ops.csr_build(edges.int().contiguous(), edge_weights.float().contiguous(), n)
    src = torch.tensor([root], dtype=torch.int32, device=edges.device)
    dist, pred = ops.sssp(row_ptr, col, w, n, src)
    return torch.arange(n, device=edges.device), pred.long(), dist


def tree_path_lengths(points: torch.Tensor, preds: torch.Tensor, root: int):
    """pred_graph + second sssp (shortest_path.py:46-55, skeletonize.py:80-85): length of the
    predecessor-tree path from the root, accumulated root -> leaf in fp32."""
    is_root = torch.zeros(preds.shape[0], dtype=torch.uint8, device=preds.device)
    is_root[root] = 1
    return ops.tree_distances(points.contiguous().float(), preds.int().contiguous(), is_root)
