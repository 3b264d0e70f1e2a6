"""/root/reference/smart_tree/util/queries.py:107-133 on st_points_to_tubes."""
from typing import List

import torch

from .. import ops
from ..data_types.tube import Tube, collate_tubes


def pts_to_nearest_tube_gpu(pts: torch.Tensor, tubes: List[Tube], device=torch.device("cuda")):
    """Vector from each point to the nearest tube surface model, tube index, tube radius there."""
    ct = collate_tubes(tubes)
    f = lambda t: t.float().contiguous().to(device)
    pts = f(pts.reshape(-1, 3))
    m, n = ct.a.shape[0], pts.shape[0]
    off = torch.arange(0, (n + 1) * m, m, dtype=torch.int32, device=device)
    rep = lambda t: f(t).repeat(n, *([1] * (t.dim() - 1)))
    vec, idx, r = ops.points_to_tubes(pts, rep(ct.a), rep(ct.b), rep(ct.r1), rep(ct.r2), off)
    return vec, idx.long(), r
