"""/root/reference/smart_tree/skeleton/path.py:49-140 on st_sample_tree (one device-resident
loop instead of a host loop with a GPU sync per predecessor hop)."""
from __future__ import annotations

from typing import Dict

import torch

from .. import ops
from ..data_types.branch import BranchSkeleton


def _cell_size(radii):
    r = float(radii.max().item()) if len(radii) else 0.0
    return max(r / 4.0, 1e-3)


def sample_tree(medial_pts, medial_radii, preds, distances, all_points=None, root_idx=0, visualize=False, pbar=None) -> Dict[int, BranchSkeleton]:
    """Same arguments as the reference.  medial_radii is [N,1]; preds int (-1 at the root);
    distances = tree path lengths.  Returns {branch_id: BranchSkeleton} with CPU tensors."""
    n = preds.shape[0]
    dev = preds.device
    pts = medial_pts.contiguous().float()
    rad = medial_radii.reshape(-1).contiguous().float()
    comp_off = torch.tensor([0, n], dtype=torch.int32, device=dev)
    path, blen, bpar, cnb, cnp = ops.sample_tree(pts, rad, preds.int().contiguous(), distances.contiguous().float(), comp_off, _cell_size(rad))
    nb, npth = int(cnb[0].item()), int(cnp[0].item())
    return branches_from_segment(pts, rad, path[:npth], blen[:nb], bpar[:nb])


def branches_from_segment(pts, rad, path, blen, bpar) -> Dict[int, BranchSkeleton]:
    """Gather node coordinates / radii for all branches of one component in one go."""
    pl = path.long()
    xyz = pts[pl].cpu()
    rr = rad[pl].reshape(-1, 1).cpu()
    lens = blen.cpu().tolist()
    pars = bpar.cpu().tolist()
    out, o = {}, 0
    for bid, (ln, par) in enumerate(zip(lens, pars)):
        out[bid] = BranchSkeleton(bid, int(par), xyz[o:o + ln], rr[o:o + ln])
        o += ln
    return out
