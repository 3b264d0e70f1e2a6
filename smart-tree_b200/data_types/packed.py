"""Packed skeleton results: what st_finish_skeletons writes (include/st_b200.h) kept as ONE host array per
skeletoniser call -- and as the device buffer the multi-GPU gather ships -- with the reference's object model
(/root/reference/smart_tree/data_types/tree.py, branch.py) materialised from it on first access.  Building a few hundred
to a few thousand Python objects per tree costs about a millisecond of host time; results that are only written to
disk, gathered over NCCL or digested never pay it."""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from .branch import BranchSkeleton, PackedBranchSkeleton
from .tree import NodeStore, TreeSkeleton


def unpack(nb: int, nrow: int, cnb_h, comp_ids, payload, post_done: bool) -> List[TreeSkeleton]:
    """payload = int32 words bmeta[B][4] | nodes[R][4] (float) | smooth[R] (float) -> one TreeSkeleton per component."""
    ncomp = len(comp_ids)
    if nb == 0:
        return [TreeSkeleton(c, {}) for c in comp_ids]
    bmeta = payload[:4 * nb].reshape(nb, 4)
    nodes = torch.from_numpy(payload[4 * nb:4 * nb + 4 * nrow].view(np.float32).reshape(nrow, 4))
    smooth = torch.from_numpy(payload[4 * nb + 4 * nrow:4 * nb + 5 * nrow].view(np.float32))
    store = NodeStore(nodes, None)
    flags = bmeta[:, 3]
    kept = np.flatnonzero(flags & 1)
    if len(kept) == 0:
        return [TreeSkeleton(c, {}) for c in comp_ids]
    conn = (flags[kept] & 2) != 0
    first = bmeta[kept, 0] + np.where(conn, 0, 1)
    cnt = bmeta[kept, 0] + bmeta[kept, 1] + 1 - first
    comp_of = np.repeat(np.arange(ncomp), cnb_h)
    local = np.arange(nb) - np.repeat(np.cumsum(cnb_h) - cnb_h, cnb_h)
    rows_l, lens_l, pars_l, flags_l = bmeta[kept, 0].tolist(), bmeta[kept, 1].tolist(), bmeta[kept, 2].tolist(), flags[kept].tolist()
    comp_l, bid_l = comp_of[kept].tolist(), local[kept].tolist()
    per_comp = [dict() for _ in range(ncomp)]
    if post_done:
        # post-processing is complete: pack the surviving rows (one gather) so that the per-branch views come
        # from two gap-free split calls (a view costs ~0.5 us; gaps would double their number)
        csum = np.cumsum(cnt)
        rows = np.repeat(first - (csum - cnt), cnt) + np.arange(int(csum[-1]))
        sm_row = np.repeat((flags[kept] & 4) != 0, cnt)
        packed = nodes[torch.from_numpy(rows)]
        radcol = torch.where(torch.from_numpy(sm_row), smooth[torch.from_numpy(rows)], packed[:, 3])
        xyz_all = packed[:, :3]
        starts = (csum - cnt).tolist()
        sizes = cnt.tolist()
        for i in range(len(rows_l)):
            # per-branch views are cut on first access (PackedBranchSkeleton)
            per_comp[comp_l[i]][bid_l[i]] = PackedBranchSkeleton(bid_l[i], pars_l[i], xyz_all, radcol, starts[i], sizes[i],
                                                                  radii_1d=bool(flags_l[i] & 4))
        return [TreeSkeleton(comp_ids[c], per_comp[c]) for c in range(ncomp)]
    # plain assembly: keep the spare rows (object-level repair writes the connection points there)
    gaps = np.empty(2 * len(kept) + 1, np.int64)
    gaps[0:-1:2] = first - np.concatenate([[0], (first + cnt)[:-1]])
    gaps[1::2] = cnt
    gaps[-1] = nrow - (first[-1] + cnt[-1])
    sizes = gaps.tolist()
    xyz_views = nodes[:, :3].split(sizes)[1::2]
    rad_views = nodes[:, 3:4].split(sizes)[1::2]
    for i in range(len(rows_l)):
        per_comp[comp_l[i]][bid_l[i]] = BranchSkeleton(bid_l[i], pars_l[i], xyz_views[i], rad_views[i],
                                                       _flat=(store, rows_l[i], lens_l[i], False))
    return [TreeSkeleton(comp_ids[c], per_comp[c]) for c in range(ncomp)]


class PackedSkeletons:
    """List-like view of one skeletoniser call's result; `DisjointTreeSkeleton.skeletons` holds one of these."""

    def __init__(self, nb, nrow, cnb_h, comp_ids, payload, dev_payload, post_done):
        self.nb, self.nrow, self.cnb_h, self.comp_ids = int(nb), int(nrow), np.asarray(cnb_h, np.int32), list(comp_ids)
        self.payload, self.dev_payload, self.post_done = payload, dev_payload, bool(post_done)
        self._objs = None

    # ---- packed access (no objects)
    @property
    def n_branches_total(self):
        return self.nb

    def wire(self, unit: int = 0):
        """Device int32 vector: [n words, B, R, n components, unit, post_done, 0, 0 | branches per component | component ids |
        bmeta | nodes | smooth] -- what dist.gather_packed all-gathers."""
        dev = self.dev_payload.device
        nc = len(self.comp_ids)
        head = np.zeros(8 + 2 * nc, np.int32)
        head[1:6] = (self.nb, self.nrow, nc, unit, int(self.post_done))
        head[8:8 + nc] = self.cnb_h[:nc]
        head[8 + nc:] = self.comp_ids
        head[0] = len(head) + int(self.dev_payload.numel())
        return torch.cat([torch.from_numpy(head).to(dev, non_blocking=True), self.dev_payload])

    @staticmethod
    def from_wire(words: np.ndarray):
        """Inverse of wire() on a host int32 array -> (unit, PackedSkeletons)."""
        n, nb, nrow, nc, unit, post_done = (int(v) for v in words[:6])
        cnb = words[8:8 + nc].copy()
        ids = words[8 + nc:8 + 2 * nc].tolist()
        payload = words[8 + 2 * nc:n].copy()
        return unit, PackedSkeletons(nb, nrow, cnb, ids, payload if nb else None, None, bool(post_done))

    # ---- the reference's object model, on demand
    def _materialise(self):
        if self._objs is None:
            self._objs = unpack(self.nb, self.nrow, self.cnb_h, self.comp_ids, self.payload, self.post_done)
        return self._objs

    def __len__(self):
        return len(self.comp_ids)

    def __iter__(self):
        return iter(self._materialise())

    def __getitem__(self, i):
        return self._materialise()[i]

    def __bool__(self):
        return len(self.comp_ids) > 0

    def __repr__(self):
        return f"PackedSkeletons({len(self.comp_ids)} skeletons, {self.nb} branches, {self.nrow} node rows)"
