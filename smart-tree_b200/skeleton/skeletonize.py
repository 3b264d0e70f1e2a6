"""Skeletonizer with the reference's constructor and forward
(/root/reference/smart_tree/skeleton/skeletonize.py:17-95).  All connected components are processed
in ONE batched pass: a single multi-source SSSP, a single tree-distance pass and a single
st_sample_tree launch with one CTA per component, instead of the reference's Python loop over
components with cudf/pandas round trips."""
from __future__ import annotations

import os
from typing import List

import torch

from .. import ops
from .._timing import section
from ..data_types.cloud import Cloud
from ..data_types.graph import Graph
from ..data_types.tree import DisjointTreeSkeleton, TreeSkeleton
from .filter import outlier_removal
from .graph import nn_graph
from .path import _cell_size, branches_from_segment


class Skeletonizer:
    def __init__(self, K: int, min_connection_length: float, minimum_graph_vertices: int,
                 device: torch.device = torch.device("cuda:0")):
        self.K = K
        self.min_connection_length = min_connection_length
        self.minimum_graph_vertices = minimum_graph_vertices
        self.device = device
        self.last = None       # intermediate tensors of the last call (for tests / diagnostics)

    @staticmethod
    def _emit(sub_medial, sub_radius, path, blen, bpar, cnb, cnp, comp_off, ncomp) -> List[TreeSkeleton]:
        """Two device->host copies in total: the per-component counts, then ONE packed buffer holding the path
        vertex ids, branch lengths, parent ids and the gathered node coordinates / radii of every branch of
        every component.  Nodes end up in one shared [P+B,4] host array (xyz, radius) with a spare row in
        front of each branch, pre-filled with the branch's first node: `repair` later only has to write the
        connection point there (its radius is the first node's radius by definition)."""
        import numpy as np

        from ..data_types.branch import BranchSkeleton
        from ..data_types.tree import NodeStore
        hdr = torch.cat([cnb, cnp, comp_off.int()]).cpu().numpy()
        cnb_h, cnp_h, off_h = hdr[:ncomp], hdr[ncomp:2 * ncomp], hdr[2 * ncomp:]
        nb, npth = int(cnb_h.sum()), int(cnp_h.sum())
        if nb == 0:
            return [TreeSkeleton(c, {}) for c in range(ncomp)]
        dev = path.device
        pseg = torch.cat([path[off_h[c]:off_h[c] + cnp_h[c]] for c in range(ncomp)]) if ncomp > 1 else path[:npth]
        lseg = torch.cat([blen[off_h[c]:off_h[c] + cnb_h[c]] for c in range(ncomp)]) if ncomp > 1 else blen[:nb]
        qseg = torch.cat([bpar[off_h[c]:off_h[c] + cnb_h[c]] for c in range(ncomp)]) if ncomp > 1 else bpar[:nb]
        gidx = pseg.long()
        if ncomp > 1:
            gidx = gidx + torch.from_numpy(np.repeat(off_h[:ncomp].astype(np.int64), cnp_h)).to(dev)
        payload = torch.cat([lseg, qseg, sub_medial[gidx].reshape(-1).view(torch.int32), sub_radius[gidx].view(torch.int32)]).cpu().numpy()
        lens, pars = payload[:nb].astype(np.int64), payload[nb:2 * nb]
        xyz = payload[2 * nb:2 * nb + 3 * npth].view(np.float32).reshape(npth, 3)
        rad = payload[2 * nb + 3 * npth:].view(np.float32)
        # host node array with one spare row per branch (copy of the branch's first node)
        start = np.cumsum(lens) - lens
        rep = np.ones(npth, np.int64)
        rep[start] = 2
        src = np.repeat(np.arange(npth), rep)
        nodes = torch.from_numpy(np.concatenate([xyz[src], rad[src, None]], 1))
        store = NodeStore(nodes, None)
        row = (start + np.arange(nb)).tolist()        # spare row of every branch
        lens_l, pars_l = lens.tolist(), pars.tolist()
        # all per-branch views from two split calls (per-branch slicing costs ~10 us of Python each)
        sizes = [v for ln in lens_l for v in (1, ln)]
        xyz_views = nodes[:, :3].split(sizes)[1::2]
        rad_views = nodes[:, 3:4].split(sizes)[1::2]
        skeletons, bi = [], 0
        for c in range(ncomp):
            branches = {}
            for bid in range(int(cnb_h[c])):
                branches[bid] = BranchSkeleton(bid, pars_l[bi], xyz_views[bi], rad_views[bi], _flat=(store, row[bi], lens_l[bi], False))
                bi += 1
            skeletons.append(TreeSkeleton(c, branches))
        return skeletons

    def forward(self, cloud: Cloud) -> DisjointTreeSkeleton:
        cloud = cloud.to_device(self.device)
        if len(cloud) == 0:
            return DisjointTreeSkeleton([])
        # skeletonize.py:34-35
        with section("skel.outlier"):
            keep = outlier_removal(cloud.medial_pts, cloud.radius.unsqueeze(1), nb_points=8)
            cloud = cloud.filter(keep)
        n = len(cloud)
        if n == 0:
            return DisjointTreeSkeleton([])
        with section("skel.medial"):
            medial = cloud.medial_pts.contiguous()
            radius = cloud.radius.contiguous()
        # skeletonize.py:37-41 (the clamp applies to graph building only, quirk C-21)
        with section("skel.nn_graph"):
            graph: Graph = nn_graph(medial, radius.clamp(min=self.min_connection_length), K=self.K)
        # skeletonize.py:43-45
        with section("skel.components"):
            label, roots, sizes = graph.ranked_components(self.minimum_graph_vertices)
        ncomp = int(roots.shape[0])
        if ncomp == 0:
            self.last = dict(keep=keep, n_components=0)
            return DisjointTreeSkeleton([])
        with section("skel.regroup_csr"):
            dev = medial.device
            # rank of each vertex's component (-1 = dropped), then vertices grouped by rank, ascending id inside
            rank_of_root = torch.full((n,), -1, dtype=torch.int64, device=dev)
            rank_of_root[roots] = torch.arange(ncomp, device=dev)
            vrank = rank_of_root[label.long()]
            sel = torch.nonzero(vrank >= 0).flatten()
            order = sel[torch.argsort(vrank[sel], stable=True)]               # new id -> old vertex id
            m = int(order.shape[0])
            new_id = torch.full((n,), -1, dtype=torch.int32, device=dev)
            new_id[order] = torch.arange(m, dtype=torch.int32, device=dev)
            comp_off = torch.zeros(ncomp + 1, dtype=torch.int64, device=dev)
            comp_off[1:] = torch.cumsum(sizes, 0)
            comp_of = torch.repeat_interleave(torch.arange(ncomp, device=dev), sizes)
            # induced edges, renumbered (skeletonize.py:60-71): done inside the CSR build through new_id
            row_ptr, col, w = ops.csr_build(graph.edges, graph.edge_weights, m, vertex_map=new_id)
            sub_xyz = cloud.xyz[order]
            sub_medial = medial[order].contiguous()
            sub_radius = radius[order].contiguous()
            # root of each component = first argmin of surface y (cloud.py:205-206)
            y = sub_xyz[:, 1].contiguous()
            if ncomp <= 64:       # components are contiguous segments: a plain argmin per segment
                off_l = comp_off.tolist()
                src = torch.stack([torch.argmin(y[off_l[c]:off_l[c + 1]]) + off_l[c] for c in range(ncomp)])
            else:
                miny = torch.full((ncomp,), float("inf"), device=dev).scatter_reduce(0, comp_of, y, "amin")
                cand = torch.where(y == miny[comp_of], torch.arange(m, device=dev), torch.full((m,), m, device=dev))
                src = torch.full((ncomp,), m, dtype=torch.int64, device=dev).scatter_reduce(0, comp_of, cand, "amin")
        # skeletonize.py:73-78
        with section("skel.sssp"):
            # threshold step of the distance-ordered SSSP schedule (any value is exact; ~1/16 of a tree height measured best)
            delta = float(os.environ.get("ST_SSSP_DELTA", 25 * self.min_connection_length))
            dist, pred = ops.sssp(row_ptr, col, w, m, src.int().contiguous(), delta=delta)
        with section("skel.tree_dist"):
            is_root = torch.zeros(m, dtype=torch.uint8, device=dev)
            is_root[src] = 1
            # skeletonize.py:80-85
            tdist = ops.tree_distances(sub_medial, pred, is_root)
            off32 = comp_off.int().contiguous()
            pred_local = torch.where(pred >= 0, pred - off32[comp_of], pred).contiguous()
        # skeletonize.py:87-93
        with section("skel.sample_tree"):
            path, blen, bpar, cnb, cnp = ops.sample_tree(sub_medial, sub_radius, pred_local, tdist, off32, _cell_size(sub_radius))
        with section("skel.emit"):
            skeletons = self._emit(sub_medial, sub_radius, path, blen, bpar, cnb, cnp, comp_off, ncomp)
        self.last = dict(keep=keep, order=order, comp_off=comp_off, pred=pred_local, dist=dist, tree_dist=tdist, roots=src,
                         path=path, branch_len=blen, branch_parent=bpar, comp_n_branches=cnb, comp_n_path=cnp,
                         n_components=ncomp, edges=graph.edges, edge_weights=graph.edge_weights)
        return DisjointTreeSkeleton(skeletons)
