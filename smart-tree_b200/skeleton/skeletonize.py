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
        """One gather + one device->host copy for the node coordinates / radii of every branch of
        every component.  Nodes live in one shared [P+B,4] array (xyz, radius) with a spare row in
        front of each branch, pre-filled with the branch's first node: `repair` later only has to
        write the connection point there (its radius is the first node's radius by definition)."""
        from ..data_types.branch import BranchSkeleton
        from ..data_types.tree import NodeStore
        counts = torch.stack([cnb, cnp]).cpu()
        cnb_h, cnp_h, off_h = counts[0].tolist(), counts[1].tolist(), comp_off.cpu().tolist()
        dev = path.device
        if sum(cnb_h) == 0:
            return [TreeSkeleton(c, {}) for c in range(ncomp)]
        seg = torch.cat([torch.arange(off_h[c], off_h[c] + cnp_h[c], device=dev) for c in range(ncomp)])
        bseg = torch.cat([torch.arange(off_h[c], off_h[c] + cnb_h[c], device=dev) for c in range(ncomp)])
        base = torch.repeat_interleave(comp_off[:-1], torch.tensor(cnp_h, device=dev))
        gidx = path[seg].long() + base
        lens_d = blen[bseg].long()
        starts = torch.cumsum(lens_d, 0) - lens_d
        rep = torch.ones_like(gidx)
        rep[starts] = 2                                          # first node of every branch twice
        gidx = torch.repeat_interleave(gidx, rep)
        nodes_dev = torch.cat([sub_medial[gidx], sub_radius[gidx].unsqueeze(1)], 1).contiguous()
        nodes = nodes_dev.cpu()
        lens = lens_d.cpu().tolist()
        pars = bpar[bseg].cpu().tolist()
        store = NodeStore(nodes, nodes_dev)
        skeletons, o, bi = [], 0, 0
        for c in range(ncomp):
            branches = {}
            for bid in range(cnb_h[c]):
                ln = lens[bi]
                branches[bid] = BranchSkeleton(bid, int(pars[bi]), nodes[o + 1:o + 1 + ln, :3], nodes[o + 1:o + 1 + ln, 3:4],
                                               _flat=(store, o, ln, False))
                o += ln + 1
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
            # induced edges, renumbered (skeletonize.py:60-71)
            e = graph.edges.long()
            esel = new_id[e[:, 0]] >= 0
            sub_edges = new_id[e[esel]].contiguous()
            sub_w = graph.edge_weights[esel].contiguous()
            row_ptr, col, w = ops.csr_build(sub_edges, sub_w, m)
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
