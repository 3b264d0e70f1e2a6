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
    def _emit(sub_medial, sub_radius, path, blen, bpar, cnb, cnp, off32, ncomp, post=None) -> List[TreeSkeleton]:
        """Branch assembly and (optionally) prune / repair / smooth on the device in one launch
        (st_finish_skeletons), then two device->host copies in total: a 4-word header + the per-component
        branch counts, and ONE packed buffer with the branch table, the node array [R,4] (xyz, radius; one
        spare row in front of each branch for the repair connection point) and the smoothed radii.
        `post` = dict(prune=(min_radius, min_length) | None, repair=bool, smooth=kernel_size) or None."""
        import numpy as np

        from ..data_types.branch import BranchSkeleton, PackedBranchSkeleton
        from ..data_types.tree import NodeStore
        post = post or {}
        prune = post.get("prune")
        k = int(post.get("smooth") or 0)
        out = ops.finish_skeletons(sub_medial, sub_radius, off32, path, blen, bpar, cnb, cnp, prune_first=prune is not None,
                                   min_radius=prune[0] if prune else 0.0, min_length=prune[1] if prune else 0.0,
                                   repair=bool(post.get("repair")), smooth_kernel=k)
        hdr = torch.cat([out[:4], cnb]).cpu().numpy()
        nb, nrow = int(hdr[0]), int(hdr[1])
        cnb_h = hdr[4:4 + ncomp]
        if nb == 0:
            return [TreeSkeleton(c, {}) for c in range(ncomp)]
        payload = out[4:4 + 4 * nb + 5 * nrow].cpu().numpy()
        bmeta = payload[:4 * nb].reshape(nb, 4)
        nodes = torch.from_numpy(payload[4 * nb:4 * nb + 4 * nrow].view(np.float32).reshape(nrow, 4))
        smooth = torch.from_numpy(payload[4 * nb + 4 * nrow:].view(np.float32))
        store = NodeStore(nodes, None)
        flags = bmeta[:, 3]
        kept = np.flatnonzero(flags & 1)
        if len(kept) == 0:
            return [TreeSkeleton(c, {}) for c in range(ncomp)]
        conn = (flags[kept] & 2) != 0
        first = bmeta[kept, 0] + np.where(conn, 0, 1)
        cnt = bmeta[kept, 0] + bmeta[kept, 1] + 1 - first
        comp_of = np.repeat(np.arange(ncomp), cnb_h)
        local = np.arange(nb) - np.repeat(np.cumsum(cnb_h) - cnb_h, cnb_h)
        rows_l, lens_l, pars_l, flags_l = bmeta[kept, 0].tolist(), bmeta[kept, 1].tolist(), bmeta[kept, 2].tolist(), flags[kept].tolist()
        comp_l, bid_l = comp_of[kept].tolist(), local[kept].tolist()
        per_comp = [dict() for _ in range(ncomp)]
        if post:
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
            return [TreeSkeleton(c, per_comp[c]) for c in range(ncomp)]
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
        return [TreeSkeleton(c, per_comp[c]) for c in range(ncomp)]

    def forward(self, cloud: Cloud, post: dict = None, shard=None) -> DisjointTreeSkeleton:
        """`post` (optional, not in the reference signature): post-processing to fuse into the device-side branch
        assembly -- see _emit.  The returned skeleton records it in `.post_applied` so that Pipeline.post_process
        does not repeat it."""
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
        self.component_ids = list(range(int(roots.shape[0])))
        if shard is not None:
            # multi-GPU plots: every rank builds the same graph and components (cheap, redundant) and then
            # extracts the skeletons of components rank, rank+world, ... of the size-ordered list only
            rank, world = shard
            self.component_ids = self.component_ids[rank::world]
            roots, sizes = roots[rank::world], sizes[rank::world]
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
        if shard is not None and post and (not self.component_ids or self.component_ids[0] != 0):
            post = {k: v for k, v in post.items() if k != "prune"}        # only the globally first skeleton is pruned (quirk C-18)
        with section("skel.emit"):
            skeletons = self._emit(sub_medial, sub_radius, path, blen, bpar, cnb, cnp, off32, ncomp, post)
            if shard is not None:
                for sk_, gid in zip(skeletons, self.component_ids):
                    sk_._id = gid                                            # global (size-ordered) component index
        self.last = dict(keep=keep, order=order, comp_off=comp_off, pred=pred_local, dist=dist, tree_dist=tdist, roots=src,
                         path=path, branch_len=blen, branch_parent=bpar, comp_n_branches=cnb, comp_n_path=cnp,
                         n_components=ncomp, edges=graph.edges, edge_weights=graph.edge_weights)
        return DisjointTreeSkeleton(skeletons, post_applied=dict(post) if post else None)
