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
        # vertices of the SSSP graph renumbered in Z-order of their medial points (st_sssp orig_id): what the CTA-local kernel
        # (ST_SSSP_LOCAL=1) needs; the default kernel relaxes the woken vertices of a warp's 32 consecutive vertices one after
        # the other, and neighbours in space wake together -- measured 10.8 ms in spatial order against 4.4 ms in the caller's
        self.spatial_sssp = bool(int(os.environ.get("ST_SSSP_SPATIAL", os.environ.get("ST_SSSP_LOCAL", "0"))))

    @staticmethod
    def _emit(sub_medial, sub_radius, path, blen, bpar, cnb, cnp, off32, ncomp, post=None, comp_ids=None):
        """Branch assembly and (optionally) prune / repair / smooth on the device in one launch
        (st_finish_skeletons), then two device->host copies in total: a 4-word header + the per-component
        branch counts, and ONE packed buffer with the branch table, the node array [R,4] (xyz, radius; one
        spare row in front of each branch for the repair connection point) and the smoothed radii.
        `post` = dict(prune=(min_radius, min_length) | None, repair=bool, smooth=kernel_size) or None.
        Returns a PackedSkeletons: the result as packed host arrays (+ the device buffer for the multi-GPU gather);
        the TreeSkeleton / BranchSkeleton objects are built from it on first access."""
        from ..data_types.packed import PackedSkeletons
        post = post or {}
        prune = post.get("prune")
        k = int(post.get("smooth") or 0)
        out = ops.finish_skeletons(sub_medial, sub_radius, off32, path, blen, bpar, cnb, cnp, prune_first=prune is not None,
                                   min_radius=prune[0] if prune else 0.0, min_length=prune[1] if prune else 0.0,
                                   repair=bool(post.get("repair")), smooth_kernel=k)
        hdr = torch.cat([out[:4], cnb]).cpu().numpy()
        nb, nrow = int(hdr[0]), int(hdr[1])
        cnb_h = hdr[4:4 + ncomp].copy()
        dev_payload = out[4:4 + 4 * nb + 5 * nrow]
        payload = dev_payload.cpu().numpy() if nb else None
        ids = list(range(ncomp)) if comp_ids is None else [int(c) for c in comp_ids]
        return PackedSkeletons(nb, nrow, cnb_h, ids, payload, dev_payload, bool(post))

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
            if ncomp == 1:                                                    # a single tree: one component, ids as they are
                order = torch.nonzero(label == roots[0]).flatten()
                m = int(order.shape[0])
                comp_off = torch.tensor([0, m], dtype=torch.int64, device=dev)
                comp_of = torch.zeros(m, dtype=torch.int64, device=dev)
            else:
                rank_of_root = torch.full((n,), -1, dtype=torch.int64, device=dev)
                rank_of_root[roots] = torch.arange(ncomp, device=dev)
                vrank = rank_of_root[label.long()]
                sel = torch.nonzero(vrank >= 0).flatten()
                vs = vrank[sel]                                               # (16-bit keys: two radix passes instead of eight)
                order = sel[torch.argsort(vs.to(torch.int16) if ncomp < 32768 else vs, stable=True)]      # new id -> old vertex id
                m = int(order.shape[0])
                comp_off = torch.zeros(ncomp + 1, dtype=torch.int64, device=dev)
                comp_off[1:] = torch.cumsum(sizes, 0)
                comp_of = torch.repeat_interleave(torch.arange(ncomp, device=dev), sizes)
            new_id = torch.full((n,), -1, dtype=torch.int32, device=dev)
            new_id[order] = torch.arange(m, dtype=torch.int32, device=dev)
            sub_xyz = cloud.xyz[order]
            sub_medial = medial[order].contiguous()
            sub_radius = radius[order].contiguous()
            # induced edges, renumbered (skeletonize.py:60-71): done inside the CSR build through a vertex map.  The CSR
            # is only read by the SSSP, which may number the graph as it likes (st_sssp orig_id): vertices in Z-order of
            # their medial points inside each component, so that a CTA's vertex range is a compact blob
            sperm = None
            if self.spatial_sssp and ncomp < 32768:
                sperm, srank = ops.spatial_order(sub_medial, comp_of)
                vmap = torch.full((n,), -1, dtype=torch.int32, device=dev)
                vmap[order] = srank
            else:
                vmap = new_id
            row_ptr, col, w = ops.csr_build(graph.edges, graph.edge_weights, m, vertex_map=vmap)
            # root of each component = first argmin of surface y (cloud.py:205-206)
            y = sub_xyz[:, 1].contiguous()
            if ncomp == 1:
                src = torch.argmin(y).reshape(1)
            elif ncomp <= 64:     # components are contiguous segments: a plain argmin per segment
                off_l = comp_off.tolist()
                src = torch.stack([torch.argmin(y[off_l[c]:off_l[c + 1]]) + off_l[c] for c in range(ncomp)])
            else:
                miny = torch.full((ncomp,), float("inf"), device=dev).scatter_reduce(0, comp_of, y, "amin")
                cand = torch.where(y == miny[comp_of], torch.arange(m, device=dev), torch.full((m,), m, device=dev))
                src = torch.full((ncomp,), m, dtype=torch.int64, device=dev).scatter_reduce(0, comp_of, cand, "amin")
        # skeletonize.py:73-78
        with section("skel.sssp"):
            # threshold step of the distance-ordered SSSP schedule (any value is exact; tools/sssp_sweep.py on the 6 m bench
            # tree: the best settings lie on a line of ~360 polls per metre of threshold step -- 0.06 m / 16 polls 3.0 ms,
            # 0.125 / 48 2.8, 0.18 / 64 2.6 -- fewer barriers win as long as the step stays a few edge lengths)
            delta = float(os.environ.get("ST_SSSP_DELTA", 9.0 * self.min_connection_length))
            if sperm is not None:
                dist, pred = ops.sssp(row_ptr, col, w, m, srank[src].contiguous(), delta=delta, orig_id=sperm)
            else:
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
            # skeleton ids = global (size-ordered) component indices, also when only this rank's share was extracted
            skeletons = self._emit(sub_medial, sub_radius, path, blen, bpar, cnb, cnp, off32, ncomp, post, comp_ids=self.component_ids)
        self.last = dict(keep=keep, order=order, comp_off=comp_off, pred=pred_local, dist=dist, tree_dist=tdist, roots=src,
                         path=path, branch_len=blen, branch_parent=bpar, comp_n_branches=cnb, comp_n_path=cnp,
                         n_components=ncomp, edges=graph.edges, edge_weights=graph.edge_weights)
        return DisjointTreeSkeleton(skeletons, post_applied=dict(post) if post else None)
