"""Tree skeleton containers and post-processing
(/root/reference/smart_tree/data_types/tree.py:21-204): prune (:94-121), repair (:73-92),
smooth (:123-134).  `repair` batches every branch of a skeleton into ONE point->tube kernel launch
(st_points_to_tubes) instead of the reference's per-branch einsum with a host round trip."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List

import torch
import torch.nn.functional as F

from .branch import BranchSkeleton


class NodeStore:
    """Shared node array of a skeletoniser call: host [R,4] (xyz, radius) + its device twin."""

    def __init__(self, host, dev):
        self.host, self.dev = host, dev


@dataclass
class TreeSkeleton:
    _id: int
    branches: Dict[int, BranchSkeleton]

    def __len__(self):
        return len(self.branches)

    def __str__(self):
        return f"Tree Skeleton ({self._id}) has {len(self)} branches..."

    def to_tubes(self):
        return [t for b in self.branches.values() for t in b.to_tubes()]

    def _flat_store(self):
        """The shared NodeStore if every branch is still an untouched view into one (fast paths)."""
        store = None
        for b in self.branches.values():
            f = b._flat
            if f is None or (store is not None and f[0] is not store):
                return None
            if b.xyz.shape[0] != f[2] or b.xyz.data_ptr() != f[0].host[f[1] + 1].data_ptr():
                return None
            store = f[0]
        return store

    def repair(self, device=None):
        """Prepend to each branch the point of its parent branch's tube model nearest to the branch
        start (tree.py:73-92, queries.py:107-133).  The reference visits branches in dict order and a
        parent always precedes its children (parent ids are ids of earlier-emitted branches), so a
        child sees its parent already repaired.  The same result is obtained level by level: all
        branches of one tree depth are resolved with ONE st_points_to_tubes launch."""
        from .. import ops
        dev = torch.device(device) if device is not None else torch.device("cuda")
        ids = self.branches
        depth, waves = {}, {}
        for bid, br in ids.items():
            if br.parent_id not in ids or br.parent_id == bid:
                continue
            if br.parent_id > bid:      # would see an un-repaired parent in dict order: handle serially
                depth[bid] = None
                continue
            d = (depth.get(br.parent_id) or 0) + 1
            depth[bid] = d
            waves.setdefault(d, []).append(br)
        late = [ids[b] for b, d in depth.items() if d is None]
        order = [waves[d] for d in sorted(waves)] + ([late] if late else [])
        if not order:
            return
        store = self._flat_store()
        if store is not None and store.dev.device == dev:
            return self._repair_flat(order, store)
        for run in order:
            pts = torch.stack([br.xyz[0] for br in run])
            pa = [ids[br.parent_id] for br in run]
            a = torch.cat([p.xyz[:-1] for p in pa]); b_ = torch.cat([p.xyz[1:] for p in pa])
            r1 = torch.cat([p.radii[:-1].reshape(-1) for p in pa]); r2 = torch.cat([p.radii[1:].reshape(-1) for p in pa])
            off = torch.tensor([0] + [len(p) - 1 for p in pa], dtype=torch.int64).cumsum(0).int()
            f = lambda t: t.float().contiguous().to(dev, non_blocking=True)
            vec, _, _ = ops.points_to_tubes(f(pts), f(a), f(b_), f(r1), f(r2), off.to(dev))
            conn = pts + vec.cpu()
            for k, br in enumerate(run):
                br.xyz = torch.cat((conn[k:k + 1], br.xyz))
                br.radii = torch.cat((br.radii[[0]], br.radii))

    def _repair_flat(self, order, store):
        """repair() on the shared node array: the tubes of all parents of a wave are read straight from
        the device twin, connection points are written into the spare rows on the device (so the next
        wave sees repaired parents) and copied to the host once at the end."""
        from .. import ops
        nd = store.dev
        dev = nd.device
        ids = self.branches
        repaired = set()
        done_rows = []
        for run in order:
            q_row = torch.tensor([br._flat[1] for br in run], dtype=torch.int64)
            ps, pc = [], []
            for br in run:
                _, o, ln = ids[br.parent_id]._flat
                s0 = o if br.parent_id in repaired else o + 1
                ps.append(s0); pc.append(o + ln - s0)                      # first tube row, tube count
            meta = torch.tensor([ps, pc], dtype=torch.int64).to(dev, non_blocking=True)
            q_row_d = q_row.to(dev, non_blocking=True)
            cnt = meta[1]
            off = torch.zeros(len(run) + 1, dtype=torch.int64, device=dev)
            off[1:] = torch.cumsum(cnt, 0)
            tube = torch.repeat_interleave(meta[0] - off[:-1], cnt) + torch.arange(int(sum(pc)), device=dev)
            pts = nd[q_row_d + 1, :3].contiguous()
            a, b_ = nd[tube], nd[tube + 1]
            vec, _, _ = ops.points_to_tubes(pts, a[:, :3].contiguous(), b_[:, :3].contiguous(), a[:, 3].contiguous(),
                                            b_[:, 3].contiguous(), off.int())
            nd[q_row_d, :3] = pts + vec
            repaired.update(br._id for br in run)
            done_rows.append(q_row)
        rows = torch.cat(done_rows)
        store.host[rows, :3] = nd[rows.to(dev), :3].cpu()
        for run in order:
            for br in run:
                _, o, ln = br._flat
                br.xyz = store.host[o:o + ln + 1, :3]
                br.radii = store.host[o:o + ln + 1, 3:4]
                br._flat = None                                            # no longer the pristine view

    def branch_lengths(self):
        """Polyline length of every branch in one vectorised pass (== BranchSkeleton.length)."""
        bs = list(self.branches.values())
        store = self._flat_store()
        if store is not None:
            # contiguous rows of the shared array: per-branch sums of consecutive-row distances
            o = torch.tensor([b._flat[1] for b in bs]); n = torch.tensor([b._flat[2] for b in bs])
            seg = (store.host[1:, :3] - store.host[:-1, :3]).norm(dim=1)
            cs = torch.cat([seg.new_zeros(1, dtype=torch.float64), seg.double().cumsum(0)])
            return (cs[o + n] - cs[o + 1]).float()
        xyz = torch.cat([b.xyz for b in bs])
        n = torch.tensor([len(b) for b in bs])
        seg = (xyz[1:] - xyz[:-1]).norm(dim=1)
        end = n.cumsum(0)
        seg = torch.cat([seg, seg.new_zeros(1)])
        seg[end - 1] = 0                                   # no segment across a branch boundary
        cs = torch.cat([seg.new_zeros(1, dtype=torch.float64), seg.double().cumsum(0)])
        return (cs[end] - cs[end - n]).float()

    def prune(self, min_radius: float, min_length: float, root_id=None):
        """tree.py:94-121: drop branches that are too short / too thin, and everything below them."""
        if not self.branches:
            return TreeSkeleton(0, {})
        root_id = min(self.branches.keys()) if root_id is None else root_id
        lengths = self.branch_lengths().tolist()
        store = self._flat_store()
        if store is not None:
            o = torch.tensor([b._flat[1] for b in self.branches.values()]); n = torch.tensor([b._flat[2] for b in self.branches.values()])
            init_r = torch.maximum(store.host[o + 1, 3], store.host[o + n, 3]).tolist()
        else:
            init_r = [float(max(b.radii[0], b.radii[-1])) for b in self.branches.values()]
        keep = {root_id: self.branches[root_id]}
        remove = {}
        for (bid, b), length, ir in zip(self.branches.items(), lengths, init_r):
            if b.parent_id not in keep and b._id != root_id:
                remove[bid] = b
            elif length < min_length:
                remove[bid] = b
            elif ir < min_radius:
                remove[bid] = b
            else:
                keep[bid] = b
        self.branches = keep
        return TreeSkeleton(0, remove)

    def smooth(self, kernel_size=5):
        """Zero-padded box filter on the radii; turns radii [N,1] into [N] (quirk C-17).  All branches
        are filtered by one conv1d over their concatenation with kernel_size//2 zeros between them,
        which is arithmetically the per-branch `padding="same"` convolution of the reference."""
        todo = [b for b in self.branches.values() if b.radii.shape[0] > kernel_size]
        if not todo:
            return
        pad = kernel_size // 2
        z = torch.zeros(pad)
        sig = torch.cat([t for b in todo for t in (b.radii.reshape(-1).float(), z)])
        kernel = torch.ones(1, 1, kernel_size) / kernel_size
        out = F.conv1d(sig.reshape(1, 1, -1), kernel, padding="same").reshape(-1)
        o = 0
        for b in todo:
            n = b.radii.shape[0]
            b.radii = out[o:o + n].clone()
            o += n + pad

    @property
    def length(self):
        return torch.sum(torch.tensor([b.length for b in self.branches.values()]))

    @property
    def max_branch_id(self):
        return max(self.branches.keys())


@dataclass
class DisjointTreeSkeleton:
    skeletons: List[TreeSkeleton]

    def prune(self, min_radius, min_length):
        # only the first skeleton is pruned (tree.py:164-168, quirk C-18)
        if self.skeletons:
            self.skeletons[0].prune(min_radius=min_radius, min_length=min_length)

    def repair(self, device=None):
        for s in self.skeletons:
            s.repair(device=device)

    def smooth(self, kernel_size=7):
        for s in self.skeletons:
            s.smooth(kernel_size=kernel_size)
