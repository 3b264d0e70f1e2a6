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
    """Shared node array of a skeletoniser call: host [R,4] (xyz, radius) + a device twin made on demand."""

    def __init__(self, host, dev=None):
        self.host, self._dev = host, dev

    def device_twin(self, device):
        if self._dev is None or self._dev.device != torch.device(device):
            self._dev = self.host.to(device)
        return self._dev


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
        """The shared NodeStore if every branch is still an unmodified view into one (fast paths).
        `_flat` = (store, spare row, node count, connection point present)."""
        store = None
        for b in self.branches.values():
            f = b._flat
            if f is None or (store is not None and f[0] is not store):
                return None
            first = f[1] + (0 if f[3] else 1)
            if b.xyz.shape[0] != f[2] + (1 if f[3] else 0) or b.xyz.data_ptr() != f[0].host[first].data_ptr():
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
        if store is not None and dev.type == "cuda":
            return self._repair_flat(order, store, dev)
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

    def _repair_flat(self, order, store, dev):
        """repair() on the shared node array in ONE kernel launch (st_repair_branches): branches sorted by
        tree depth, a block barrier between depths so that children see repaired parents, connection
        points written into the spare rows on the device and copied to the host once."""
        from .. import ops
        nd = store.device_twin(dev)
        ids = self.branches
        listed = [br for run in order for br in run]
        in_order = {id(br) for br in listed}
        level0 = [br for br in ids.values() if id(br) not in in_order]
        seq = level0 + listed
        index = {br._id: i for i, br in enumerate(seq)}
        repaired_ids = {br._id for br in listed}
        row = [br._flat[1] for br in seq]
        ln = [br._flat[2] for br in seq]
        par = [index.get(br.parent_id, -1) if id(br) in in_order else -1 for br in seq]
        prep = [1 if (id(br) in in_order and br.parent_id in repaired_ids) else 0 for br in seq]
        level_off = [len(level0)]
        for run in order:
            level_off.append(level_off[-1] + len(run))
        meta = torch.tensor([row, ln, par], dtype=torch.int32).to(dev, non_blocking=True)
        ops.repair_branches(nd, meta[0].contiguous(), meta[1].contiguous(), meta[2].contiguous(),
                            torch.tensor(prep, dtype=torch.uint8).to(dev, non_blocking=True),
                            torch.tensor(level_off, dtype=torch.int32).to(dev, non_blocking=True))
        rows = torch.tensor([br._flat[1] for br in listed], dtype=torch.int64)
        store.host[rows] = nd[rows.to(dev)].cpu()
        host = store.host
        # views of the repaired branches (spare row included) from two split calls
        listed_sorted = sorted(listed, key=lambda b: b._flat[1])
        sizes, prev = [], 0
        for br in listed_sorted:
            _, o, n, _ = br._flat
            sizes += [o - prev, n + 1]
            prev = o + n + 1
        sizes.append(host.shape[0] - prev)
        xv = host[:, :3].split(sizes)[1::2]
        rv = host[:, 3:4].split(sizes)[1::2]
        for br, x, r in zip(listed_sorted, xv, rv):
            st, o, n, _ = br._flat
            br.xyz, br.radii = x, r
            br._flat = (st, o, n, True)

    def branch_lengths(self):
        """Polyline length of every branch in one vectorised pass (== BranchSkeleton.length)."""
        bs = list(self.branches.values())
        store = self._flat_store()
        if store is not None:
            # contiguous rows of the shared array: per-branch sums of consecutive-row distances
            first = torch.tensor([b._flat[1] + (0 if b._flat[3] else 1) for b in bs])
            last = torch.tensor([b._flat[1] + b._flat[2] for b in bs])
            seg = (store.host[1:, :3] - store.host[:-1, :3]).norm(dim=1)
            cs = torch.cat([seg.new_zeros(1, dtype=torch.float64), seg.double().cumsum(0)])
            return (cs[last] - cs[first]).float()
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
            first = torch.tensor([b._flat[1] + (0 if b._flat[3] else 1) for b in self.branches.values()])
            last = torch.tensor([b._flat[1] + b._flat[2] for b in self.branches.values()])
            init_r = torch.maximum(store.host[first, 3], store.host[last, 3]).tolist()
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
        store = self._flat_store()
        if store is not None and kernel_size % 2 == 1:
            return self._smooth_flat(todo, store, kernel_size)
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

    def _smooth_flat(self, todo, store, kernel_size):
        """Box filter of every branch at once on the shared array: windowed sums from one float64
        running sum, window clipped to the branch (== zero padding), divided by the kernel size."""
        h = kernel_size // 2
        first = torch.tensor([b._flat[1] + (0 if b._flat[3] else 1) for b in todo])
        cnt = torch.tensor([b.radii.shape[0] for b in todo])
        last = first + cnt - 1
        rows = torch.repeat_interleave(first - torch.cumsum(cnt, 0) + cnt, cnt) + torch.arange(int(cnt.sum()))
        lo = torch.maximum(rows - h, torch.repeat_interleave(first, cnt))
        hi = torch.minimum(rows + h, torch.repeat_interleave(last, cnt))
        cs = torch.cat([torch.zeros(1, dtype=torch.float64), store.host[:, 3].double().cumsum(0)])
        out = ((cs[hi + 1] - cs[lo]) / kernel_size).float()
        for b, r in zip(todo, out.split(cnt.tolist())):
            b.radii = r
            b._flat = None                       # radii no longer live in the shared array

    @property
    def length(self):
        return torch.sum(torch.tensor([b.length for b in self.branches.values()]))

    @property
    def max_branch_id(self):
        return max(self.branches.keys())


@dataclass
class DisjointTreeSkeleton:
    skeletons: List[TreeSkeleton]
    # post-processing already done on the device by Skeletonizer.forward(post=...): {"prune", "repair", "smooth"}
    post_applied: dict = None

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
