"""Multi-GPU plumbing: one process per GPU, independent units (trees / blocks) sharded over
ranks with no data-path collective; the only exchange is the final gather of packed skeletons
(SURVEY §8e), using the reference's own skeleton schema (util/file.py:73-93) as wire format."""
from __future__ import annotations

import os
from typing import List

import numpy as np
import torch
import torch.distributed as dist

from .data_types.tree import DisjointTreeSkeleton, TreeSkeleton
from .util.file import skeleton_arrays, skeleton_from_arrays


def init_from_env(backend=None):
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard(items: List, rank: int, world: int) -> List:
    """Unit i goes to rank i % world (callers sort units by descending size first for balance)."""
    return items[rank::world]


def pack_skeletons(skeletons: List[TreeSkeleton], unit_ids: List[int]):
    """-> float32 node table [P,4] (xyz, radius) and int64 branch table [B,5] = (unit, skeleton id, branch id,
    parent id, node count): the reference's npz skeleton schema (util/file.py:73-93) plus a unit column."""
    xyz, rad, meta = [], [], []
    for unit, sk in zip(unit_ids, skeletons):
        for b in sk.branches.values():
            xyz.append(b.xyz)
            rad.append(b.radii.reshape(-1))
            meta.append((unit, sk._id, b._id, b.parent_id, b.xyz.shape[0]))
    if not meta:
        return torch.zeros((0, 4), dtype=torch.float32), torch.zeros((0, 5), dtype=torch.int64)
    nodes = torch.cat([torch.cat(xyz).float(), torch.cat(rad).float().unsqueeze(1)], 1)
    return nodes, torch.tensor(meta, dtype=torch.int64)


def unpack_skeletons(nodes: torch.Tensor, meta: torch.Tensor):
    """Inverse of pack_skeletons -> {(unit, skeleton id): TreeSkeleton}."""
    nodes, meta = nodes.cpu().numpy(), meta.cpu().numpy()
    out = {}
    off = np.concatenate([[0], np.cumsum(meta[:, 4])]) if len(meta) else np.zeros(1, np.int64)
    keys = {}
    for i, (unit, sid, bid, par, cnt) in enumerate(meta.tolist()):
        keys.setdefault((unit, sid), []).append(i)
    for key, rows in keys.items():
        data = {"branch_id": meta[rows, 2], "branch_parent_id": meta[rows, 3], "branch_num_elements": meta[rows, 4],
                "skeleton_xyz": np.concatenate([nodes[off[r]:off[r + 1], :3] for r in rows]),
                "skeleton_radii": np.concatenate([nodes[off[r]:off[r + 1], 3:4] for r in rows])}
        out[key] = skeleton_from_arrays(data, tree_id=key[1])
    return out


class GatheredSkeletons:
    """Result of the gather: every rank's packed tables, materialised into TreeSkeleton objects on demand
    (building thousands of Python objects on every rank is not part of the exchange)."""

    def __init__(self, tables):
        self.tables = tables                      # list of (nodes [P,4], meta [B,5]) per rank, CPU tensors

    @property
    def n_branches(self):
        return int(sum(m.shape[0] for _, m in self.tables))

    @property
    def n_nodes(self):
        return int(sum(n.shape[0] for n, _ in self.tables))

    def skeletons(self):
        out = {}
        for nodes, meta in self.tables:
            out.update(unpack_skeletons(nodes, meta))
        return out

    # dict-like access for callers that want the objects
    def __getitem__(self, key):
        return self.skeletons()[key]

    def keys(self):
        return self.skeletons().keys()

    def items(self):
        return self.skeletons().items()

    def __len__(self):
        return len({(int(u), int(s)) for _, m in self.tables for u, s in m[:, :2].tolist()})


def gather_skeletons(local: List[DisjointTreeSkeleton], unit_ids: List[int], device=None) -> GatheredSkeletons:
    """All-gather every rank's skeletons (NCCL on GPUs, gloo on CPU).  `local[i]` is the result for global
    unit `unit_ids[i]`.  Every rank receives every rank's packed tables."""
    sk, units = [], []
    for d, u in zip(local, unit_ids):
        for s in d.skeletons:
            sk.append(s); units.append(u)
    nodes, meta = pack_skeletons(sk, units)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return GatheredSkeletons([(nodes, meta)])
    world = dist.get_world_size()
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    counts = torch.tensor([nodes.shape[0], meta.shape[0]], dtype=torch.int64, device=dev)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    all_counts = torch.stack(all_counts).cpu()
    max_n, max_m = int(all_counts[:, 0].max()), int(all_counts[:, 1].max())
    # one padded payload per rank: node table followed by the branch table viewed as float32 pairs
    nbuf = torch.zeros((max(max_n, 1), 4), dtype=torch.float32, device=dev); nbuf[:nodes.shape[0]] = nodes.to(dev)
    mbuf = torch.zeros((max(max_m, 1), 5), dtype=torch.int64, device=dev); mbuf[:meta.shape[0]] = meta.to(dev)
    ng = [torch.empty_like(nbuf) for _ in range(world)]
    mg = [torch.empty_like(mbuf) for _ in range(world)]
    dist.all_gather(ng, nbuf)
    dist.all_gather(mg, mbuf)
    return GatheredSkeletons([(ng[r][:int(all_counts[r, 0])].cpu(), mg[r][:int(all_counts[r, 1])].cpu()) for r in range(world)])


# ---------------------------------------------------------------------------------------------- packed gather (one collective)
class GatheredPacked:
    """Every rank's packed skeleton buffer after ONE all-gather (device-resident int32 [world, cap]); host copy and the
    reference's object model only on demand."""

    def __init__(self, buf, world, cap=None, redo=None, works=None):
        self.buf, self.world = buf, world
        self.cap = int(buf.shape[1]) if cap is None else int(cap)
        self._redo = redo
        self._works = works if works is not None else []
        self._host = None
        self.host_bytes = 0

    def wait(self):
        """Join the (asynchronous) collective: afterwards `buf` may be read on the current stream."""
        while self._works:
            self._works.pop(0).wait()
        return self

    def to_host(self):
        """Read the `world` length words, then ONE device->host copy of the used prefixes (not of the capacity-padded
        buffer) -> list of (unit, PackedSkeletons) per rank.  Collective if a rank's result overflowed (gather_packed)."""
        if self._host is None:
            from .data_types.packed import PackedSkeletons
            self.wait()
            lens = [int(v) for v in self.buf[:, 0].cpu().tolist()]
            if max(lens) > self.cap:
                if self._redo is None:
                    raise RuntimeError("gathered skeleton buffer overflowed and cannot be re-exchanged")
                self.buf = self._redo(max(lens))
                self.cap = max(lens)
                self.wait()
            flat = torch.cat([self.buf[r, :max(lens[r], 8)] for r in range(self.world)]).cpu().numpy()
            self.host_bytes = int(flat.nbytes + 4 * self.world)
            out, o = [], 0
            for r in range(self.world):
                m = max(lens[r], 8)
                out.append(PackedSkeletons.from_wire(flat[o:o + m]))
                o += m
            self._host = out
        return self._host

    @property
    def n_branches(self):
        return int(sum(p.nb for _, p in self.to_host()))

    def skeletons(self):
        """{(unit, skeleton id): TreeSkeleton} over all ranks."""
        out = {}
        for unit, p in self.to_host():
            for s in p:
                out[(unit, s._id)] = s
        return out

    def __len__(self):
        return sum(len(p) for _, p in self.to_host())


def gather_packed(local: DisjointTreeSkeleton, unit: int, capacity: int = 1 << 20, device=None) -> GatheredPacked:
    """All-gather of the packed result of one skeletoniser call per rank (the path's only exchange).
    The device buffer st_finish_skeletons wrote is shipped as it is -- no Python re-packing of branches: ONE all-gather
    of `capacity` int32 words per rank (the used prefix carries its own length), no host synchronisation: the collective
    overlaps whatever the caller does next.  The lengths are only read when the result is brought to the host
    (GatheredPacked.to_host: the `world` length words, then the used prefixes in one copy); if a rank's result did not fit,
    to_host repeats the collective once with the largest size -- every rank sees the same lengths, so every rank takes the
    same branch, but every rank has to call it (rare: the default capacity holds ~200 k nodes)."""
    from .data_types.packed import PackedSkeletons
    packed = local.skeletons
    if not isinstance(packed, PackedSkeletons):
        if len(packed) != 0:
            raise TypeError("gather_packed needs the packed result of Skeletonizer.forward (use gather_skeletons for object lists)")
        d0 = device if device is not None else torch.device("cpu")
        packed = PackedSkeletons(0, 0, [], [], None, torch.zeros(0, dtype=torch.int32, device=d0), False)      # nothing to skeletonise
    world = dist.get_world_size() if dist.is_initialized() else 1
    wire = packed.wire(unit)
    dev = device if device is not None else wire.device
    if world > 1 and dist.get_backend() != "nccl":
        dev = torch.device("cpu")
    wire = wire.to(dev)
    need = int(wire.numel())

    def exchange(cap):
        send = torch.zeros(cap, dtype=torch.int32, device=dev)
        m = min(need, cap)
        send[:m] = wire[:m]
        send[0] = need                        # true length, also when truncated: the receivers see the overflow
        if world == 1:
            return send.unsqueeze(0)
        recv = torch.empty((world, cap), dtype=torch.int32, device=dev)
        if dist.get_backend() == "nccl":
            # (async_op=True, the compute stream not waiting for the collective, was measured and bought nothing: the pipeline's
            # stage timers synchronise the device, and NCCL's kernel shares the SMs with the resident-grid SSSP launch)
            dist.all_gather_into_tensor(recv, send)
        else:
            dist.all_gather(list(recv.unbind(0)), send)
        return recv

    works = []
    cap = max(int(capacity), 16)
    return GatheredPacked(exchange(cap), world, cap, exchange, works)


# ---------------------------------------------------------------------------------------------- labelled voxels (plots)
def labelled_part(lc, voxel_block):
    """One rank's labelled voxels as two tables: f[n,9] = xyz, rgb, medial_vector and i[n,2] = class, global block."""
    n = lc.xyz.shape[0]
    rgb = lc.rgb if lc.rgb is not None else torch.zeros_like(lc.xyz)
    f = torch.cat([lc.xyz.float(), rgb.float(), lc.medial_vector.float()], 1) if n else torch.zeros((0, 9), device=lc.xyz.device)
    i = torch.stack([lc.class_l.reshape(-1).long(), voxel_block.long()], 1) if n else torch.zeros((0, 2), dtype=torch.int64, device=lc.xyz.device)
    return f.contiguous(), i.contiguous()


def merge_labelled(parts, device=None):
    """All ranks' parts -> one labelled Cloud in BLOCK order (stable: voxels keep their order inside a block), i.e.
    exactly the cloud ModelInference.forward returns on one GPU, whatever the number of ranks."""
    from .data_types.cloud import Cloud
    f = torch.cat([p[0] for p in parts])
    i = torch.cat([p[1] for p in parts])
    if device is not None:
        f, i = f.to(device), i.to(device)
    order = torch.argsort(i[:, 1], stable=True)
    f, i = f[order], i[order]
    return Cloud(xyz=f[:, 0:3].contiguous(), rgb=f[:, 3:6].contiguous(), medial_vector=f[:, 6:9].contiguous(),
                 class_l=i[:, 0:1].contiguous())


def gather_labelled(part, device=None):
    """All-gather of every rank's labelled voxels (counts first, then payloads padded to the largest): the one
    exchange of the block-sharded plot path (~25 B per voxel; NCCL over NVLink on GPUs, gloo on CPU)."""
    f, i = part
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [part]
    world = dist.get_world_size()
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    cnt = torch.tensor([f.shape[0]], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    cnts = [int(c) for c in torch.cat(cnts).cpu()]
    m = max(max(cnts), 1)
    fb = torch.zeros((m, 9), dtype=torch.float32, device=dev); fb[:f.shape[0]] = f.to(dev)
    ib = torch.zeros((m, 2), dtype=torch.int64, device=dev); ib[:i.shape[0]] = i.to(dev)
    fg = [torch.empty_like(fb) for _ in range(world)]
    ig = [torch.empty_like(ib) for _ in range(world)]
    dist.all_gather(fg, fb)
    dist.all_gather(ig, ib)
    return [(fg[r][:cnts[r]], ig[r][:cnts[r]]) for r in range(world)]
