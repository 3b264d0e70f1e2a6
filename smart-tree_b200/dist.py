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
    """-> (float32 payload [P,4] = xyz,radius per node ; int64 header [B,4] = unit, skeleton, branch id, parent, + counts)."""
    nodes, meta = [], []
    for unit, sk in zip(unit_ids, skeletons):
        a = skeleton_arrays(sk)
        if len(a["branch_id"]) == 0:
            continue
        nodes.append(np.concatenate([a["skeleton_xyz"].astype(np.float32), a["skeleton_radii"].astype(np.float32).reshape(-1, 1)], 1))
        m = np.stack([np.full(len(a["branch_id"]), unit), np.full(len(a["branch_id"]), int(a["tree_id"])), a["branch_id"],
                      a["branch_parent_id"], a["branch_num_elements"]], 1).astype(np.int64)
        meta.append(m)
    nodes = np.concatenate(nodes) if nodes else np.zeros((0, 4), np.float32)
    meta = np.concatenate(meta) if meta else np.zeros((0, 5), np.int64)
    return torch.from_numpy(nodes), torch.from_numpy(meta)


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


def gather_skeletons(local: List[DisjointTreeSkeleton], unit_ids: List[int], device=None):
    """All-gather every rank's skeletons (NCCL on GPUs, gloo on CPU).  `local[i]` is the result for
    global unit `unit_ids[i]`.  Returns {(unit, skeleton id): TreeSkeleton} on every rank."""
    sk, units = [], []
    for d, u in zip(local, unit_ids):
        for s in d.skeletons:
            sk.append(s); units.append(u)
    nodes, meta = pack_skeletons(sk, units)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return unpack_skeletons(nodes, meta)
    world = dist.get_world_size()
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    counts = torch.tensor([nodes.shape[0], meta.shape[0]], dtype=torch.int64, device=dev)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    all_counts = torch.stack(all_counts).cpu()
    max_n, max_m = int(all_counts[:, 0].max()), int(all_counts[:, 1].max())
    nbuf = torch.zeros((max(max_n, 1), 4), dtype=torch.float32, device=dev); nbuf[:nodes.shape[0]] = nodes.to(dev)
    mbuf = torch.zeros((max(max_m, 1), 5), dtype=torch.int64, device=dev); mbuf[:meta.shape[0]] = meta.to(dev)
    ng = [torch.zeros_like(nbuf) for _ in range(world)]
    mg = [torch.zeros_like(mbuf) for _ in range(world)]
    dist.all_gather(ng, nbuf)
    dist.all_gather(mg, mbuf)
    out = {}
    for r in range(world):
        out.update(unpack_skeletons(ng[r][:int(all_counts[r, 0])], mg[r][:int(all_counts[r, 1])]))
    return out
