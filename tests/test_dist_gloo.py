"""Multi-process host logic on CPU (gloo, world_size 2): unit sharding and the skeleton gather
(pack -> all_gather -> unpack) used by bench.py / smart_tree_b200.dist at N > 1."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smart_tree_b200 import dist as stdist
from smart_tree_b200.data_types.branch import BranchSkeleton
from smart_tree_b200.data_types.tree import DisjointTreeSkeleton, TreeSkeleton


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_skeleton(unit, nbranch):
    g = torch.Generator().manual_seed(unit)
    br = {}
    for b in range(nbranch):
        n = 2 + (unit + b) % 5
        br[b] = BranchSkeleton(b, b - 1, torch.randn(n, 3, generator=g), torch.rand(n, 1, generator=g))
    return DisjointTreeSkeleton([TreeSkeleton(0, br), TreeSkeleton(1, {0: BranchSkeleton(0, -1, torch.randn(3, 3, generator=g), torch.rand(3, 1, generator=g))})])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = stdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    units = stdist.shard(list(range(5)), rank, world)            # 5 units over 2 ranks: uneven on purpose
    local = [_fake_skeleton(u, 3 + u) for u in units]
    if rank == 1:
        local[-1] = DisjointTreeSkeleton([])                         # a unit with no skeleton at all
    merged = stdist.gather_skeletons(local, units, device=torch.device("cpu"))
    assert merged.n_branches > 0 and len(merged) == len(merged.skeletons())
    out[rank] = {k: (len(v.branches), float(sum(b.xyz.sum() + b.radii.sum() for b in v.branches.values()))) for k, v in merged.items()}
    dist.destroy_process_group()


def test_skeleton_gather_world2():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] == out[1]                                          # every rank holds the same merged result
    keys = sorted(out[0].keys())
    expect = {}
    for u in range(5):
        if u == 3:                                                   # rank 1's last unit (units 1, 3) was emptied
            continue
        sk = _fake_skeleton(u, 3 + u)
        for s in sk.skeletons:
            expect[(u, s._id)] = (len(s.branches), float(sum(b.xyz.sum() + b.radii.sum() for b in s.branches.values())))
    assert keys == sorted(expect.keys())
    for k in keys:
        assert out[0][k][0] == expect[k][0] and abs(out[0][k][1] - expect[k][1]) < 1e-4


def test_pack_unpack_roundtrip_single_process():
    sk = _fake_skeleton(7, 6)
    nodes, meta = stdist.pack_skeletons(sk.skeletons, [4, 4])
    back = stdist.unpack_skeletons(nodes, meta)
    assert sorted(back.keys()) == [(4, 0), (4, 1)]
    for s in sk.skeletons:
        got = back[(4, s._id)]
        for bid, b in s.branches.items():
            assert got.branches[bid].parent_id == b.parent_id
            assert torch.equal(got.branches[bid].xyz, b.xyz) and torch.equal(got.branches[bid].radii, b.radii)


def test_shard_covers_all_units():
    units = list(range(11))
    parts = [stdist.shard(units, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == units
