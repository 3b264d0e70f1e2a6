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


# ---------------------------------------------------------------- block-sharded plots: labelled-voxel exchange
def _fake_labelled(nblocks, seed=0):
    """A labelled voxel cloud in block order, as ModelInference.forward returns it, with its per-voxel block ids."""
    from smart_tree_b200.data_types.cloud import Cloud
    g = torch.Generator().manual_seed(seed)
    per = [int(v) for v in torch.randint(0, 7, (nblocks,), generator=g)]           # some blocks contribute nothing
    blk = torch.repeat_interleave(torch.arange(nblocks), torch.tensor(per))
    n = int(blk.shape[0])
    lc = Cloud(xyz=torch.randn(n, 3, generator=g), rgb=torch.rand(n, 3, generator=g), medial_vector=torch.randn(n, 3, generator=g),
               class_l=torch.randint(0, 2, (n, 1), generator=g))
    return lc, blk


def _labelled_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    stdist.init_from_env(backend="gloo")
    lc, blk = _fake_labelled(11)
    mine = (blk % world) == rank                                    # this rank labelled blocks rank, rank+world, ...
    part = stdist.labelled_part(lc.filter(mine), blk[mine])
    merged = stdist.merge_labelled(stdist.gather_labelled(part, device=torch.device("cpu")))
    out[rank] = (merged.xyz.numpy().tobytes(), merged.rgb.numpy().tobytes(), merged.medial_vector.numpy().tobytes(),
                 merged.class_l.numpy().tobytes())
    dist.destroy_process_group()


def test_labelled_voxel_exchange_world2_restores_block_order():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_labelled_worker, args=(2, port, out), nprocs=2, join=True)
    lc, _ = _fake_labelled(11)
    expect = (lc.xyz.numpy().tobytes(), lc.rgb.numpy().tobytes(), lc.medial_vector.numpy().tobytes(), lc.class_l.numpy().tobytes())
    assert out[0] == expect and out[1] == expect                    # bit-identical to the unsharded cloud on every rank


def test_merge_labelled_any_world_single_process():
    lc, blk = _fake_labelled(23, seed=3)
    for world in (1, 3, 8):
        parts = [stdist.labelled_part(lc.filter((blk % world) == r), blk[(blk % world) == r]) for r in range(world)]
        m = stdist.merge_labelled(parts)
        assert torch.equal(m.xyz, lc.xyz) and torch.equal(m.medial_vector, lc.medial_vector) and torch.equal(m.class_l, lc.class_l)


# ---------------------------------------------------------------- packed gather: the device buffer of st_finish_skeletons, one collective
def _fake_packed(unit, nbranch, post_done=True):
    """A PackedSkeletons as Skeletonizer._emit builds it (include/st_b200.h: bmeta | nodes | smooth), two components."""
    from smart_tree_b200.data_types.packed import PackedSkeletons
    g = np.random.default_rng(unit)
    lens = [2 + (unit + b) % 5 for b in range(nbranch)] + [3]
    cnb = np.array([nbranch, 1], np.int32)
    nb, nrow = len(lens), sum(lens) + len(lens)
    bmeta = np.zeros((nb, 4), np.int32)
    row = 0
    for b, ln in enumerate(lens):
        local = b if b < nbranch else 0
        bmeta[b] = (row, ln, local - 1, 1 | (2 if local > 0 and post_done else 0) | (4 if ln > 4 and post_done else 0))
        row += ln + 1
    nodes = g.standard_normal((nrow, 4)).astype(np.float32)
    smooth = g.uniform(0, 1, nrow).astype(np.float32)
    payload = np.concatenate([bmeta.reshape(-1), nodes.view(np.int32).reshape(-1), smooth.view(np.int32)])
    return PackedSkeletons(nb, nrow, cnb, [0, 1], payload, torch.from_numpy(payload.copy()), post_done)


def _summary(packed_list):
    out = {}
    for unit, p in packed_list:
        for s in p:
            out[(unit, s._id)] = (len(s.branches), float(sum(float(b.xyz.sum()) + float(b.radii.sum()) for b in s.branches.values())),
                                  [b.parent_id for b in s.branches.values()])
    return out


def _packed_worker(rank, world, port, out, capacity):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    stdist.init_from_env(backend="gloo")
    local = DisjointTreeSkeleton(_fake_packed(10 + rank, 3 + 4 * rank))
    g = stdist.gather_packed(local, unit=10 + rank, capacity=capacity, device=torch.device("cpu"))
    assert g.n_branches == sum(3 + 4 * r + 1 for r in range(world)) and len(g) == 2 * world
    out[rank] = _summary(g.to_host())
    dist.destroy_process_group()


@pytest.mark.parametrize("capacity", [1 << 16, 64])        # 64 words: every rank overflows -> the collective is repeated at the largest size
def test_packed_gather_world2(capacity):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_packed_worker, args=(2, port, out, capacity), nprocs=2, join=True)
    expect = _summary([(10 + r, _fake_packed(10 + r, 3 + 4 * r)) for r in range(2)])
    assert out[0] == expect and out[1] == expect


def test_packed_wire_roundtrip_and_lazy_objects():
    from smart_tree_b200.data_types.packed import PackedSkeletons
    for post_done in (True, False):
        p = _fake_packed(3, 5, post_done)
        assert p._objs is None and len(p) == 2 and bool(p)                  # nothing materialised yet
        unit, q = PackedSkeletons.from_wire(p.wire(unit=42).numpy())
        assert unit == 42 and q.nb == p.nb and q.nrow == p.nrow and q.comp_ids == [0, 1]
        assert _summary([(0, p)]) == _summary([(0, q)])
        sk = p[0]
        assert sorted(sk.branches) == list(range(5)) and p._objs is not None
        b1 = sk.branches[1]
        assert b1.parent_id == 0 and b1.xyz.shape[0] == (2 + (3 + 1) % 5) + (1 if post_done else 0)      # + the repair connection point
