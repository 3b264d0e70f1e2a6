"""Pin the oracle's FRNN / cugraph restatements against scipy and
the accelerated oracle paths against their definitional brute-force forms."""
import heapq

import numpy as np
import pytest
from scipy import sparse
from scipy.sparse import csgraph
from scipy.spatial import cKDTree

from oracle import skeleton_ref as S


def test_knn_matches_bruteforce_and_ckdtree():
    rng = np.random.default_rng(0)
    p = rng.uniform(0, 1, (600, 3)).astype(np.float32)
    q = rng.uniform(0, 1, (200, 3)).astype(np.float32)
    for K, r in [(8, 0.15), (16, 0.3), (1, 0.05)]:
        i1, d1 = S.knn(q, p, K, r)
        i2, d2 = S.knn_bruteforce(q, p, K, r)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
        dd, ii = cKDTree(p.astype(np.float64)).query(q.astype(np.float64), k=K, distance_upper_bound=r)
        ii = np.where(np.isfinite(dd), ii, -1).reshape(len(q), K)
        assert (ii == i1).mean() > 0.999            # float64 vs fp32 only differs on near ties
    # self query returns self first with d=0; padding is idx -1 / d2 -1 -> sqrt = NaN
    i3, d3 = S.knn(p, p, 4, 0.01)
    assert np.array_equal(i3[:, 0], np.arange(len(p))) and np.all(d3[:, 0] == 0)
    assert np.isnan(S.knn_dist(d3)[i3 == -1]).all()


def test_knn_tie_break_lower_index():
    p = np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    i, d = S.knn(p[:1], p, 3, 2.0)
    assert i.tolist() == [[0, 1, 2]]


def _fp32_dijkstra(n, edges, w, root):
    adj = [[] for _ in range(n)]
    for (a, b), ww in zip(edges, w):
        if a != b:
            adj[a].append((b, np.float32(ww))); adj[b].append((a, np.float32(ww)))
    dist = [np.float32(np.inf)] * n
    dist[root] = np.float32(0)
    h = [(np.float32(0), root)]
    while h:
        d, u = heapq.heappop(h)
        if d > dist[u]:
            continue
        for v, ww in adj[u]:
            nd = np.float32(d + ww)
            if nd < dist[v]:
                dist[v] = nd; heapq.heappush(h, (nd, v))
    return np.array(dist, np.float32)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_sssp_is_fp32_dijkstra_and_close_to_scipy(seed):
    rng = np.random.default_rng(seed)
    p = rng.uniform(0, 1, (400, 3)).astype(np.float32)
    e, w = S.nn_graph(p, np.full(len(p), 0.2, np.float32), 6)
    pred, dist = S.sssp(len(p), e, w, 5)
    ref = _fp32_dijkstra(len(p), e, w, 5)
    reach = np.isfinite(ref)
    assert np.array_equal(dist[reach], ref[reach])
    assert np.all(dist[~reach] == np.finfo(np.float32).max) and np.all(pred[~reach] == -1)
    a = sparse.coo_matrix((w.astype(np.float64), (e[:, 0], e[:, 1])), shape=(len(p),) * 2).tocsr()
    d64 = csgraph.dijkstra(a.maximum(a.T), directed=False, indices=5)
    np.testing.assert_allclose(dist[reach], d64[reach], rtol=1e-5, atol=1e-6)
    # predecessor property: lowest u with dist[u]+w == dist[v]; tree reaches the root
    for v in np.nonzero(reach)[0]:
        if v == 5:
            assert pred[v] == -1; continue
        u = pred[v]
        cands = [a_ if b_ == v else b_ for (a_, b_), ww in zip(e, w)
                 if (a_ == v or b_ == v) and a_ != b_ and np.float32(dist[a_ if b_ == v else b_] + np.float32(ww)) == dist[v]]
        assert u == min(cands)


def test_connected_components_match_scipy_and_ordering():
    rng = np.random.default_rng(1)
    blobs = [rng.normal(c, 0.03, (n, 3)) for c, n in [((0, 0, 0), 80), ((1, 0, 0), 120), ((0, 1, 0), 40), ((1, 1, 1), 10)]]
    p = np.concatenate(blobs).astype(np.float32)
    p = p[rng.permutation(len(p))]
    e, w = S.nn_graph(p, np.full(len(p), 0.2, np.float32), 8)
    comps = S.connected_components(len(p), e, 32)
    assert [len(c) for c in comps] == sorted([len(c) for c in comps], reverse=True)
    assert all(np.all(np.diff(c) > 0) for c in comps)
    a = sparse.coo_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(len(p),) * 2)
    ncomp, lab = csgraph.connected_components(a, directed=False)
    big = sorted([int((lab == l).sum()) for l in range(ncomp) if (lab == l).sum() >= 32], reverse=True)
    assert big == [len(c) for c in comps]


def test_make_edges_quirks():
    p = np.array([[0, 0, 0], [0.01, 0, 0], [0.02, 0, 0]], np.float32)
    e, w = S.nn_graph(p, np.full(3, 0.05, np.float32), 3)
    assert not np.any(e[:, 1] == 0)                      # `idxs > 0`: vertex 0 never a destination
    assert [1, 1] in e.tolist() and [2, 2] in e.tolist()  # self edges kept
    assert [0, 0] not in e.tolist()


def test_tree_distances_running_sum():
    p = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [1, 1, 2], [5, 5, 5]], np.float32)
    pred = np.array([-1, 0, 1, 2, -1])
    d = S.tree_distances(p, pred, 0)
    assert d[:4].tolist() == [0, 1, 2, 4] and d[4] == np.finfo(np.float32).max


def _sample_tree_literal(pts, rad, preds, distances):
    """Line-by-line restatement of path.py:49-140 with brute-force nn (small n)."""
    n = len(pts)
    dist = distances.copy(); dist[~(preds > 0)] = -1
    term = set(); bids = np.full(n, -1); out = []; bid = 0
    while True:
        f = int(np.argmax(dist))
        if dist[f] <= 0: break
        path = []; i = f
        while i >= 0 and i not in term:
            path.append(i); i = int(preds[i])
        path = path[::-1]
        r = np.float32(rad[path].max())
        idx, d2 = S.knn_bruteforce(pts, pts[path], 1, r)
        idx, d = idx[:, 0], np.sqrt(np.where(d2[:, 0] < 0, np.nan, d2[:, 0]))
        on = np.zeros(n, bool)
        has = idx >= 0
        on[has] = d[has] < rad[np.asarray(path)[idx[has]]]
        ip = np.nonzero(on)[0]
        dist[ip] = -1; dist[path] = -1
        term |= set(ip.tolist()) | set(path)
        if len(path) < 2: continue
        out.append((bid, int(bids[i]), path))
        bids[path] = bid; bids[ip] = bid; bid += 1
    return out


@pytest.mark.parametrize("seed", [0, 3])
def test_sample_tree_matches_literal_restatement(seed):
    rng = np.random.default_rng(seed)
    # a Y-shaped set of noisy medial points
    t = rng.uniform(0, 1, 500)
    arm = rng.integers(0, 3, 500)
    dirs = np.array([[0, 1, 0], [0.7, 0.7, 0], [-0.6, 0.8, 0.2]])
    base = np.array([[0, 0, 0], [0, 1, 0], [0, 1, 0]])
    p = (base[arm] + dirs[arm] * t[:, None] + rng.normal(0, 0.004, (500, 3))).astype(np.float32)
    rad = np.where(arm == 0, 0.05, 0.03).astype(np.float32)
    e, w = S.nn_graph(p, np.maximum(rad, 0.02), 16)
    comp = S.connected_components(len(p), e, 32)[0]
    loc = np.full(len(p), -1); loc[comp] = np.arange(len(comp))
    sel = np.isin(e[:, 0], comp)
    root = int(np.argmin(p[comp, 1]))
    pred, _ = S.sssp(len(comp), loc[e[sel]], w[sel], root)
    dist = S.tree_distances(p[comp], pred, root)
    got = S.sample_tree(p[comp], rad[comp], pred, dist)
    ref = _sample_tree_literal(p[comp], rad[comp], pred, dist)
    assert len(got) == len(ref) and len(got) >= 2
    for g, r in zip(got, ref):
        assert (g.id, g.parent_id) == (r[0], r[1]) and g.path.tolist() == r[2]


def test_devoxelised_inference_restates_its_definition():
    """oracle.pipeline_ref.infer_points against a direct per-point definition: the point's block is the
    one whose inner half-open cube holds it, its voxel the one the block's voxeliser assigned it to, its label the
    network's output for that voxel (taken from oracle.infer's raw per-voxel predictions)."""
    import os
    import torch
    from oracle import pipeline_ref as P
    from oracle import unet_ref as U
    from smart_tree_b200 import synth
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sd = torch.load(os.path.join(root, "smart-tree_b200", "model", "weights", "noble-elevator-58_model_weights.pt"), map_location="cpu", weights_only=True)
    params = U.to_numpy_params(sd)
    xyz = P.centre_cloud(synth.make_tree(2, 2500).xyz)
    xyz = np.concatenate([xyz, xyz[:2] + np.float32(30)])               # two points in a block that is dropped (<= 20 points)
    rgb = np.zeros_like(xyz)
    got = P.infer_points(params, xyz, rgb, 0.02, 1.0, 0.4)
    raw = P.infer(params, xyz, rgb, 0.02, 1.0, 0.4, return_raw=True)["raw"]
    vmed = (np.exp(raw["preds"]["radius"]) * raw["preds"]["direction"]).astype(np.float32)
    centres, members = P.compute_blocks(xyz, 1.0, 0.4)
    first = np.concatenate([[0], np.cumsum([int((raw["coords"][:, 0] == b).sum()) for b in range(len(centres))])])
    seen = np.zeros(len(xyz), bool)
    for b, (c, m) in enumerate(zip(centres, members)):
        m = np.asarray(m)
        _, _, pcid, _ = P.voxelize_block(np.concatenate([xyz[m], rgb[m]], 1), 0.02)
        for j, p in enumerate(m):
            if P.cube_mask(xyz[p:p + 1], c, 1.0)[0] and pcid[j] >= 0:
                assert not seen[p]                                       # exactly one block's inner cube holds a point
                seen[p] = True
                assert np.array_equal(got["medial_vector"][p], vmed[first[b] + pcid[j]])
    assert np.array_equal(seen, got["class_l"] >= 0) and not seen[-2:].any() and seen[:-2].mean() > 0.9     # (sparse 1 m blocks of <= 20 points are dropped, dataset.py:178)


def test_sssp_oracle_above_46341_vertices():
    """Regression: scipy's dijkstra returns int32 predecessors, and `pred * n + v` overflowed for n > 46340 vertices --
    the oracle then produced distances below the true fixed point and predecessor chains that never reach the root
    (found by the full-size parity run: 245 529 skeleton vertices at BASELINE config C2)."""
    rng = np.random.default_rng(0)
    n = 60000
    pts = rng.uniform(0, 1, (n, 3)).astype(np.float32) * np.float32([4, 0.3, 0.3])
    idx, d2 = S.knn(pts, pts, 8, 0.08)
    src = np.repeat(np.arange(n), 8)
    ok = idx.reshape(-1) > 0
    e = np.stack([src[ok], idx.reshape(-1)[ok]], 1)
    w = np.sqrt(d2.reshape(-1)[ok]).astype(np.float32)
    comp = S.connected_components(n, e, 32)[0]
    assert len(comp) > 46341
    loc = np.full(n, -1, np.int64); loc[comp] = np.arange(len(comp))
    sel = np.isin(e[:, 0], comp)
    pred, dist = S.sssp(len(comp), loc[e[sel]], w[sel], 0)
    assert (dist < np.finfo(np.float32).max).all() and (pred >= 0).sum() == len(comp) - 1
    # the fixed point: nothing relaxes any further, and every predecessor edge is tight
    u, v = loc[e[sel]][:, 0], loc[e[sel]][:, 1]
    assert not (dist[u] + w[sel] < dist[v]).any() and not (dist[v] + w[sel] < dist[u]).any()
    td = S.tree_distances(pts[comp], pred, 0)
    assert (td < np.finfo(np.float32).max).all()             # every chain reaches the root
