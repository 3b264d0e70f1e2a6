"""Oracle: FRNN / cugraph semantics + the reference skeletonizer on the CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Parity unpinned at the FRNN / cugraph
boundary; every (L)-level choice is *defined* here and the CUDA
path must reproduce it bit for bit:
  * d2 = (dx*dx + dy*dy) + dz*dz in fp32, no FMA contraction; dist = IEEE sqrt(d2)
  * kNN: neighbours with d2 < fl32(r*r), ascending (d2, index); missing -> idx -1
  * SSSP: dist = least fixed point of d[v] = min_u fl32(d[u] + w(u,v)) (== fp32 Dijkstra);
    predecessor = lowest u != v with fl32(d[u] + w(u,v)) == d[v]; -1 at the root
  * components ordered by size descending, then smallest vertex id
  * argmax / argmin -> first extremum

Follows
  /root/reference/smart_tree/skeleton/graph.py:12-60        knn, nn, nn_graph, make_edges
  /root/reference/smart_tree/skeleton/filter.py:6-11        outlier_removal
  /root/reference/smart_tree/data_types/graph.py:32-66      connected_cugraph_components
  /root/reference/smart_tree/skeleton/skeletonize.py:31-95  Skeletonizer.forward / process_subgraph
  /root/reference/smart_tree/skeleton/shortest_path.py:12-21,46-74  shortest_paths, pred_graph
  /root/reference/smart_tree/skeleton/path.py:9-140         trace_route, select_path_points, sample_tree
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy import sparse
from scipy.sparse import csgraph
from scipy.spatial import cKDTree

F32 = np.float32


def d2_f32(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = a.astype(F32, copy=False)
    b = b.astype(F32, copy=False)
    dx = a[..., 0] - b[..., 0]
    dy = a[..., 1] - b[..., 1]
    dz = a[..., 2] - b[..., 2]
    return (dx * dx + dy * dy) + dz * dz


# ----------------------------------------------------------------------------- B7
def knn_bruteforce(src, dst, K, r):
    """Definitional kNN (small inputs only)."""
    r2 = F32(F32(r) * F32(r))
    n = len(src)
    idx = np.full((n, K), -1, np.int64)
    d2o = np.full((n, K), -1, F32)
    for i in range(n):
        d2 = d2_f32(src[i][None], dst)
        cand = np.nonzero(d2 < r2)[0]
        o = np.lexsort((cand, d2[cand]))[:K]
        idx[i, :len(o)] = cand[o]
        d2o[i, :len(o)] = d2[cand[o]]
    return idx, d2o


def knn(src, dst, K, r, slack=8, tree=None):
    """frnn.frnn_grid_points restated: K nearest of dst within r for each src row.
    Returns idx[N,K] (-1 padded) and SQUARED distances d2[N,K] (-1 padded), sorted by
    (d2, idx).  Candidates come from a float64 KD-tree (K+slack of them), the decision is
    made on fp32 d2 exactly as the CUDA kernel computes it."""
    src = np.ascontiguousarray(src, F32)
    dst = np.ascontiguousarray(dst, F32)
    n, m = len(src), len(dst)
    idx = np.full((n, K), -1, np.int64)
    d2o = np.full((n, K), -1, F32)
    if n == 0 or m == 0:
        return idx, d2o
    r2 = F32(F32(r) * F32(r))
    tree = tree or cKDTree(dst.astype(np.float64))
    kk = min(K + slack, m)
    _, ci = tree.query(src.astype(np.float64), k=kk, distance_upper_bound=float(r) * (1 + 1e-5) + 1e-9, workers=-1)
    if kk == 1:
        ci = ci[:, None]
    miss = ci >= m
    ci_safe = np.where(miss, 0, ci)
    d2 = d2_f32(src[:, None, :], dst[ci_safe])
    bad = miss | ~(d2 < r2)
    d2s = np.where(bad, np.inf, d2).astype(F32)
    cis = np.where(bad, np.iinfo(np.int64).max, ci_safe)
    o = np.lexsort((cis, d2s), axis=1)[:, :K]
    ti = np.take_along_axis(cis, o, 1)
    td = np.take_along_axis(d2s, o, 1)
    ok = np.isfinite(td)
    kq = min(K, kk)
    idx[:, :kq] = np.where(ok, ti, -1)
    d2o[:, :kq] = np.where(ok, td, F32(-1))
    return idx, d2o


def knn_dist(d2):
    """graph.py:26  dists.sqrt(): NaN for the -1 padding (Appendix C-7)."""
    with np.errstate(invalid="ignore"):
        return np.sqrt(d2.astype(F32))


# ----------------------------------------------------------------------------- filter / graph
def outlier_removal(points, radii, nb_points=8):
    """filter.py:6-11.  radii [N] raw (unclamped) radius; keep iff all nb_points nearest
    (self included) exist and lie strictly inside the point's own radius."""
    radii = radii.astype(F32)
    idx, d2 = knn(points, points, nb_points, float(radii.max()) if len(radii) else 0.0)
    d = knn_dist(d2)
    with np.errstate(invalid="ignore"):
        keep = (d < radii[:, None]) & (idx != -1)
    return keep.sum(1) == nb_points


def nn_graph(points, radii, K=16):
    """graph.py:36-40,52-60.  Returns directed candidate edges [E,2] i64 and weights [E] f32.
    Neighbour dropped if d > radius_i (NaN never drops, but idx is already -1); destination
    0 is never emitted (`idxs > 0`, Appendix C-6); self edges are emitted."""
    radii = radii.astype(F32)
    idx, d2 = knn(points, points, K, float(radii.max()) if len(radii) else 0.0)
    d = knn_dist(d2)
    with np.errstate(invalid="ignore"):
        idx = np.where(d > radii[:, None], -1, idx)
    n = len(points)
    parent = np.repeat(np.arange(n, dtype=np.int64), K)
    flat = idx.reshape(-1)
    valid = flat > 0
    return np.stack([parent[valid], flat[valid]], 1), d.reshape(-1)[valid]


# ----------------------------------------------------------------------------- B8
def connected_components(n, edges, minimum_vertices=32):
    """data_types/graph.py:32-51.  Vertex set 0..n-1 (renumber=False), weakly connected
    labels; components with >= minimum_vertices vertices, size desc then min vertex id.
    Returns list of ascending vertex-id arrays."""
    if n == 0:
        return []
    a = sparse.coo_matrix((np.ones(len(edges), np.int8), (edges[:, 0], edges[:, 1])), shape=(n, n))
    _, lab = csgraph.connected_components(a, directed=False)
    order = np.argsort(lab, kind="stable")
    counts = np.bincount(lab)
    starts = np.concatenate([[0], np.cumsum(counts)])
    comps = [order[starts[l]:starts[l + 1]] for l in range(len(counts)) if counts[l] >= minimum_vertices]
    comps.sort(key=lambda v: (-len(v), v[0]))
    return comps


def sssp(n, edges, weights, root):
    """cugraph.sssp on the undirected weighted graph, fp32 (see module docstring)."""
    w = weights.astype(F32)
    u = np.concatenate([edges[:, 0], edges[:, 1]]).astype(np.int64)
    v = np.concatenate([edges[:, 1], edges[:, 0]]).astype(np.int64)
    ww = np.concatenate([w, w])
    nl = u != v                                   # self loops never relax anything
    u, v, ww = u[nl], v[nl], ww[nl]
    # duplicate undirected edges collapse (weights are equal by construction: d(i,j) == d(j,i))
    key = u * n + v
    o = np.lexsort((ww, key))
    key, u, v, ww = key[o], u[o], v[o], ww[o]
    first = np.ones(len(key), bool)
    first[1:] = key[1:] != key[:-1]
    key, u, v, ww = key[first], u[first], v[first], ww[first]
    dist = np.full(n, np.inf, F32)
    dist[root] = 0
    if len(u):
        a = sparse.csr_matrix((ww.astype(np.float64) + 1e-30, (u, v)), shape=(n, n))
        # float64 Dijkstra tree -> fp32 running sums along it: an upper bound of the fp32 fixed point
        d64, p64 = csgraph.dijkstra(a, directed=True, indices=root, return_predecessors=True)
        p64 = p64.astype(np.int64)             # scipy returns int32: `par * n + v` below overflowed for n > 46340 vertices
        done = np.zeros(n, bool)
        done[root] = True
        pending = np.nonzero(p64 >= 0)[0]
        while len(pending):
            par = p64[pending]
            rdy = done[par]
            vv, pp = pending[rdy], par[rdy]
            ew = ww[np.searchsorted(key, pp * n + vv)]
            dist[vv] = dist[pp] + ew
            done[vv] = True
            pending = pending[~rdy]
        # Bellman-Ford sweeps in fp32 down to the least fixed point
        while True:
            nd = dist.copy()
            np.minimum.at(nd, v, dist[u] + ww)
            if np.array_equal(nd, dist):
                break
            dist = nd
    # predecessor = lowest u with fl32(dist[u] + w) == dist[v]
    pred = np.full(n, -1, np.int64)
    if len(u):
        hit = ((dist[u] + ww) == dist[v]) & np.isfinite(dist[v])
        big = np.full(n, np.iinfo(np.int64).max, np.int64)
        np.minimum.at(big, v[hit], u[hit])
        pred = np.where(big == np.iinfo(np.int64).max, -1, big)
    pred[root] = -1
    dist_out = np.where(np.isfinite(dist), dist, np.finfo(F32).max).astype(F32)
    return pred, dist_out


def tree_distances(points, pred, root):
    """skeletonize.py:80-85 + shortest_path.py:46-55: SSSP over the predecessor tree with
    weights ||p_v - p_pred(v)|| == root->leaf fp32 running sum along the tree."""
    n = len(pred)
    w = np.sqrt(d2_f32(points, points[np.maximum(pred, 0)])).astype(F32)
    dist = np.full(n, np.finfo(F32).max, F32)
    dist[root] = 0
    done = np.zeros(n, bool)
    done[root] = True
    pending = np.nonzero((pred >= 0))[0]
    while len(pending):
        rdy = done[pred[pending]]
        if not rdy.any():
            break
        vv = pending[rdy]
        dist[vv] = dist[pred[vv]] + w[vv]
        done[vv] = True
        pending = pending[~rdy]
    return dist


# ----------------------------------------------------------------------------- sample_tree
@dataclass
class Branch:
    id: int
    parent_id: int
    path: np.ndarray          # vertex indices (component-local), root side first
    xyz: np.ndarray = field(default=None)
    radii: np.ndarray = field(default=None)


def sample_tree(medial_pts, medial_radii, preds, distances):
    """path.py:49-140, exact fp32 decisions.  medial_radii [n]."""
    pts = np.ascontiguousarray(medial_pts, F32)
    rad = medial_radii.astype(F32)
    n = len(pts)
    dist = distances.astype(F32).copy()
    dist[~(preds > 0)] = -1                                   # path.py:71-72
    allocated = np.zeros(n, bool)                             # termination_pts
    branch_ids = np.full(n, -1, np.int64)
    tree = cKDTree(pts.astype(np.float64))
    branches = []
    bid = 0
    while n:
        f = int(np.argmax(dist))
        if not dist[f] > 0:
            break
        path = []
        i = f
        while i >= 0 and not allocated[i]:                    # trace_route
            path.append(i)
            i = int(preds[i])
        term = i
        path = np.asarray(path[::-1], np.int64)
        # select_path_points: nearest path vertex within r = max path radius, then own-radius test
        r = F32(rad[path].max())
        r2 = F32(r * r)
        ball = tree.query_ball_point(pts[path].astype(np.float64), float(r) * (1 + 1e-5) + 1e-9)
        best_d2 = {}
        cand_p = np.concatenate([np.asarray(b, np.int64) for b in ball]) if len(ball) else np.zeros(0, np.int64)
        cand_j = np.concatenate([np.full(len(b), j, np.int64) for j, b in enumerate(ball)]) if len(ball) else cand_p
        if len(cand_p):
            d2 = d2_f32(pts[cand_p], pts[path[cand_j]])
            ok = d2 < r2
            cand_p, cand_j, d2 = cand_p[ok], cand_j[ok], d2[ok]
            o = np.lexsort((cand_j, d2, cand_p))               # per point: min d2, then lowest path position
            cand_p, cand_j, d2 = cand_p[o], cand_j[o], d2[o]
            firstp = np.ones(len(cand_p), bool)
            firstp[1:] = cand_p[1:] != cand_p[:-1]
            cand_p, cand_j, d2 = cand_p[firstp], cand_j[firstp], d2[firstp]
            on = np.sqrt(d2) < rad[path[cand_j]]
            idx_points = cand_p[on]
        else:
            idx_points = cand_p
        dist[idx_points] = -1
        dist[path] = -1
        allocated[idx_points] = True
        allocated[path] = True
        if len(path) < 2:
            continue
        parent = int(branch_ids[term])                         # term == -1 wraps to the last vertex (C-10)
        branches.append(Branch(bid, parent, path, pts[path].copy(), rad[path].copy()))
        branch_ids[path] = bid
        branch_ids[idx_points] = bid
        bid += 1
    return branches


# ----------------------------------------------------------------------------- Skeletonizer
@dataclass
class Skeleton:
    id: int
    vertex_ids: np.ndarray     # indices into the (outlier-filtered) cloud
    root: int
    preds: np.ndarray
    distances: np.ndarray
    branches: list


def skeletonize(xyz, medial_vector, K=16, min_connection_length=0.02, minimum_graph_vertices=32,
                return_filtered=False):
    """Skeletonizer.forward (skeletonize.py:31-55)."""
    xyz = np.ascontiguousarray(xyz, F32)
    mv = np.ascontiguousarray(medial_vector, F32)
    medial = xyz + mv
    radius = np.sqrt((mv[:, 0] * mv[:, 0] + mv[:, 1] * mv[:, 1]) + mv[:, 2] * mv[:, 2])   # cloud.py:255-256
    keep = outlier_removal(medial, radius, 8) if len(xyz) else np.zeros(0, bool)
    xyz, medial, radius = xyz[keep], medial[keep], radius[keep]
    edges, w = nn_graph(medial, np.maximum(radius, F32(min_connection_length)), K) if len(xyz) else (np.zeros((0, 2), np.int64), np.zeros(0, F32))
    comps = connected_components(len(xyz), edges, minimum_graph_vertices)
    lab = np.full(len(xyz), -1, np.int64)
    for ci, vids in enumerate(comps):
        lab[vids] = ci
    skeletons = []
    e_lab = lab[edges[:, 0]] if len(edges) else np.zeros(0, np.int64)
    for ci, vids in enumerate(comps):
        local = np.full(len(xyz), -1, np.int64)
        local[vids] = np.arange(len(vids))
        sel = e_lab == ci
        le = local[edges[sel]]
        root = int(np.argmin(xyz[vids, 1]))                   # cloud.py:205-206 on the sub-cloud
        pred, _ = sssp(len(vids), le, w[sel], root)
        dist = tree_distances(medial[vids], pred, root)
        br = sample_tree(medial[vids], radius[vids], pred, dist)
        skeletons.append(Skeleton(ci, vids, root, pred, dist, br))
    if return_filtered:
        return skeletons, keep
    return skeletons
