"""Oracle: the reference's Python glue around the network and the skeletonizer, on the CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  /root/reference/smart_tree/dataset/augmentations.py:38-41   CentreCloud
  /root/reference/smart_tree/dataset/dataset.py:166-226       compute_blocks, __getitem__
  /root/reference/smart_tree/util/maths.py:135-155            cube_filter
  /root/reference/smart_tree/model/sparse.py:40-61            batch_collate
  /root/reference/smart_tree/model/model_inference.py:49-100  ModelInference.forward
  /root/reference/smart_tree/pipeline.py:55-106               process_cloud / post_process
  /root/reference/smart_tree/data_types/tree.py:73-134        repair / prune / smooth
  /root/reference/smart_tree/util/queries.py:89-133           pts_to_nearest_tube_gpu
Blocks are processed in index order (the reference shuffles them by accident, Appendix C-1), so
clouds must be compared order-insensitively against the real reference.
"""
from __future__ import annotations

import numpy as np
import torch

from . import skeleton_ref as S
from . import unet_ref as U

F32 = np.float32


def centre_cloud(xyz: np.ndarray) -> np.ndarray:
    xyz = xyz.astype(F32)
    lo, hi = xyz.min(0), xyz.max(0)
    half = (hi - lo) / F32(2)
    centre = lo + half
    shift = -centre + np.array([0, half[1], 0], F32)
    return xyz + shift


def cube_mask(points, centre, cube_size):
    lo = centre.astype(F32) - F32(cube_size / 2)
    hi = centre.astype(F32) + F32(cube_size / 2)
    return np.all((points >= lo) & (points < hi), axis=1)


def compute_blocks(xyz, block_size=4, buffer_size=0.4, min_points=20):
    """dataset.py:166-190.  Returns block centres [B,3] and the list of point-index arrays."""
    q = torch.div(torch.from_numpy(xyz), block_size, rounding_mode="floor")
    ids, counts = torch.unique(q, return_counts=True, dim=0)
    ids = ids[counts > min_points]
    centres = (ids * block_size + (block_size / 2)).numpy()
    members = [np.nonzero(cube_mask(xyz, c, block_size + buffer_size * 2))[0] for c in centres]
    return centres, members


def round_half_away(q):
    t = np.trunc(q)
    return t + ((q - t) >= F32(0.5))


def voxelize_block(pts6, voxel_size):
    """dataset.py:192-216: range = the block cloud's own bounding box."""
    lo, hi = pts6[:, :3].min(0), pts6[:, :3].max(0)
    return point_to_voxel_exact(pts6, voxel_size, lo, hi)


def point_to_voxel_exact(points, vsize, lo, hi):
    """unet_ref.point_to_voxel with std::round evaluated exactly (trunc + fraction test)."""
    points = np.ascontiguousarray(points, F32)
    vs = F32(vsize)
    lo = np.asarray(lo, F32)
    hi = np.asarray(hi, F32)
    grid = round_half_away((hi - lo) / vs).astype(np.int64)
    c = np.floor((points[:, :3] - lo) / vs).astype(np.int64)
    ok = np.all((c >= 0) & (c < grid), axis=1)
    lin = (c[:, 2] * (grid[1] + 1) + c[:, 1]) * (grid[0] + 1) + c[:, 0]
    idx_ok = np.nonzero(ok)[0]
    uniq, first, inv = np.unique(lin[ok], return_index=True, return_inverse=True)
    appearance = np.argsort(first, kind="stable")
    rank = np.empty(len(uniq), np.int64)
    rank[appearance] = np.arange(len(uniq))
    pcid = np.full(len(points), -1, np.int64)
    pcid[idx_ok] = rank[inv]
    rep = idx_ok[first[appearance]]
    return points[rep], c[rep][:, ::-1].astype(np.int32), pcid, rep


def infer(params, xyz, rgb, voxel_size=0.01, block_size=4, buffer_size=0.4, eps=1e-4, dtype=np.float32, return_raw=False):
    """ModelInference.forward: returns the masked labelled voxel cloud (xyz, rgb, medial_vector, class)."""
    centres, members = compute_blocks(xyz, block_size, buffer_size)
    feats, coords, masks = [], [], []
    for b, (c, m) in enumerate(zip(centres, members)):
        vox, zyx, _, _ = voxelize_block(np.concatenate([xyz[m], rgb[m]], 1), voxel_size)
        feats.append(vox)
        coords.append(np.concatenate([np.full((len(zyx), 1), b, np.int32), zyx], 1))
        masks.append(cube_mask(vox[:, :3], c, block_size))
    if not feats:
        z = np.zeros((0, 3), F32)
        return dict(xyz=z, rgb=z, medial_vector=z, class_l=np.zeros(0, np.int64))
    feats, coords, masks = np.concatenate(feats), np.concatenate(coords), np.concatenate(masks)
    out = U.forward(params, feats[:, :3], coords, eps=eps, dtype=dtype)
    medial = (np.exp(out["radius"]) * out["direction"]).astype(F32)
    cls = out["class_l"].argmax(1)
    res = dict(xyz=feats[masks, :3], rgb=feats[masks, 3:6], medial_vector=medial[masks], class_l=cls[masks])
    if return_raw:
        res["raw"] = dict(feats=feats, coords=coords, mask=masks, preds=out)
    return res


def infer_points(params, xyz, rgb, voxel_size=0.01, block_size=4, buffer_size=0.4, eps=1e-4, dtype=np.float32):
    """Devoxelised inference: every input point takes the prediction of its voxel
    (pc_voxel_id of voxelize_block, which the reference computes in dataset.py:214 and drops) from the block whose
    inner half-open cube contains the point; class -1 / zero vector where no block covers it."""
    centres, members = compute_blocks(xyz, block_size, buffer_size)
    medial = np.zeros((len(xyz), 3), F32)
    cls = np.full(len(xyz), -1, np.int64)
    feats, coords, ids = [], [], []
    off = 0
    for b, (c, m) in enumerate(zip(centres, members)):
        vox, zyx, pcid, _ = voxelize_block(np.concatenate([xyz[m], rgb[m]], 1), voxel_size)
        feats.append(vox)
        coords.append(np.concatenate([np.full((len(zyx), 1), b, np.int32), zyx], 1))
        inner = cube_mask(xyz[m], c, block_size) & (pcid >= 0)
        ids.append((np.asarray(m)[inner], pcid[inner] + off))
        off += len(vox)
    if not feats:
        return dict(medial_vector=medial, class_l=cls)
    feats, coords = np.concatenate(feats), np.concatenate(coords)
    out = U.forward(params, feats[:, :3], coords, eps=eps, dtype=dtype)
    vmed = (np.exp(out["radius"]) * out["direction"]).astype(F32)
    vcls = out["class_l"].argmax(1)
    for pts, rows in ids:
        medial[pts] = vmed[rows]
        cls[pts] = vcls[rows]
    return dict(medial_vector=medial, class_l=cls)


# ----------------------------------------------------------------------------- post-processing
def branch_length(xyz):
    d = xyz[1:] - xyz[:-1]
    return F32(np.sqrt((d * d).sum(1, dtype=F32)).sum(dtype=F32)) if len(xyz) > 1 else F32(0)


def prune(branches: dict, min_radius, min_length):
    """tree.py:94-121 on {id: (parent_id, xyz, radii)}; returns the kept dict."""
    root_id = min(branches.keys())
    keep = {root_id: branches[root_id]}
    for bid, (par, xyz, rad) in branches.items():
        if par not in keep and bid != root_id:
            continue
        if branch_length(xyz) < min_length:
            continue
        if max(rad[0], rad[-1]) < min_radius:
            continue
        keep[bid] = (par, xyz, rad)
    return keep


def nearest_tube_vector(p, xyz, rad):
    """queries.py:89-133 for one point against the tubes of one branch (fp32, einsum order)."""
    a, b = xyz[:-1].astype(F32), xyz[1:].astype(F32)
    r1, r2 = rad[:-1].astype(F32), rad[1:].astype(F32)
    ab = b - a
    ap = p.astype(F32)[None] - a
    with np.errstate(invalid="ignore", divide="ignore"):
        t = np.clip((ap * ab).sum(1, dtype=F32) / (ab * ab).sum(1, dtype=F32), 0, 1).astype(F32)
    proj = a + t[:, None] * ab
    d = np.sqrt(((proj - p) ** 2).sum(1, dtype=F32))
    r = (1 - t) * r1 + t * r2
    i = int(np.argmin(np.abs(d - r)))
    return proj[i] - p, i


def repair(branches: dict):
    """tree.py:73-92 (in dict order, each child sees its parent's current geometry)."""
    ids = set(branches.keys())
    for bid in list(branches.keys()):
        par, xyz, rad = branches[bid]
        if par not in ids:
            continue
        _, pxyz, prad = branches[par]
        v, _ = nearest_tube_vector(xyz[0], pxyz, prad)
        branches[bid] = (par, np.concatenate([(xyz[0] + v)[None], xyz]).astype(F32), np.concatenate([rad[:1], rad]))
    return branches


def smooth(branches: dict, kernel_size):
    """tree.py:123-134: zero-padded box filter."""
    k = np.ones(kernel_size, F32) / F32(kernel_size)
    for bid, (par, xyz, rad) in branches.items():
        if len(rad) > kernel_size:
            branches[bid] = (par, xyz, np.convolve(rad.astype(F32), k, mode="same").astype(F32))
    return branches


def process_cloud(params, xyz, rgb, voxel_size=0.01, block_size=4, buffer_size=0.4, K=16, min_connection_length=0.02,
                  minimum_graph_vertices=32, branch_classes=(0,), prune_skeletons=True, min_skeleton_radius=0.01,
                  min_skeleton_length=0.02, repair_skeletons=True, smooth_skeletons=True, smooth_kernel_size=11, eps=1e-4,
                  labelled=None):
    """Pipeline.process_cloud (pipeline.py:55-106).  `labelled` short-circuits the network (used to
    give the oracle and the CUDA skeletoniser bit-identical inputs)."""
    if labelled is None:
        labelled = infer(params, centre_cloud(xyz), rgb, voxel_size, block_size, buffer_size, eps)
    sel = np.isin(labelled["class_l"], list(branch_classes))
    skels = S.skeletonize(labelled["xyz"][sel], labelled["medial_vector"][sel], K, min_connection_length, minimum_graph_vertices)
    out = []
    for s in skels:
        br = {b.id: (b.parent_id, b.xyz, b.radii) for b in s.branches}
        out.append(br)
    if prune_skeletons and out and out[0]:
        out[0] = prune(out[0], min_skeleton_radius, min_skeleton_length)
    if repair_skeletons:
        out = [repair(b) for b in out]
    if smooth_skeletons:
        out = [smooth(b, smooth_kernel_size) for b in out]
    return labelled, skels, out
