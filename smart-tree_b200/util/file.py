"""Loader / writer subset of /root/reference/smart_tree/util/file.py: `.npz` clouds (:156-167),
the skeleton npz schema (:73-116) and a dependency-free ASCII/binary-little-endian PLY reader
(the reference delegates PLY to open3d, which is a GUI dependency and out of scope)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from ..data_types.branch import BranchSkeleton
from ..data_types.cloud import Cloud
from ..data_types.tree import TreeSkeleton

_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4",
              "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4",
              "float32": "f4", "float64": "f8"}


def _read_ply(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"ply", "not a PLY file"
        fmt, n, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline().decode("ascii").strip()
            if line == "end_header":
                break
            tok = line.split()
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                props.append((tok[2], _PLY_TYPES[tok[1]]))
        if fmt == "ascii":
            data = np.loadtxt(f, max_rows=n, ndmin=2) if n else np.zeros((0, len(props)))
            cols = {name: data[:, i] for i, (name, _) in enumerate(props)}
        else:
            end = "<" if fmt == "binary_little_endian" else ">"
            rec = np.frombuffer(f.read(n * np.dtype([(nm, end + t) for nm, t in props]).itemsize),
                                dtype=np.dtype([(nm, end + t) for nm, t in props]), count=n)
            cols = {name: rec[name] for name, _ in props}
    xyz = np.stack([cols["x"], cols["y"], cols["z"]], 1).astype(np.float64)
    if all(c in cols for c in ("red", "green", "blue")):
        rgb = np.stack([cols["red"], cols["green"], cols["blue"]], 1).astype(np.float64)
        if dict(props)["red"] in ("u1", "i1", "u2", "i2"):          # integer colour channels are 0..255 (open3d does the same)
            rgb = rgb / 255.0
    else:
        rgb = np.zeros_like(xyz)
    return xyz, rgb


def load_cloud(path: Path, pin_memory: bool = False) -> Cloud:
    """pin_memory: page-lock the host tensors (when a CUDA runtime is present) so that `cloud.to_device(dev,
    non_blocking=True)` is an asynchronous copy -- what the pipeline's loader wants."""
    path = Path(path)
    if path.suffix == ".npz":
        cloud = Cloud.from_numpy(**np.load(path))
    elif path.suffix == ".ply":
        xyz, rgb = _read_ply(path)
        cloud = Cloud.from_numpy(xyz=xyz, rgb=rgb)
    else:
        raise ValueError(f"unsupported cloud format {path.suffix} (supported: .npz, .ply)")
    if pin_memory:
        import torch
        if torch.cuda.is_available():
            cloud = cloud.pin_memory()
    return cloud


def skeleton_arrays(skeleton: TreeSkeleton) -> dict:
    """The reference's skeleton schema (file.py:73-93); also the wire format of the multi-GPU gather."""
    bs = list(skeleton.branches.values())
    return {
        "tree_id": np.asarray(skeleton._id),
        "skeleton_xyz": np.concatenate([np.asarray(b.xyz) for b in bs]) if bs else np.zeros((0, 3), np.float32),
        "skeleton_radii": (np.concatenate([np.asarray(b.radii).reshape(-1) for b in bs])[..., np.newaxis]
                           if bs else np.zeros((0, 1), np.float32)),
        "branch_id": np.asarray([b._id for b in bs], np.int64),
        "branch_parent_id": np.asarray([b.parent_id for b in bs], np.int64),
        "branch_num_elements": np.asarray([len(b) for b in bs], np.int64),
    }


def save_skeleton(skeleton: TreeSkeleton, save_location):
    np.savez(save_location, **skeleton_arrays(skeleton))


def skeleton_from_arrays(data, tree_id=0) -> TreeSkeleton:
    import torch
    sizes = np.asarray(data["branch_num_elements"]).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(sizes)])
    xyz, radii = np.asarray(data["skeleton_xyz"]), np.asarray(data["skeleton_radii"]).reshape(-1, 1)
    branches = {}
    for i, (bid, par) in enumerate(zip(data["branch_id"], data["branch_parent_id"])):
        s = slice(off[i], off[i + 1])
        branches[int(bid)] = BranchSkeleton(int(bid), int(par), torch.as_tensor(xyz[s]).float(), torch.as_tensor(radii[s]).float())
    return TreeSkeleton(int(tree_id), branches)


def load_skeleton(path) -> TreeSkeleton:
    return skeleton_from_arrays(np.load(path))
