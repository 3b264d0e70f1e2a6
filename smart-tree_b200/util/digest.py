"""Result digest of a list of TreeSkeletons: what bench.py prints and the full-size parity tests verify against
the oracle (tests/test_gpu_fullsize.py), so that a benchmark number can be tied to a checked result."""
from __future__ import annotations

import hashlib

import numpy as np


def skeleton_digest(skeletons) -> dict:
    """{"branches": B, "nodes": P, "topology": sha1 over (skeleton id, branch id, parent id, node count) rows,
    "coords_1e-4": sha1 over the node coordinates rounded to 1e-4 m (the north star's coordinate tolerance)}."""
    topo, xyz = [], []
    for s in skeletons:
        for bid in sorted(s.branches.keys()):
            b = s.branches[bid]
            topo.append((int(s._id), int(bid), int(b.parent_id), int(b.xyz.shape[0])))
            xyz.append(np.asarray(b.xyz, dtype=np.float64))
    t = np.asarray(topo, dtype=np.int64).reshape(-1, 4)
    x = np.concatenate(xyz) if xyz else np.zeros((0, 3))
    q = np.round(x * 1e4).astype(np.int64)
    return {"branches": int(t.shape[0]), "nodes": int(x.shape[0]), "topology": hashlib.sha1(t.tobytes()).hexdigest()[:16],
            "coords_1e-4": hashlib.sha1(q.tobytes()).hexdigest()[:16]}
