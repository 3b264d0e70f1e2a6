"""Deterministic synthetic tube-tree clouds (SURVEY.md §8d "Synthetic inputs").

The reference ships no input clouds (its dataset is an external download,
/root/reference/README.md:24), so every benchmark and parity case in this repo uses
this generator.  A tree is a recursive set of tapered tubes; points are sampled on the
tube surfaces with 2 mm Gaussian noise.  Ground-truth medial vectors (surface point ->
tube axis) and a class label (0 = branch, 1 = foliage) are returned as well so the
skeleton stage can be exercised without the network.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SynthTree:
    xyz: np.ndarray            # [N,3] f32 metres
    rgb: np.ndarray            # [N,3] f32 zeros
    medial_vector: np.ndarray  # [N,3] f32 ground truth (point -> axis)
    class_l: np.ndarray        # [N]   i64 0 branch / 1 foliage
    segments: np.ndarray       # [S,8] f64 a(3) b(3) r0 r1


def _unit(v):
    return v / np.linalg.norm(v)


def tube_segments(rng: np.random.Generator, levels: int = 5, children: int = 3) -> np.ndarray:
    """Trunk from the origin along +y (r=0.15 m, 2 m long, tapering to 0.7 r), then
    ``children`` children per segment for ``levels`` levels."""
    segs = []
    frontier = [(np.zeros(3), np.array([0.0, 1.0, 0.0]), 0.15, 2.0)]
    for level in range(levels + 1):
        nxt = []
        for a, d, r0, length in frontier:
            b = a + d * length
            segs.append(np.concatenate([a, b, [r0, 0.7 * r0]]))
            if level == levels:
                continue
            for _ in range(children):
                cd = _unit(d + rng.normal(0.0, 0.5, 3))
                t = rng.uniform(0.4, 1.0)
                nxt.append((a + d * length * t, cd, 0.6 * r0, 0.7 * length))
        frontier = nxt
    return np.asarray(segs)


def _frame(d):
    ref = np.array([1.0, 0.0, 0.0]) if abs(d[0]) < 0.9 else np.array([0.0, 0.0, 1.0])
    u = _unit(np.cross(d, ref))
    return u, np.cross(d, u)


def make_tree(seed: int = 0, n_points: int = 50_000, foliage_fraction: float = 0.0,
              noise: float = 0.002, offset=(0.0, 0.0, 0.0)) -> SynthTree:
    rng = np.random.default_rng(seed)
    segs = tube_segments(rng)
    n_fol = int(round(n_points * foliage_fraction))
    n_branch = n_points - n_fol

    length = np.linalg.norm(segs[:, 3:6] - segs[:, 0:3], axis=1)
    w = length * (segs[:, 6] + segs[:, 7])
    quota = w / w.sum() * n_branch
    counts = np.floor(quota).astype(np.int64)
    rem = n_branch - counts.sum()
    counts[np.argsort(-(quota - counts), kind="stable")[:rem]] += 1

    xyz, mv = [], []
    for s, c in zip(segs, counts):
        if c == 0:
            continue
        a, b, r0, r1 = s[0:3], s[3:6], s[6], s[7]
        d = _unit(b - a)
        u, v = _frame(d)
        t = rng.uniform(0.0, 1.0, c)
        th = rng.uniform(0.0, 2 * np.pi, c)
        r = r0 + (r1 - r0) * t
        radial = np.cos(th)[:, None] * u + np.sin(th)[:, None] * v
        axis = a + t[:, None] * (b - a)
        p = axis + r[:, None] * radial + rng.normal(0.0, noise, (c, 3))
        xyz.append(p)
        # medial vector = from the (noisy) surface point to its projection on the axis
        tt = np.clip(((p - a) @ (b - a)) / ((b - a) @ (b - a)), 0.0, 1.0)
        mv.append(a + tt[:, None] * (b - a) - p)
    cls = [np.zeros(n_branch, np.int64)]
    if n_fol:
        leaves = segs[-(3 ** 5):]
        tip = leaves[rng.integers(0, len(leaves), n_fol), 3:6]
        fp = tip + rng.normal(0.0, 0.15, (n_fol, 3))
        xyz.append(fp)
        mv.append(np.zeros((n_fol, 3)))
        cls.append(np.ones(n_fol, np.int64))
    xyz = np.concatenate(xyz) + np.asarray(offset)
    mv = np.concatenate(mv)
    cls = np.concatenate(cls)
    perm = rng.permutation(len(xyz))  # scanner order is not segment order
    return SynthTree(xyz[perm].astype(np.float32), np.zeros((len(xyz), 3), np.float32),
                     mv[perm].astype(np.float32), cls[perm], segs)


def make_forest(seeds, n_points_each: int, pitch: float = 5.0, cols: int = 8, **kw) -> SynthTree:
    """Trees on a grid in the x-z plane (C5: 40 trees, 8x5 grid, 5 m pitch)."""
    parts = [make_tree(s, n_points_each, offset=((i % cols) * pitch, 0.0, (i // cols) * pitch), **kw)
             for i, s in enumerate(seeds)]
    return SynthTree(np.concatenate([p.xyz for p in parts]), np.concatenate([p.rgb for p in parts]),
                     np.concatenate([p.medial_vector for p in parts]),
                     np.concatenate([p.class_l for p in parts]),
                     np.concatenate([p.segments for p in parts]))
