"""Loader / writer mirrors of /root/reference/smart_tree/util/file.py:73-116,156-167 and data_types/cloud.py:234-252:
`.npz` clouds (current and legacy `vector` key), ASCII and binary PLY, the skeleton npz schema round trip."""
import struct

import numpy as np
import pytest
import torch

from smart_tree_b200.data_types.branch import BranchSkeleton
from smart_tree_b200.data_types.tree import TreeSkeleton
from smart_tree_b200.util.file import load_cloud, load_skeleton, save_skeleton, skeleton_arrays


def _cloud(n=37, seed=0):
    g = np.random.default_rng(seed)
    return g.normal(size=(n, 3)), g.uniform(size=(n, 3)), g.normal(size=(n, 3)) * 0.05, g.integers(0, 2, n)


def test_load_cloud_npz_current_and_legacy_keys(tmp_path):
    xyz, rgb, mv, cls = _cloud()
    np.savez(tmp_path / "a.npz", xyz=xyz, rgb=rgb, medial_vector=mv, class_l=cls)
    c = load_cloud(tmp_path / "a.npz")
    assert c.xyz.dtype == torch.float32 and tuple(c.xyz.shape) == (37, 3)
    np.testing.assert_allclose(c.xyz.numpy(), xyz.astype(np.float32))
    np.testing.assert_allclose(c.rgb.numpy(), rgb.astype(np.float32))
    np.testing.assert_allclose(c.medial_vector.numpy(), mv.astype(np.float32))
    assert np.array_equal(c.class_l.numpy().reshape(-1), cls)
    np.savez(tmp_path / "legacy.npz", xyz=xyz, rgb=rgb, vector=mv, class_l=cls)      # cloud.py:234-252: `vector` is the old name
    np.testing.assert_allclose(load_cloud(tmp_path / "legacy.npz").medial_vector.numpy(), mv.astype(np.float32))
    np.savez(tmp_path / "bare.npz", xyz=xyz, rgb=rgb)
    b = load_cloud(tmp_path / "bare.npz")
    assert b.medial_vector is None and len(b) == 37


def _write_ply(path, xyz, rgb_u8, fmt):
    n = len(xyz)
    hdr = ("ply\nformat %s 1.0\ncomment test\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
           "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face 0\nproperty list uchar int vertex_indices\n"
           "end_header\n" % (fmt, n)).encode()
    with open(path, "wb") as f:
        f.write(hdr)
        for p, c in zip(xyz, rgb_u8):
            if fmt == "ascii":
                f.write(("%r %r %r %d %d %d\n" % (float(p[0]), float(p[1]), float(p[2]), c[0], c[1], c[2])).encode())
            else:
                f.write(struct.pack("<fffBBB" if fmt == "binary_little_endian" else ">fffBBB", p[0], p[1], p[2], c[0], c[1], c[2]))


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
def test_load_cloud_ply(tmp_path, fmt):
    xyz = np.random.default_rng(1).normal(size=(23, 3)).astype(np.float32)
    rgb = np.random.default_rng(2).integers(0, 256, (23, 3)).astype(np.uint8)
    _write_ply(tmp_path / "c.ply", xyz, rgb, fmt)
    c = load_cloud(tmp_path / "c.ply")
    np.testing.assert_allclose(c.xyz.numpy(), xyz, rtol=1e-6)
    np.testing.assert_allclose(c.rgb.numpy(), rgb / 255.0, rtol=1e-6)      # uchar colours -> [0, 1], decided by the property type


def test_ply_dark_uchar_colours_and_empty_cloud(tmp_path):
    xyz = np.zeros((3, 3), np.float32)
    _write_ply(tmp_path / "dark.ply", xyz, np.ones((3, 3), np.uint8), "binary_little_endian")      # all channels <= 1: still /255
    np.testing.assert_allclose(load_cloud(tmp_path / "dark.ply").rgb.numpy(), 1 / 255.0, rtol=1e-6)
    for fmt in ("ascii", "binary_little_endian"):
        _write_ply(tmp_path / "empty.ply", xyz[:0], np.zeros((0, 3), np.uint8), fmt)
        assert len(load_cloud(tmp_path / "empty.ply")) == 0
    with pytest.raises(ValueError):
        load_cloud(tmp_path / "cloud.xyz")


def test_skeleton_npz_schema_round_trip(tmp_path):
    g = torch.Generator().manual_seed(0)
    br = {b: BranchSkeleton(b, b // 2 - 1 if b else -1, torch.randn(3 + b, 3, generator=g), torch.rand(3 + b, 1, generator=g)) for b in (0, 1, 2, 5)}
    sk = TreeSkeleton(3, br)
    arr = skeleton_arrays(sk)
    # the reference's keys and shapes (util/file.py:73-93)
    assert set(arr) == {"tree_id", "skeleton_xyz", "skeleton_radii", "branch_id", "branch_parent_id", "branch_num_elements"}
    assert arr["skeleton_xyz"].shape == (sum(3 + b for b in br), 3) and arr["skeleton_radii"].shape == (arr["skeleton_xyz"].shape[0], 1)
    assert arr["branch_num_elements"].tolist() == [3 + b for b in br] and arr["branch_id"].tolist() == list(br)
    save_skeleton(sk, tmp_path / "s.npz")
    back = load_skeleton(tmp_path / "s.npz")
    assert sorted(back.branches) == sorted(br)
    for b, ref in br.items():
        got = back.branches[b]
        assert got.parent_id == ref.parent_id and torch.equal(got.xyz, ref.xyz) and torch.equal(got.radii, ref.radii)
    empty = skeleton_arrays(TreeSkeleton(0, {}))
    assert empty["skeleton_xyz"].shape == (0, 3) and empty["branch_id"].shape == (0,)


def test_load_cloud_pin_memory_flag_is_harmless_without_cuda(tmp_path):
    """load_cloud(pin_memory=True) returns the same cloud (pinned only where a CUDA runtime exists); to_device accepts non_blocking."""
    import numpy as np
    import torch
    from smart_tree_b200.util.file import load_cloud
    rng = np.random.default_rng(0)
    xyz = rng.standard_normal((100, 3)).astype(np.float32)
    rgb = rng.uniform(0, 1, (100, 3)).astype(np.float32)
    p = tmp_path / "c.npz"
    np.savez(p, xyz=xyz, rgb=rgb)
    a, b = load_cloud(p), load_cloud(p, pin_memory=True)
    assert torch.equal(a.xyz, b.xyz) and torch.equal(a.rgb, b.rgb)
    assert b.xyz.is_pinned() == torch.cuda.is_available()
    c = b.to_device(torch.device("cpu"), non_blocking=True)
    assert torch.equal(c.xyz, a.xyz)
