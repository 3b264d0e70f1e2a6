"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/st_b200.h declares (no compute calls here); the ctypes table matches the header."""
import os
import re

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "st_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(st_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from smart_tree_b200 import _lib
    from smart_tree_b200.build import build
    build()
    lib = _lib.load()
    names = _declared()
    assert len(names) > 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in st_b200.h but not exported by libst_b200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.st_version() >= 100
    assert lib.st_hash_capacity(1000) == 2048
    assert lib.st_conv_tc_weight_floats(27, 64, 64) == 54 * 2 * 64 * 32
    assert lib.st_conv_tc_weight_floats(27, 24, 40) == -1


def test_ops_refuse_cpu_tensors():
    import torch
    from smart_tree_b200 import _lib, ops
    with pytest.raises(_lib.StB200Error):
        ops.knn(torch.zeros(4, 3), torch.zeros(4, 3), 2, 0.1)
    with pytest.raises(_lib.StB200Error):
        ops.CoordTable(torch.zeros(4, 4, dtype=torch.int32))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from smart_tree_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.StB200Error, match="no CPU fallback"):
        _lib.load()


def test_host_modules_mirror_reference_names():
    import smart_tree_b200.compat as compat
    from smart_tree_b200.config import instantiate, load_config
    from smart_tree_b200.model.model import Smart_Tree
    cfg = load_config(overrides=["+path=foo.npz", "pipeline.skeletonizer.K=8"])
    assert cfg["path"] == "foo.npz" and cfg["pipeline"]["skeletonizer"]["K"] == 8
    # the reference's own _target_ paths resolve to the B200 classes
    sk = instantiate({"_target_": "smart_tree.skeleton.skeletonize.Skeletonizer", "K": 16, "min_connection_length": 0.02,
                      "minimum_graph_vertices": 32})
    assert type(sk).__module__ == "smart_tree_b200.skeleton.skeletonize" and sk.K == 16
    m = Smart_Tree(3, [8, 16, 32, 64], [8, 8, 4, 1], [8, 8, 4, 3], [8, 8, 4, 2])
    import torch
    sd = torch.load(os.path.join(ROOT, "smart-tree_b200", "model", "weights", "peach-forest-65_model_weights.pt"), map_location="cpu",
                    weights_only=True)
    assert m.load_state_dict(sd).missing_keys == []
    assert compat.install() is not None
    import spconv.pytorch as sp  # noqa: F401  (the stand-in, registered by compat.install)
    assert hasattr(sp, "SubMConv3d") and hasattr(sp, "SparseConvTensor")


def test_packed_branch_skeleton_is_a_lazy_branch_skeleton():
    """Host logic of the skeleton emit path: branches cut lazily from one packed array behave like BranchSkeleton."""
    import torch
    from smart_tree_b200.data_types.branch import BranchSkeleton, PackedBranchSkeleton
    nodes, rad = torch.randn(12, 3), torch.rand(12)
    b = PackedBranchSkeleton(4, 1, nodes, rad, 3, 5)
    assert isinstance(b, BranchSkeleton) and len(b) == 5 and (b._id, b.parent_id, b.child_id) == (4, 1, None)
    assert torch.equal(b.xyz, nodes[3:8]) and b.radii.shape == (5, 1) and torch.equal(b.radii[:, 0], rad[3:8])
    assert torch.equal(b.length, (nodes[4:8] - nodes[3:7]).norm(dim=1).sum()) and len(b.to_tubes()) == 4
    s = PackedBranchSkeleton(5, 4, nodes, rad, 0, 3, radii_1d=True)      # smoothed radii are 1-D (reference quirk)
    assert s.radii.shape == (3,)
    f = b.filter(torch.tensor([True, False, True, True, False]))
    assert type(f) is BranchSkeleton and len(f) == 3
    b.xyz, b.radii = nodes[:2], rad[:2].unsqueeze(1)                      # object-level post-processing may replace them
    assert len(b) == 2 and b.radii.shape == (2, 1)


def test_pyproject_console_script_mirrors_the_reference():
    """`run-smart-tree` (reference pyproject.toml:36) is declared and resolves to a callable; every sub-package listed."""
    import importlib
    import tomllib
    with open(os.path.join(ROOT, "pyproject.toml"), "rb") as f:
        cfg = tomllib.load(f)
    target = cfg["project"]["scripts"]["run-smart-tree"]
    assert target == "smart_tree_b200.cli:main"
    mod, fn = target.split(":")
    assert callable(getattr(importlib.import_module(mod), fn))
    pkgs = set(cfg["tool"]["setuptools"]["packages"])
    here = os.path.join(ROOT, "smart-tree_b200")
    found = {"smart_tree_b200"} | {"smart_tree_b200." + d for d in os.listdir(here)
                                   if os.path.exists(os.path.join(here, d, "__init__.py"))}
    assert pkgs == found
