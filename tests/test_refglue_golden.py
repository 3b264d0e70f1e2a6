"""Golden fixtures produced by the REFERENCE's own Python glue running on the oracle's third-party stand-ins
(oracle/refglue.py, committed with its output under tests/golden/).  CPU tests: the oracle's end-to-end
restatement reproduces what the reference code computes.  GPU tests: so does the CUDA path."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, WEIGHTS
from oracle import pipeline_ref as P
from oracle import skeleton_ref as S
from oracle import unet_ref as U


def _g(name):
    return np.load(os.path.join(GOLDEN, name))


def _params(tag, g):
    if tag == "random":
        return {k[3:]: g[k] for k in g.files if k.startswith("sd_")}
    return U.to_numpy_params(torch.load(os.path.join(WEIGHTS, f"{tag}_model_weights.pt"), map_location="cpu", weights_only=True))


def _sorted_rows(xyz):
    return np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))


@pytest.mark.parametrize("tag", ["noble-elevator-58", "peach-forest-65", "random"])
def test_oracle_forward_reproduces_reference_model_code(tag):
    g = _g(f"refglue_forward_{tag}.npz")
    out = U.forward(_params(tag, g), g["feats"], g["coords"])
    for k in ("radius", "direction", "class_l"):
        scale = np.abs(g[k]).max()
        assert np.abs(out[k] - g[k]).max() <= 2e-5 * scale, k


def test_oracle_inference_reproduces_reference_glue():
    g = _g("refglue_inference_noble-elevator-58.npz")
    sd = torch.load(os.path.join(WEIGHTS, "noble-elevator-58_model_weights.pt"), map_location="cpu", weights_only=True)
    xyz = P.centre_cloud(g["in_xyz"])
    lab = P.infer(U.to_numpy_params(sd), xyz, np.zeros_like(xyz), float(g["voxel"]), 4, 0.4)
    # the reference shuffles blocks (DataLoader quirk, Appendix C-1): compare as sets of voxels
    a, b = _sorted_rows(lab["xyz"]), _sorted_rows(g["xyz"])
    assert np.array_equal(lab["xyz"][a], g["xyz"][b])
    np.testing.assert_allclose(lab["medial_vector"][a], g["medial_vector"][b], rtol=0, atol=2e-6)
    assert np.array_equal(lab["class_l"][a], g["class_l"][b])


def _golden_skeletons(g):
    out = []
    for si in range(int(g["n_skeletons"])):
        ln = g[f"s{si}_len"]
        off = np.concatenate([[0], np.cumsum(ln)])
        out.append([(int(b), int(p), g[f"s{si}_xyz"][off[i]:off[i + 1]], g[f"s{si}_radii"][off[i]:off[i + 1]])
                    for i, (b, p) in enumerate(zip(g[f"s{si}_branch_id"], g[f"s{si}_parent_id"]))])
    return out


def test_oracle_skeleton_reproduces_reference_glue():
    g = _g("refglue_skeleton.npz")
    xyz, mv = g["xyz"], g["medial_vector"]
    med = xyz + mv
    rad = np.sqrt((mv[:, 0] * mv[:, 0] + mv[:, 1] * mv[:, 1]) + mv[:, 2] * mv[:, 2])
    keep = S.outlier_removal(med, rad, 8)
    assert np.array_equal(keep, g["keep"])
    e, w = S.nn_graph(med[keep], np.maximum(rad[keep], np.float32(0.02)), 16)
    assert np.array_equal(e, g["edges"]) and np.array_equal(w, g["edge_weights"])
    sk = S.skeletonize(xyz, mv, 16, 0.02, 32)
    gold = _golden_skeletons(g)
    assert len(sk) == len(gold)
    for s, gs in zip(sk, gold):
        assert len(s.branches) == len(gs)
        for b, (bid, par, bxyz, brad) in zip(s.branches, gs):
            assert (b.id, b.parent_id) == (bid, par)
            assert np.array_equal(b.xyz, bxyz)
            np.testing.assert_allclose(b.radii, brad, rtol=1e-6)


# ------------------------------------------------------------------ the CUDA path against the same fixtures
@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["noble-elevator-58", "peach-forest-65", "random"])
def test_cuda_forward_matches_reference_model_code(tag):
    from smart_tree_b200.engine import SmartTreeEngine
    g = _g(f"refglue_forward_{tag}.npz")
    sd = {k: torch.from_numpy(v) for k, v in _params(tag, g).items()}
    eng = SmartTreeEngine(sd, device="cuda")
    out = eng.forward(torch.from_numpy(g["feats"]).cuda(), torch.from_numpy(g["coords"]).cuda())
    for k in ("radius", "direction", "class_l"):
        e = np.abs(out[k].cpu().numpy() - g[k]) / max(np.abs(g[k]).max(), 1e-12)     # peach's class head is identically 0
        assert np.quantile(e, 0.999) < 1e-3 and e.max() < (1e-2 if k == "direction" else 1e-3), (k, e.max())


@pytest.mark.gpu
def test_cuda_inference_matches_reference_glue():
    from smart_tree_b200.data_types.cloud import Cloud
    from smart_tree_b200.dataset.augmentations import CentreCloud
    from smart_tree_b200.model.model_inference import ModelInference
    g = _g("refglue_inference_noble-elevator-58.npz")
    xyz = torch.from_numpy(g["in_xyz"]).cuda()
    mi = ModelInference(None, os.path.join(WEIGHTS, "noble-elevator-58_model_weights.pt"), float(g["voxel"]), 4, 0.4,
                        device=torch.device("cuda"))
    lc = mi.forward(CentreCloud()(Cloud(xyz=xyz, rgb=torch.zeros_like(xyz))))
    got_xyz = lc.xyz.cpu().numpy()
    a, b = _sorted_rows(got_xyz), _sorted_rows(g["xyz"])
    assert np.array_equal(got_xyz[a], g["xyz"][b])
    mvg, mvr = lc.medial_vector.cpu().numpy()[a], g["medial_vector"][b]
    assert np.abs(mvg - mvr).max() <= 1e-3 * np.abs(mvr).max()
    assert (lc.class_l.cpu().numpy().reshape(-1)[a] == g["class_l"][b]).mean() > 0.999


@pytest.mark.gpu
def test_cuda_skeleton_matches_reference_glue():
    from smart_tree_b200.data_types.cloud import Cloud
    from smart_tree_b200.skeleton.skeletonize import Skeletonizer
    g = _g("refglue_skeleton.npz")
    cloud = Cloud(xyz=torch.from_numpy(g["xyz"]).cuda(), medial_vector=torch.from_numpy(g["medial_vector"]).cuda())
    sk = Skeletonizer(16, 0.02, 32, device=torch.device("cuda")).forward(cloud)
    gold = _golden_skeletons(g)
    assert len(sk.skeletons) == len(gold)
    for s, gs in zip(sk.skeletons, gold):
        assert len(s.branches) == len(gs)
        for bid, par, bxyz, brad in gs:
            b = s.branches[bid]
            assert b.parent_id == par and np.array_equal(b.xyz.numpy(), bxyz)      # node coordinates bit-identical
            np.testing.assert_allclose(b.radii.numpy().reshape(-1), brad, rtol=1e-6)
