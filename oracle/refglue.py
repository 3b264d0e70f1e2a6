"""Run the REFERENCE's own Python glue (imported from /root/reference, unmodified) on the oracle's
third-party stand-ins and freeze the results as golden fixtures under tests/golden/.

    python -m oracle.refglue            # regenerates tests/golden/refglue_*.npz   (needs /root/reference)

TEST INFRASTRUCTURE.  What this pins: every line of the reference's model / dataset / skeleton glue
(module wiring, BN/ReLU/cat order, block tiling and masks, exp*direction, argmax, outlier rule, edge rule,
component loop, sample_tree incl. all its quirks).  What it cannot pin: the arithmetic inside spconv / FRNN /
cugraph, which are not available; those calls land in oracle/fake_thirdparty.py."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
WEIGHTS = os.path.join(ROOT, "smart-tree_b200", "model", "weights")
CPU = torch.device("cpu")


class _TorchOnCpu:
    """`torch` as seen by reference modules that hard-code torch.device("cuda") (skeleton/path.py:74,78)."""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(*a, **k):
        return CPU


def _ieee_sqrt(self):
    """torch's CPU float32 sqrt is NOT correctly rounded for ~0.05 % of inputs (vectorised approximation),
    whereas the reference runs on CUDA, whose sqrtf is IEEE round-to-nearest (nvcc default -prec-sqrt=true).
    Generating the fixtures with the CPU quirk would bake a CPU artefact into them, so Tensor.sqrt is routed
    through numpy (correctly rounded) while the reference code runs here."""
    if self.dtype == torch.float32 and self.device.type == "cpu":
        with np.errstate(invalid="ignore"):
            return torch.from_numpy(np.sqrt(self.detach().numpy()))
    return _TORCH_SQRT(self)


_TORCH_SQRT = torch.Tensor.sqrt


def _setup():
    from . import fake_thirdparty
    fake_thirdparty.install()
    torch.Tensor.sqrt = _ieee_sqrt
    sys.path.insert(0, os.path.join(ROOT, "smart-tree_b200"))
    import smart_tree.skeleton.path as path_mod
    path_mod.torch = _TorchOnCpu()
    path_mod.tqdm = lambda *a, **k: _NoBar()
    return path_mod


class _NoBar:
    n = 0

    def update(self, n=0):
        pass

    def refresh(self):
        pass


def reference_model(weights_name):
    """The checkpoint architecture, assembled from the reference's own classes: HEAD's
    Smart_Tree with its MLP heads swapped for the SparseFC heads the checkpoints were trained with, BN eps 1e-4."""
    import functools

    import smart_tree.model.model_blocks as mb
    from smart_tree.model.model import Smart_Tree
    m = Smart_Tree(3, [8, 16, 32, 64], [8, 8, 4, 1], [8, 8, 4, 3], [8, 8, 4, 2])
    m.radius_head = mb.SparseFC([8, 8, 4, 1], torch.nn.BatchNorm1d, torch.nn.ReLU)
    m.direction_head = mb.SparseFC([8, 8, 4, 3], torch.nn.BatchNorm1d, torch.nn.ReLU)
    m.class_head = mb.SparseFC([8, 8, 4, 2], torch.nn.BatchNorm1d, torch.nn.ReLU)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.eps = 1e-4
    sd = torch.load(os.path.join(WEIGHTS, f"{weights_name}_model_weights.pt"), map_location="cpu", weights_only=True)
    m.load_state_dict(sd)
    return m.eval(), sd


def gen_inference(weights_name="noble-elevator-58", n_points=12000, voxel=0.02, seed=0):
    """ModelInference.forward (model_inference.py:49-100) incl. SingleTreeInference / DataLoader / collate."""
    import synth
    import smart_tree.model.model_inference as mi
    from smart_tree.data_types.cloud import Cloud
    from smart_tree.dataset.augmentations import CentreCloud
    model, _ = reference_model(weights_name)
    mi.load_model = lambda *a, **k: model
    tr = synth.make_tree(seed, n_points)
    # two trees 6 m apart so that several 4 m blocks (and their buffers) are exercised
    xyz = np.concatenate([tr.xyz, tr.xyz[::3] + np.float32([3.0, 0, 2.0])]).astype(np.float32)
    cloud = CentreCloud()(Cloud(xyz=torch.from_numpy(xyz), rgb=torch.zeros(len(xyz), 3)))
    inf = mi.ModelInference(None, None, voxel, 4, 0.4, num_workers=0, batch_size=4, device=CPU)
    torch.manual_seed(0)
    with torch.no_grad():
        lc = inf.forward(cloud)
    np.savez_compressed(os.path.join(GOLDEN, f"refglue_inference_{weights_name}.npz"), in_xyz=xyz, voxel=np.float32(voxel),
                        xyz=lc.xyz.numpy(), rgb=lc.rgb.numpy(), medial_vector=lc.medial_vector.numpy(),
                        class_l=lc.class_l.numpy().reshape(-1))
    return lc


def gen_model_forward(weights_name, n_points=6000, voxel=0.03, random_weights=False):
    """Smart_Tree.forward (model.py:77-87) on one voxelised cloud; optionally random weights / features."""
    import synth
    from oracle import pipeline_ref as P
    from smart_tree.model.sparse import sparse_from_batch
    model, sd = reference_model(weights_name)
    tag = weights_name
    if random_weights:
        g = torch.Generator().manual_seed(11)
        new = {}
        for k, v in sd.items():
            if k.endswith("num_batches_tracked"):
                new[k] = v
            elif k.endswith("running_var"):
                new[k] = torch.rand(v.shape, generator=g) + 0.5
            elif v.dim() == 1:
                new[k] = torch.randn(v.shape, generator=g) * 0.1 + (1.0 if k.endswith("weight") else 0.0)
            else:
                new[k] = torch.randn(v.shape, generator=g) * (2.0 / v[0].numel()) ** 0.5
        model.load_state_dict(new)
        sd, tag = new, "random"
    tr = synth.make_tree(1, n_points)
    xyz = P.centre_cloud(tr.xyz)
    vox, zyx, _, _ = P.point_to_voxel_exact(np.concatenate([xyz, tr.rgb], 1), voxel, xyz.min(0), xyz.max(0))
    coords = np.concatenate([np.zeros((len(zyx), 1), np.int32), zyx], 1)
    feats = vox[:, :3].copy()
    if random_weights:
        feats = np.random.default_rng(5).standard_normal(feats.shape).astype(np.float32)
    st = sparse_from_batch(torch.from_numpy(feats), torch.from_numpy(coords), device=CPU)
    with torch.no_grad():
        out = model(st)
    np.savez_compressed(os.path.join(GOLDEN, f"refglue_forward_{tag}.npz"), feats=feats, coords=coords,
                        radius=out["radius"].numpy(), direction=out["direction"].numpy(), class_l=out["class_l"].numpy(),
                        **({"sd_" + k: v.numpy() for k, v in sd.items() if not k.endswith("num_batches_tracked")} if random_weights else {}))


def gen_skeleton(n_points=16000, voxel=0.03, seed=0):
    """outlier_removal, nn_graph, Skeletonizer.forward (skeletonize.py:31-95) incl. sample_tree (path.py)."""
    import synth
    from oracle import pipeline_ref as P
    from smart_tree.data_types.cloud import Cloud
    from smart_tree.skeleton.filter import outlier_removal
    from smart_tree.skeleton.graph import nn_graph
    from smart_tree.skeleton.skeletonize import Skeletonizer
    import smart_tree.skeleton.skeletonize as sk_mod
    import smart_tree.data_types.graph as g_mod
    sk_mod.tqdm = lambda it, **k: it
    g_mod.tqdm = lambda it, **k: it
    tr = synth.make_tree(seed, n_points)
    vox, _, _, _ = P.point_to_voxel_exact(np.concatenate([tr.xyz, tr.medial_vector], 1), voxel, tr.xyz.min(0), tr.xyz.max(0))
    rng = np.random.default_rng(seed)
    xyz = np.concatenate([vox[:, :3], rng.uniform(-3, 3, (40, 3)).astype(np.float32)])
    mv = np.concatenate([vox[:, 3:6], rng.normal(0, 0.02, (40, 3)).astype(np.float32)])
    cloud = Cloud(xyz=torch.from_numpy(xyz), medial_vector=torch.from_numpy(mv))
    keep = outlier_removal(cloud.medial_pts, cloud.radius.unsqueeze(1), nb_points=8)
    fc = cloud.filter(keep)
    graph = nn_graph(fc.medial_pts, fc.radius.clamp(min=0.02), K=16)
    skel = Skeletonizer(K=16, min_connection_length=0.02, minimum_graph_vertices=32, device=CPU).forward(cloud)
    out = dict(xyz=xyz, medial_vector=mv, keep=keep.numpy(), edges=graph.edges.numpy(), edge_weights=graph.edge_weights.numpy(),
               n_skeletons=np.int64(len(skel.skeletons)))
    for si, s in enumerate(skel.skeletons):
        bs = list(s.branches.values())
        out[f"s{si}_branch_id"] = np.array([b._id for b in bs])
        out[f"s{si}_parent_id"] = np.array([b.parent_id for b in bs])
        out[f"s{si}_len"] = np.array([len(b) for b in bs])
        out[f"s{si}_xyz"] = np.concatenate([b.xyz.numpy() for b in bs]) if bs else np.zeros((0, 3), np.float32)
        out[f"s{si}_radii"] = np.concatenate([b.radii.numpy().reshape(-1) for b in bs]) if bs else np.zeros(0, np.float32)
    np.savez_compressed(os.path.join(GOLDEN, "refglue_skeleton.npz"), **out)
    return skel


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    _setup()
    gen_model_forward("noble-elevator-58")
    gen_model_forward("peach-forest-65")
    gen_model_forward("noble-elevator-58", random_weights=True)
    gen_inference("noble-elevator-58")
    sk = gen_skeleton()
    print("golden fixtures written to", GOLDEN, [len(s.branches) for s in sk.skeletons])


if __name__ == "__main__":
    main()
