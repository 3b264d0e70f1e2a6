"""CPU stand-ins for the third-party packages the reference imports (spconv, frnn, cugraph, cudf,
cupy, open3d, torchtyping, cmapy), backed by the oracle restatements -- TEST INFRASTRUCTURE.

Purpose: run the reference's OWN Python (model.py, model_blocks.py, dataset.py, sparse.py,
model_inference.py, filter.py, graph.py, path.py, skeletonize.py, data_types/*) unmodified in this
container and freeze its outputs as golden fixtures (oracle/refglue.py -> tests/golden/).  This pins
every line of reference glue; the arithmetic inside the faked packages is the oracle's restatement
(SURVEY Appendix B), so parity at the third-party boundary itself stays unpinned.
"""
from __future__ import annotations

import enum
import sys
import types

import numpy as np
import pandas as pd
import torch
import torch.nn as nn

from . import skeleton_ref as S
from . import unet_ref as U


# ----------------------------------------------------------------------------- spconv
class ConvAlgo(enum.Enum):
    Native = 0


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, indice_dict=None, **_):
        self.features, self.indices, self.spatial_shape, self.batch_size = features, indices, spatial_shape, batch_size
        self.indice_dict = indice_dict if indice_dict is not None else {}

    def replace_feature(self, f):
        return SparseConvTensor(f, self.indices, self.spatial_shape, self.batch_size, self.indice_dict)


class SparseModule(nn.Module):
    pass


class SparseSequential(SparseModule):
    def __init__(self, *mods):
        super().__init__()
        for i, m in enumerate(mods):
            self.add_module(str(i), m)

    def add(self, m, name=None):
        self.add_module(name or str(len(self._modules)), m)

    def forward(self, x):
        for m in self._modules.values():
            if isinstance(m, SparseModule):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.indices.shape[0]:
                    x = x.replace_feature(m(x.features))
            else:
                x = m(x)
        return x


_MAPS = {}


def _np(t):
    return t.detach().cpu().numpy()


def _subm(indices):
    key = (indices.data_ptr(), indices.shape[0])
    if key not in _MAPS:
        _MAPS.clear()
        _MAPS[key] = U.subm_map(_np(indices))
    return _MAPS[key]


class _Conv(SparseModule):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, **_):
        super().__init__()
        k = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
        self.k, self.indice_key = k, indice_key
        self.weight = nn.Parameter(torch.randn(out_channels, k, k, k, in_channels) * 0.1)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def _apply_conv(self, feats, nbr, n_out):
        w = _np(self.weight)
        out = U.linear(_np(feats), w) if self.k == 1 else U.gather_conv(_np(feats), w, nbr, n_out)
        out = torch.from_numpy(np.ascontiguousarray(out, np.float32))
        return out + self.bias if self.bias is not None else out


class SubMConv3d(_Conv):
    def forward(self, x):
        n = x.features.shape[0]
        nbr = None if self.k == 1 else _subm(x.indices)
        return x.replace_feature(self._apply_conv(x.features, nbr, n))


class SparseConv3d(_Conv):
    def forward(self, x):
        oc, down, up = U.strided_maps(_np(x.indices))
        feats = self._apply_conv(x.features, down, len(oc))
        x.indice_dict[self.indice_key] = (up, x.indices, x.spatial_shape)
        return SparseConvTensor(feats, torch.from_numpy(oc), x.spatial_shape, x.batch_size, x.indice_dict)


class SparseInverseConv3d(_Conv):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key=None, bias=True, algo=None, **kw):
        super().__init__(in_channels, out_channels, kernel_size, bias=bias, indice_key=indice_key)

    def forward(self, x):
        up, indices, shape = x.indice_dict[self.indice_key]
        return SparseConvTensor(self._apply_conv(x.features, up, indices.shape[0]), indices, shape, x.batch_size, x.indice_dict)


class PointToVoxel:
    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_voxels, max_num_points_per_voxel, device=None):
        assert max_num_points_per_voxel == 1
        self.vs = float(vsize_xyz[0])
        self.rng = [float(v) for v in coors_range_xyz]

    def generate_voxel_with_id(self, pc):
        from .pipeline_ref import point_to_voxel_exact
        vox, zyx, pcid, _ = point_to_voxel_exact(_np(pc), self.vs, np.float32(self.rng[:3]), np.float32(self.rng[3:]))
        return (torch.from_numpy(vox).unsqueeze(1), torch.from_numpy(zyx.copy()), torch.ones(len(vox), dtype=torch.int32),
                torch.from_numpy(pcid))


# ----------------------------------------------------------------------------- frnn
def frnn_grid_points(p1, p2, l1=None, l2=None, K=-1, r=-1.0, grid=None, return_nn=False, return_sorted=True, radius_cell_ratio=2.0):
    idx, d2 = S.knn(_np(p1[0]), _np(p2[0]), int(K), float(r))
    return torch.from_numpy(d2)[None], torch.from_numpy(idx)[None], None, None


# ----------------------------------------------------------------------------- cugraph / cudf / cupy
class _Graph:
    def __init__(self, directed=False):
        self.directed = directed

    def from_cudf_edgelist(self, df, source="source", destination="destination", edge_attr=None, renumber=False):
        src, dst = np.asarray(df[source]).astype(np.int64), np.asarray(df[destination]).astype(np.int64)
        self.w = np.asarray(df[edge_attr]).astype(np.float32) if edge_attr else np.ones(len(src), np.float32)
        self.e = np.stack([src, dst], 1)
        self.n = int(self.e.max()) + 1 if len(self.e) else 0
        # undirected, de-duplicated view (lower id first), as cugraph stores it
        a, b = np.minimum(src, dst), np.maximum(src, dst)
        key = a * max(self.n, 1) + b
        _, first = np.unique(key, return_index=True)
        self.ue, self.uw = np.stack([a[first], b[first]], 1), self.w[first]
        self._nodes = np.unique(self.ue)

    def nodes(self):
        return pd.Series(self._nodes)

    def edges(self):
        return pd.DataFrame({"src": self.ue[:, 0], "dst": self.ue[:, 1]})


def _connected_components(g):
    from scipy import sparse
    from scipy.sparse import csgraph
    a = sparse.coo_matrix((np.ones(len(g.ue)), (g.ue[:, 0], g.ue[:, 1])), shape=(g.n, g.n))
    _, lab = csgraph.connected_components(a, directed=False)
    # label = smallest vertex id of the component; rows ordered so that `unique()` yields the oracle's
    # component order (size desc, then smallest vertex) -- the reference's order is implementation-defined
    first = np.full(lab.max() + 1, g.n, np.int64)
    np.minimum.at(first, lab, np.arange(g.n))
    label = first[lab]
    size = np.bincount(lab)[lab]
    order = np.lexsort((np.arange(g.n), label, -size))
    return FakeFrame({"vertex": np.arange(g.n)[order], "labels": label[order]})


def _subgraph(g, vertices):
    v = np.asarray(vertices)
    keep = np.isin(g.ue[:, 0], v) & np.isin(g.ue[:, 1], v)
    s = _Graph()
    s.n, s.ue, s.uw = g.n, g.ue[keep], g.uw[keep]
    s.e, s.w = s.ue, s.uw
    s._nodes = np.unique(s.ue)
    return s


def _to_pandas_edgelist(g):
    return pd.DataFrame({"src": g.ue[:, 0], "dst": g.ue[:, 1], "weights": g.uw})


def _sssp(g, source):
    n = g.n
    pred, dist = S.sssp(n, g.ue, g.uw, int(source))
    # pred_graph (shortest_path.py:46-55) feeds a TREE: cugraph's distances there are root->leaf running sums
    return FakeFrame({"vertex": np.arange(n), "distance": dist, "predecessor": pred})


class FakeSeries(pd.Series):
    """cudf.Series flavour: unique() returns a Series, to_pandas() exists."""

    @property
    def _constructor(self):
        return FakeSeries

    @property
    def _constructor_expanddim(self):
        return FakeFrame

    def unique(self):
        return FakeSeries(pd.Series.unique(self))

    def to_pandas(self):
        return pd.Series(self)


class FakeFrame(pd.DataFrame):
    @property
    def _constructor(self):
        return FakeFrame

    @property
    def _constructor_sliced(self):
        return FakeSeries


class _Cupy(types.ModuleType):
    @staticmethod
    def asarray(x):
        return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)

    unique = staticmethod(np.unique)


class _TensorType:
    def __class_getitem__(cls, item):
        return torch.Tensor


class _Anything(types.ModuleType):
    """open3d & co: importable, attribute access yields further stubs; never called on the hot path."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Anything(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        raise RuntimeError(f"{self.__name__} is a stub")


def install(reference_root="/root/reference"):
    """Register the stand-ins and put the reference on sys.path.  Idempotent."""
    if "spconv" in sys.modules and getattr(sys.modules["spconv"], "_oracle_fake", False):
        return

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    pt = mod("spconv.pytorch", ConvAlgo=ConvAlgo, SparseConvTensor=SparseConvTensor, SparseModule=SparseModule,
             SparseSequential=SparseSequential, SubMConv3d=SubMConv3d, SparseConv3d=SparseConv3d,
             SparseInverseConv3d=SparseInverseConv3d)
    ut = mod("spconv.pytorch.utils", PointToVoxel=PointToVoxel)
    pt.utils = ut
    sp = mod("spconv", pytorch=pt, ConvAlgo=ConvAlgo, _oracle_fake=True)
    sp.__path__ = []
    pt.__path__ = []
    mod("frnn", frnn_grid_points=frnn_grid_points)
    mod("cudf", DataFrame=FakeFrame)
    cp = _Cupy("cupy")
    sys.modules["cupy"] = cp
    mod("cugraph", Graph=_Graph, connected_components=_connected_components, subgraph=_subgraph, sssp=_sssp,
        to_pandas_edgelist=_to_pandas_edgelist)
    mod("torchtyping", TensorType=_TensorType, patch_typeguard=lambda: None)
    mod("typeguard", typechecked=lambda f=None, **k: f if f is not None else (lambda g: g))
    for name in ("open3d", "open3d.visualization", "open3d.visualization.rendering", "cmapy"):
        sys.modules[name] = _Anything(name)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
