"""spconv.pytorch-compatible modules on libst_b200 (inference only; fp32).

Semantics follow SURVEY.md Appendix B1-B5.  Differences from real spconv, all deliberate:
  * by default the grid is unbounded -- `spatial_shape` is carried but not used for clipping (SURVEY C-3); set
    `smart_tree_b200.spconv.pytorch.STRICT_SPATIAL_SHAPE = True` (or ST_STRICT_SPCONV_BOUNDS=1) to reproduce spconv's
    bound checks against the declared shape (neighbour locations >= shape are invisible, strided outputs >= out_shape
    are not created).  Coordinates must satisfy 0 <= z,y,x <= 65533, 0 <= batch < 32768 (checked, raises otherwise)
  * the sub-manifold neighbour map is built once per index set and shared by every SubMConv3d
    that sees the same `indices` tensor (spconv rebuilds it when no indice_key is given, C-4)
  * the output rows of a strided SparseConv3d are sorted by (b,z,y,x) (spconv: hash order)"""
from __future__ import annotations

import enum
import weakref
from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import ops
from ..engine import LevelIndex, _conv_w


import os as _os

STRICT_SPATIAL_SHAPE = bool(int(_os.environ.get("ST_STRICT_SPCONV_BOUNDS", "0")))


def _shape_tensor(spatial_shape, device):
    if spatial_shape is None or not STRICT_SPATIAL_SHAPE:
        return None
    if torch.is_tensor(spatial_shape):
        return spatial_shape.to(device=device, dtype=torch.int32).contiguous()
    return torch.tensor([int(v) for v in spatial_shape], dtype=torch.int32, device=device)


class ConvAlgo(enum.Enum):
    Native = 0
    MaskImplicitGemm = 1
    MaskSplitImplicitGemm = 2


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None, indice_dict=None,
                 benchmark=False):
        self.features = features
        self.indices = indices
        self.spatial_shape = spatial_shape
        self.batch_size = batch_size
        self.indice_dict = indice_dict if indice_dict is not None else {}

    def replace_feature(self, feature):
        return SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, indice_dict=self.indice_dict)

    def level(self) -> LevelIndex:
        """Coordinate table + sub-manifold map of this tensor's index set.  Cached per `indices`
        tensor OBJECT (weakly referenced): ResBlock / UBlock in the reference build new
        SparseConvTensors from (features, indices, ...) (model_blocks.py:149-151,227-232) but pass the
        same indices tensor along, so every conv of a level shares one map."""
        return _level_of(self.indices, self.spatial_shape)

    def dense(self):
        raise NotImplementedError("dense() is not part of the hot path")


_LEVEL_CACHE: Dict[int, tuple] = {}


def _register_level(indices, lv):
    if len(_LEVEL_CACHE) > 32:
        for k in [k for k, (r, _) in _LEVEL_CACHE.items() if r() is None]:
            del _LEVEL_CACHE[k]
        if len(_LEVEL_CACHE) > 32:
            _LEVEL_CACHE.clear()
    _LEVEL_CACHE[id(indices)] = (weakref.ref(indices), lv)


def _level_of(indices, spatial_shape=None) -> LevelIndex:
    ent = _LEVEL_CACHE.get(id(indices))
    if ent is not None and ent[0]() is indices and ent[2:] == ():
        return ent[1]
    if not indices.is_cuda:
        raise RuntimeError("smart_tree_b200.spconv runs on CUDA tensors only (no CPU fallback)")
    lv = LevelIndex(indices.int().contiguous(), _shape_tensor(spatial_shape, indices.device))
    lv.table.check()
    lv.child = None
    _register_level(indices, lv)
    return lv


class SparseModule(nn.Module):
    pass


def _is_sparse_mod(m):
    return isinstance(m, SparseModule)


class SparseSequential(SparseModule):
    def __init__(self, *args):
        super().__init__()
        for i, m in enumerate(args):
            self.add_module(str(i), m)

    def add(self, module, name=None):
        self.add_module(name if name is not None else str(len(self._modules)), module)

    def __len__(self):
        return len(self._modules)

    def __getitem__(self, i):
        return list(self._modules.values())[i]

    def forward(self, x):
        for m in self._modules.values():
            if _is_sparse_mod(m):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.indices.shape[0] != 0:
                    x = x.replace_feature(m(x.features))
            else:
                x = m(x)
        return x


class _SparseConvBase(SparseModule):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, subm=False, inverse=False, **_):
        super().__init__()
        ks = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
        if ks not in (1, 3) or dilation != 1 or groups != 1:
            raise NotImplementedError("libst_b200 supports 1x1x1 and 3x3x3 kernels, dilation 1, groups 1")
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, [ks] * 3
        self.stride, self.padding, self.indice_key, self.algo = stride, padding, indice_key, algo
        self.subm, self.inverse = subm, inverse
        self.conv1x1 = ks == 1
        self.weight = nn.Parameter(torch.empty(out_channels, ks, ks, ks, in_channels))      # spconv KRSC layout
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        nn.init.kaiming_uniform_(self.weight.view(out_channels, -1), a=5 ** 0.5)
        self._wt = None

    def _w(self):
        v = (self.weight._version, self.weight.data_ptr(), self.weight.device)
        if self._wt is None or self._wt[0] != v:
            self._wt = (v, _conv_w(self.weight.detach()).to(self.weight.device))
        return self._wt[1]

    def _run(self, feats, nbr, n_out):
        if not feats.is_cuda:
            raise RuntimeError("smart_tree_b200.spconv runs on CUDA tensors only (no CPU fallback)")
        return ops.conv_gather(feats.float(), nbr, self._w(), n_out, shift=self.bias)


class SubMConv3d(_SparseConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, **kw):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, indice_key, algo, subm=True)

    def forward(self, x: SparseConvTensor):
        n = x.features.shape[0]
        if self.conv1x1:
            return x.replace_feature(self._run(x.features, None, n))
        return x.replace_feature(self._run(x.features, x.level().nbr, n))


class SparseConv3d(_SparseConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, **kw):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, indice_key, algo)
        if not (self.kernel_size[0] == 3 and stride == 2 and padding == 1):
            raise NotImplementedError("only SparseConv3d(kernel_size=3, stride=2, padding=1) is on the hot path")

    def forward(self, x: SparseConvTensor):
        lv = x.level()
        if lv.down is None:
            out_shape = None
            if lv.spatial_shape is not None:
                out_shape = (torch.div(lv.spatial_shape - 1, 2, rounding_mode="floor") + 1).int().contiguous()
            oc = ops.strided_coords(lv.coords, out_shape=out_shape)
            lv.child = LevelIndex(oc, out_shape)
            lv.child.child = None
            lv.down, lv.up = ops.strided_maps(lv.coords, oc, lv.child.table)
        child = lv.child
        feats = self._run(x.features, lv.down, child.n)
        shape = [(int(s) - 1) // 2 + 1 for s in (x.spatial_shape.tolist() if torch.is_tensor(x.spatial_shape) else x.spatial_shape)]
        out = SparseConvTensor(feats, child.coords, shape, x.batch_size, indice_dict=x.indice_dict)
        _register_level(child.coords, child)
        if self.indice_key is not None:
            if self.indice_key in x.indice_dict and x.indice_dict[self.indice_key][0] is not lv:
                raise AssertionError(f"indice_key {self.indice_key} reused with a different index set")
            x.indice_dict[self.indice_key] = (lv, x.indices, x.spatial_shape)
        return out


class SparseInverseConv3d(_SparseConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key=None, bias=True, algo=None, **kw):
        super().__init__(in_channels, out_channels, kernel_size, 1, 0, 1, 1, bias, indice_key, algo, inverse=True)

    def forward(self, x: SparseConvTensor):
        if self.indice_key not in x.indice_dict:
            raise AssertionError(f"SparseInverseConv3d: indice_key {self.indice_key} was never registered by a SparseConv3d")
        lv, indices, shape = x.indice_dict[self.indice_key]
        feats = self._run(x.features, lv.up, lv.n)
        return SparseConvTensor(feats, indices, shape, x.batch_size, indice_dict=x.indice_dict)
