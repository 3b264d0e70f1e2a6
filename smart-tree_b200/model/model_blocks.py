"""Network blocks with the reference's module tree and state_dict keys
(/root/reference/smart_tree/model/model_blocks.py:8-320), built on the libst_b200 spconv stand-in.
Used for API compatibility (load_state_dict, module-by-module execution); the production path is
the fused engine (engine.py), which these modules feed with their parameters."""
import torch
import torch.nn as nn

from .. import spconv
from ..spconv import SparseModule


def _seq(*mods):
    return spconv.SparseSequential(*mods)


class SubMConvBlock(SparseModule):
    def __init__(self, input_channels, output_channels, kernel_size, norm_fn, activation_fn, stride=1, padding=1,
                 algo=spconv.ConvAlgo.Native, bias=False):
        super().__init__()
        self.sequence = _seq(spconv.SubMConv3d(input_channels, output_channels, kernel_size, stride=stride, padding=padding,
                                               bias=bias, algo=algo), norm_fn(output_channels), activation_fn())

    def forward(self, x):
        return self.sequence(x)


class EncoderBlock(SparseModule):
    def __init__(self, input_channels, output_channels, kernel_size, norm_fn, activation_fn, stride=2, padding=1, key=None,
                 algo=spconv.ConvAlgo.Native, bias=False):
        super().__init__()
        self.sequence = _seq(spconv.SparseConv3d(input_channels, output_channels, kernel_size, stride=stride, indice_key=key,
                                                 algo=algo, bias=bias, padding=padding), norm_fn(output_channels), activation_fn())

    def forward(self, x):
        return self.sequence(x)


class DecoderBlock(SparseModule):
    def __init__(self, input_channels, output_channels, kernel_size, norm_fn, activation_fn, key=None,
                 algo=spconv.ConvAlgo.Native, bias=False):
        super().__init__()
        self.sequence = _seq(spconv.SparseInverseConv3d(input_channels, output_channels, kernel_size, indice_key=key, algo=algo,
                                                        bias=bias), norm_fn(output_channels), activation_fn())

    def forward(self, x):
        return self.sequence(x)


class ResBlock(nn.Module):
    def __init__(self, input_channels, output_channels, kernel_size, norm_fn, activation_fn, algo=spconv.ConvAlgo.Native, bias=False):
        super().__init__()
        if input_channels == output_channels:
            self.identity = _seq(nn.Identity())
        else:
            self.identity = _seq(spconv.SubMConv3d(input_channels, output_channels, kernel_size=1, padding=1, bias=False, algo=algo))
        self.sequence = _seq(
            spconv.SubMConv3d(input_channels, output_channels, kernel_size, bias=False, algo=algo), norm_fn(output_channels),
            activation_fn(),
            spconv.SubMConv3d(output_channels, output_channels, kernel_size, bias=False, algo=algo), norm_fn(output_channels))
        self.activation_fn = _seq(activation_fn())

    def forward(self, x):
        ident = spconv.SparseConvTensor(x.features, x.indices, x.spatial_shape, x.batch_size)
        out = self.sequence(x)
        out = out.replace_feature(out.features + self.identity(ident).features)
        return self.activation_fn(out)


class UBlock(nn.Module):
    def __init__(self, n_planes, norm_fn, activation_fn, kernel_size=3, key_id=1, algo=spconv.ConvAlgo.Native, bias=False):
        super().__init__()
        self.n_planes = list(n_planes)
        p = self.n_planes
        self.Head = ResBlock(p[0], p[0], kernel_size, norm_fn, activation_fn, algo=algo, bias=bias)
        if len(p) > 1:
            self.Encode = EncoderBlock(p[0], p[1], kernel_size, norm_fn, activation_fn, stride=2, key=key_id, algo=algo, bias=bias)
            self.U = UBlock(p[1:], norm_fn, activation_fn, kernel_size, key_id + 1, algo, bias)
            self.Decode = DecoderBlock(p[1], p[0], kernel_size, norm_fn, activation_fn, key=key_id, algo=algo, bias=bias)
            self.Tail = ResBlock(p[0] * 2, p[0], kernel_size, norm_fn, activation_fn, algo=algo, bias=bias)

    def forward(self, x):
        out = self.Head(x)
        skip = spconv.SparseConvTensor(out.features, out.indices, out.spatial_shape, out.batch_size)
        if len(self.n_planes) > 1:
            out = self.Decode(self.U(self.Encode(out)))
            out = out.replace_feature(torch.cat((skip.features, out.features), dim=1))
            out = self.Tail(out)
        return out


class SparseFC(nn.Module):
    """Heads of the shipped checkpoints: 1x1 sub-manifold convs without bias (model_blocks.py:246-285)."""

    def __init__(self, n_planes, norm_fn, activation_fn=None, kernel_size=1, algo=spconv.ConvAlgo.Native, bias=False):
        super().__init__()
        self.sequence = spconv.SparseSequential()
        for i in range(len(n_planes) - 2):
            self.sequence.add(spconv.SubMConv3d(n_planes[i], n_planes[i + 1], kernel_size=kernel_size, bias=False, algo=algo, padding=0))
            self.sequence.add(norm_fn(n_planes[i + 1]))
            self.sequence.add(activation_fn())
        self.sequence.add(spconv.SubMConv3d(n_planes[-2], n_planes[-1], kernel_size=kernel_size, bias=False, algo=algo, padding=0))

    def forward(self, x):
        return self.sequence(x)


class MLP(nn.Module):
    """Heads of the reference's HEAD code: nn.Linear stacks (model_blocks.py:288-320)."""

    def __init__(self, n_planes, norm_fn, activation_fn=None, bias=False):
        super().__init__()
        self.sequence = spconv.SparseSequential()
        for i in range(len(n_planes) - 2):
            self.sequence.add(nn.Linear(n_planes[i], n_planes[i + 1], bias=bias))
            self.sequence.add(norm_fn(n_planes[i + 1]))
            self.sequence.add(activation_fn())
        self.sequence.add(nn.Linear(n_planes[-2], n_planes[-1], bias=bias))

    def forward(self, x):
        return self.sequence(x)
