"""Smart_Tree with the reference's constructor, module tree and forward contract
(/root/reference/smart_tree/model/model.py:10-87): forward(SparseConvTensor) -> {"radius",
"direction", "class_l"}.  Two execution modes:
  * fused (default in eval mode): the whole network runs on the fused engine compiled from this
    module's own state_dict (engine.SmartTreeEngine) -- one index build per level, BN/ReLU/residual/
    concat folded into the conv kernels, one heads kernel;
  * layerwise: module by module through the spconv stand-in (what the reference's own source does
    when it runs on smart_tree_b200.compat).
`heads="sparse_fc"` builds the checkpoint architecture, `heads="mlp"` the one in
the reference's HEAD source."""
import functools

import torch.nn as nn
import torch.nn.functional as F

from .. import spconv
from ..engine import SmartTreeEngine
from .model_blocks import MLP, SparseFC, SubMConvBlock, UBlock


class Smart_Tree(nn.Module):
    def __init__(self, input_channels, unet_planes, radius_fc_planes, direction_fc_planes, class_fc_planes, bias=False,
                 algo=spconv.ConvAlgo.Native, heads="sparse_fc", bn_eps=1e-4, fused=True):
        super().__init__()
        norm_fn = functools.partial(nn.BatchNorm1d, eps=bn_eps, momentum=0.1)
        activation_fn = nn.ReLU
        self.bn_eps = bn_eps
        self.fused = fused
        self.input_conv = SubMConvBlock(input_channels, unet_planes[0], kernel_size=1, padding=1, norm_fn=norm_fn,
                                        activation_fn=activation_fn)
        self.UNet = UBlock(unet_planes, norm_fn, activation_fn, key_id=1, algo=algo)
        if heads == "sparse_fc":
            mk = lambda planes: SparseFC(planes, norm_fn, activation_fn, algo=algo)
        elif heads == "mlp":
            mk = lambda planes: MLP(planes, norm_fn, activation_fn, bias=True)
        else:
            raise ValueError(heads)
        self.radius_head = mk(radius_fc_planes)
        self.direction_head = mk(direction_fc_planes)
        self.class_head = mk(class_fc_planes)
        self.apply(self.set_bn_init)
        self._engine = None

    @staticmethod
    def set_bn_init(m):
        if m.__class__.__name__.find("BatchNorm") != -1:
            m.weight.data.fill_(1.0)
            m.bias.data.fill_(0.0)

    @classmethod
    def from_state_dict(cls, sd, bn_eps=1e-4, fused=True):
        """Rebuild the module from a checkpoint alone (the reference needs the pickled module,
        model_inference.py:11-16, which drags spconv/cumm/omegaconf into the unpickler)."""
        planes, pre = [], "UNet."
        while pre + "Head.sequence.0.weight" in sd:
            planes.append(sd[pre + "Head.sequence.0.weight"].shape[0])
            pre += "U."
        heads = "mlp" if sd["radius_head.sequence.0.weight"].dim() == 2 else "sparse_fc"

        def head_planes(name):
            out, i = [sd[f"{name}.sequence.0.weight"].shape[-1]], 0
            while f"{name}.sequence.{i}.weight" in sd:
                out.append(sd[f"{name}.sequence.{i}.weight"].shape[0])
                i += 3
            return out

        m = cls(sd["input_conv.sequence.0.weight"].shape[-1], planes, head_planes("radius_head"), head_planes("direction_head"),
                head_planes("class_head"), heads=heads, bn_eps=bn_eps, fused=fused)
        m.load_state_dict(sd)
        return m.eval()

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    def engine(self) -> SmartTreeEngine:
        dev = self.input_conv.sequence[0].weight.device
        if self._engine is None or self._engine.device != dev:
            self._engine = SmartTreeEngine(self.state_dict(), device=dev, eps=self.bn_eps)
        return self._engine

    def forward(self, input):
        if self.fused and not self.training:
            return self.engine().forward(input.features, input.indices)
        x = self.input_conv(input)
        unet_out = self.UNet(x)
        return {"radius": self.radius_head(unet_out).features,
                "direction": F.normalize(self.direction_head(unet_out).features),
                "class_l": self.class_head(unet_out).features}
