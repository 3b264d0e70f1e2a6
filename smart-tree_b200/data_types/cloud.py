"""Point-cloud container with the reference's value semantics
(/root/reference/smart_tree/data_types/cloud.py:21-28,72-103,138-161,194-231,255-260).
The open3d converters / viewers of the reference are out of scope (open3d is a GUI dependency)."""
from __future__ import annotations

from dataclasses import dataclass, fields
from pathlib import Path
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

_TENSOR_FIELDS = ("xyz", "rgb", "medial_vector", "branch_direction", "branch_ids", "class_l")


@dataclass
class Cloud:
    xyz: torch.Tensor                                  # [N,3]
    rgb: Optional[torch.Tensor] = None                 # [N,3]
    medial_vector: Optional[torch.Tensor] = None       # [N,3]
    branch_direction: Optional[torch.Tensor] = None    # [N,3]
    branch_ids: Optional[torch.Tensor] = None          # [N,1]
    class_l: Optional[torch.Tensor] = None             # [N,1]
    filename: Optional[Path] = None

    def __post_init__(self):
        if self.xyz.dim() != 2 or self.xyz.shape[1] != 3:
            raise TypeError(f"xyz must be [N,3], got {tuple(self.xyz.shape)}")
        n = self.xyz.shape[0]
        for name in _TENSOR_FIELDS[1:]:
            t = getattr(self, name)
            if t is not None and (t.dim() != 2 or t.shape[0] != n):
                raise TypeError(f"{name} must be [N,k] with N={n}, got {tuple(t.shape)}")

    def __len__(self):
        return self.xyz.shape[0]

    def __str__(self):
        return (f"Cloud with {len(self)} points, min {self.xyz.min(0)[0].tolist()}, "
                f"max {self.xyz.max(0)[0].tolist()}, device {self.xyz.device}")

    def _map(self, fn) -> "Cloud":
        kw = {f.name: getattr(self, f.name) for f in fields(self)}
        for name in _TENSOR_FIELDS:
            if kw[name] is not None:
                kw[name] = fn(kw[name])
        return Cloud(**kw)

    # cloud.py:72-95 -- boolean mask or index tensor; every field is gathered
    def filter(self, mask) -> "Cloud":
        mask = mask.to(self.xyz.device)
        if mask.dtype == torch.bool:          # one compaction of the mask, then plain row gathers for every field
            mask = mask.nonzero().squeeze(1)
        return self._map(lambda t: t.index_select(0, mask))

    # cloud.py:97-103
    def filter_by_class(self, classes) -> "Cloud":
        classes = torch.as_tensor(classes, device=self.class_l.device)
        return self.filter(torch.isin(self.class_l, classes).view(-1))

    def to_device(self, device, non_blocking: bool = False) -> "Cloud":
        """non_blocking: asynchronous host->device copies (effective for pinned host tensors: load_cloud(pin_memory=True))."""
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def cpu(self) -> "Cloud":
        return self.to_device(torch.device("cpu"))

    def pin_memory(self) -> "Cloud":
        return self._map(lambda t: t.pin_memory())

    def cat(self):
        return torch.cat((self.xyz, self.rgb), 1)

    # cloud.py:194-202 -- the transforms keep only xyz and rgb (quirk C-12)
    def scale(self, factor) -> "Cloud":
        return Cloud(self.xyz * factor, self.rgb)

    def translate(self, xyz) -> "Cloud":
        return Cloud(self.xyz + xyz.to(self.xyz.device), self.rgb)

    def rotate(self, rot_mat) -> "Cloud":
        return Cloud(torch.matmul(self.xyz, rot_mat.to(self.xyz.dtype).to(self.xyz.device)), self.rgb)

    @property
    def root_idx(self) -> int:
        return int(torch.argmin(self.xyz[:, 1]).item())

    @property
    def number_classes(self) -> int:
        return 1 if self.class_l is None else int(self.class_l.max().item()) + 1

    @property
    def max_xyz(self):
        return self.xyz.max(0)[0]

    @property
    def min_xyz(self):
        return self.xyz.min(0)[0]

    @property
    def bbox(self):
        half = (self.max_xyz - self.min_xyz) / 2
        return self.min_xyz + half, half

    @property
    def medial_pts(self):
        return self.xyz + self.medial_vector

    @property
    def radius(self):
        # cloud.py:255-256 (`pow(2).sum(1).sqrt()`); the summation order is spelled out so that the
        # value is bit-identical on every device: sqrt((x*x + y*y) + z*z), one rounding per operation
        v = self.medial_vector
        return ((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]).sqrt()

    @property
    def direction(self):
        return F.normalize(self.medial_vector)

    # cloud.py:234-252 (incl. the legacy "vector" key)
    @staticmethod
    def from_numpy(**kwargs) -> "Cloud":
        out = {}
        for key, value in kwargs.items():
            if key in _TENSOR_FIELDS:
                out[key] = torch.tensor(np.asarray(value)).float()
            elif key == "vector":
                out["medial_vector"] = torch.tensor(np.asarray(value))
        for key in ("branch_ids", "class_l"):
            if key in out and out[key].dim() == 1:
                out[key] = out[key].unsqueeze(1)
        return Cloud(**out)
