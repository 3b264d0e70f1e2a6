"""One skeleton branch (/root/reference/smart_tree/data_types/branch.py:19-75)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch

from .tube import Tube


@dataclass
class BranchSkeleton:
    _id: int
    parent_id: int
    xyz: torch.Tensor     # [N,3]
    radii: torch.Tensor   # [N,1]
    child_id: Optional[int] = None
    # (store, row, length): this branch's nodes are rows row+1 .. row+length of a shared [*,4] node array
    # with one spare row in front (used by the batched post-processing fast paths); None = standalone
    _flat: Optional[tuple] = field(default=None, repr=False, compare=False)

    def __post_init__(self):
        # the reference type-checks these shapes with torchtyping (tests/type_checks.py:9-13)
        if self.xyz.dim() != 2 or self.xyz.shape[1] != 3:
            raise TypeError(f"xyz must be [N,3], got {tuple(self.xyz.shape)}")
        if self.radii.dim() != 2 or self.radii.shape != (self.xyz.shape[0], 1):
            raise TypeError(f"radii must be [N,1], got {tuple(self.radii.shape)}")

    def __len__(self):
        return self.xyz.shape[0]

    def to_tubes(self) -> List[Tube]:
        return [Tube(self.xyz[i], self.xyz[i + 1], self.radii[i], self.radii[i + 1]) for i in range(len(self) - 1)]

    def filter(self, mask) -> "BranchSkeleton":
        return BranchSkeleton(self._id, self.parent_id, self.xyz[mask], self.radii[mask], self.child_id)

    @property
    def length(self):
        return (self.xyz[1:] - self.xyz[:-1]).norm(dim=1).sum()

    @property
    def initial_radius(self):
        return torch.max(self.radii[0], self.radii[-1])

    @property
    def biggest_radius_idx(self):
        return torch.argmax(self.radii)

    @property
    def biggest_radius(self):
        return torch.max(self.radii)
