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


class PackedBranchSkeleton(BranchSkeleton):
    """A branch whose nodes are rows [start, start+count) of one packed host array shared by the whole result
    (nodes[R,3], radii[R]): `xyz` / `radii` are sliced out on first access instead of when the skeleton is
    assembled (a few hundred to thousands of branches per tree; a tensor view costs ~1 us of host time each).
    Same attributes, methods and shapes as BranchSkeleton; assigning xyz / radii replaces the view."""

    def __init__(self, _id, parent_id, nodes, radii, start, count, radii_1d=False):
        self._id = _id
        self.parent_id = parent_id
        self.child_id = None
        self._flat = None
        self._src = (nodes, radii, start, count, radii_1d)
        self._xyz = None
        self._radii = None

    def __len__(self):
        return self._src[3] if self._xyz is None else self._xyz.shape[0]

    @property
    def xyz(self):
        if self._xyz is None:
            nodes, _, s, c, _ = self._src
            self._xyz = nodes[s:s + c]
        return self._xyz

    @xyz.setter
    def xyz(self, v):
        self._xyz = v

    @property
    def radii(self):
        if self._radii is None:
            _, rad, s, c, one_d = self._src
            r = rad[s:s + c]
            self._radii = r if one_d else r.unsqueeze(1)      # smoothed radii are 1-D (quirk C-17)
        return self._radii

    @radii.setter
    def radii(self, v):
        self._radii = v

    def __repr__(self):
        return f"BranchSkeleton(_id={self._id}, parent_id={self.parent_id}, nodes={len(self)})"
