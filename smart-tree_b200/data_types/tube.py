"""Tapered tube segment (/root/reference/smart_tree/data_types/tube.py:8-50)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch


@dataclass
class Tube:
    a: torch.Tensor   # [3] start
    b: torch.Tensor   # [3] end
    r1: torch.Tensor  # start radius
    r2: torch.Tensor  # end radius


@dataclass
class CollatedTube:
    a: torch.Tensor   # [M,3]
    b: torch.Tensor   # [M,3]
    r1: torch.Tensor  # [M]
    r2: torch.Tensor  # [M]


def collate_tubes(tubes: List[Tube]) -> CollatedTube:
    return CollatedTube(torch.stack([t.a.reshape(3) for t in tubes]), torch.stack([t.b.reshape(3) for t in tubes]),
                        torch.stack([t.r1.reshape(()) for t in tubes]), torch.stack([t.r2.reshape(()) for t in tubes]))
