from .branch import BranchSkeleton
from .cloud import Cloud
from .tree import DisjointTreeSkeleton, TreeSkeleton
from .tube import Tube

__all__ = ["Cloud", "BranchSkeleton", "TreeSkeleton", "DisjointTreeSkeleton", "Tube"]
