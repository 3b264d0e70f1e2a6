"""Minimal `spconv` / `spconv.pytorch` stand-in backed by libst_b200 -- exactly the surface the
reference model code touches: SparseConvTensor, SparseModule,
SparseSequential, SubMConv3d, SparseConv3d, SparseInverseConv3d, ConvAlgo, utils.PointToVoxel.
`smart_tree_b200.compat.install()` registers it as `spconv` so the reference's own
smart_tree/model/{model,model_blocks}.py run unchanged on the B200 kernels."""
from .pytorch import (ConvAlgo, SparseConv3d, SparseConvTensor, SparseInverseConv3d, SparseModule,
                      SparseSequential, SubMConv3d)

__all__ = ["ConvAlgo", "SparseConvTensor", "SparseModule", "SparseSequential", "SubMConv3d", "SparseConv3d",
           "SparseInverseConv3d"]
