"""CPU oracle for the smart-tree hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``smart-tree_b200/``) never imports it and fails loudly when its CUDA library is missing.

What it restates (numpy / scipy / torch-CPU, fp32 like the reference):
  * ``unet_ref``      spconv semantics (PointToVoxel, SubMConv3d, SparseConv3d,
                      SparseInverseConv3d, BatchNorm1d eval) + Smart_Tree.forward
                      -- /root/reference/smart_tree/model/{model,model_blocks}.py
  * ``skeleton_ref``  FRNN kNN, outlier removal, nn_graph/make_edges, cugraph connected
                      components / sssp, pred_graph distances, sample_tree
                      -- /root/reference/smart_tree/skeleton/*.py, data_types/graph.py
  * ``pipeline_ref``  CentreCloud, block tiling, ModelInference.forward, class filter,
                      Skeletonizer.forward, prune / repair / smooth
                      -- /root/reference/smart_tree/{pipeline.py,dataset/dataset.py,...}

PARITY STATUS: **parity unpinned at the third-party boundary, pinned for the reference's own glue** (the
reference's Python runs unmodified on oracle/fake_thirdparty.py -> tests/golden/refglue_*.npz).  The reference holds no
golden vectors, known-answer tests or fixtures for this path and the
libraries that carry its arithmetic (spconv, FRNN, cugraph 23.02) are neither vendored
nor installable here, so the oracle's spconv/FRNN/cugraph semantics are restated from
their documented behaviour and pinned only by (a) dense-equivalence checks against
torch.nn.functional.conv3d / conv_transpose3d, scipy cKDTree and scipy csgraph
(tests/test_oracle_*.py) and (b) the reference's OWN Python glue executed here against
these restatements (oracle/refglue.py -> tests/golden/*.npz).
"""
