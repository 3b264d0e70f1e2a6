"""smart-tree_b200: B200-native (sm_100a) implementation of smart-tree's hot path --
sparse 3D UNet inference -> medial-axis projection -> skeleton graph extraction --
behind the reference's own Python interface.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"
