"""ctypes binding of libst_b200.so (C ABI in include/st_b200.h).

There is NO CPU fallback: if the shared library is missing this module raises at first use,
and every op raises if the tensors are not on a CUDA device."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libst_b200.so")

_lib = None

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_f = C.c_float
_sz = C.c_size_t
_pi64 = C.POINTER(C.c_int64)
_pi32 = C.POINTER(C.c_int32)

# name -> (restype, argtypes); mirrors include/st_b200.h one to one
SIGNATURES = {
    "st_version": (C.c_int, []),
    "st_last_error": (C.c_char_p, []),
    "st_device_check": (C.c_int, [C.c_int]),
    "st_sm_count": (C.c_int, [C.c_int]),
    "st_voxelize_workspace_bytes": (_sz, [_i64]),
    "st_voxelize": (C.c_int, [_p, _i64, C.c_int, _p, _p, _p, _i32, _f, _p, _p, _p, _pi64, _p, _sz, _p]),
    "st_block_workspace_bytes": (_sz, [_i64, _i64]),
    "st_block_list": (C.c_int, [_p, _i64, _f, C.c_int, _p, _p, _i32, _pi64, _p, _sz, _p]),
    "st_block_count": (C.c_int, [_p, _i64, _p, _i32, _f, _f, _f, C.c_int, _p, _pi64, _p, _sz, _p]),
    "st_block_emit": (C.c_int, [_p, _i64, _p, _i32, _f, _f, _f, C.c_int, _p, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "st_gather_rows": (C.c_int, [_p, _p, _i64, C.c_int, _p, _p]),
    "st_devoxelize": (C.c_int, [_p, _i64, _p, _p, _p, _i64, _p, _f, _p, _p, _p, _p, _p, _p]),
    "st_hash_capacity": (_i64, [_i64]),
    "st_hash_build": (C.c_int, [_p, _i64, _p, _p, _i64, _p, _p]),
    "st_subm_map": (C.c_int, [_p, _i64, _p, _p, _i64, _p, _p, _p]),
    "st_strided_coords_workspace_bytes": (_sz, [_i64]),
    "st_strided_coords": (C.c_int, [_p, _i64, C.c_int, _p, _p, _pi64, _p, _sz, _p]),
    "st_morton_workspace_bytes": (_sz, [_i64]),
    "st_morton_perm": (C.c_int, [_p, _i64, _p, _p, _sz, _p]),
    "st_strided_maps": (C.c_int, [_p, _i64, _i64, _p, _p, _i64, _p, _p, _p]),
    "st_conv_gather": (C.c_int, [_p, C.c_int, _p, _i64, C.c_int, _p, C.c_int, C.c_int, _p, _p, _p, C.c_int,
                                 _p, C.c_int, _p, C.c_int, _p, C.c_int, C.c_int, _p]),
    "st_stem_conv": (C.c_int, [_p, C.c_int, _p, _i64, _p, C.c_int, C.c_int, _p, _p, _p, C.c_int, C.c_int, _p]),
    "st_conv_gather_inv": (C.c_int, [_p, C.c_int, _p, _p, _p, _i64, C.c_int, _p, C.c_int, C.c_int, _p, _p, _p, C.c_int, C.c_int, _p]),
    "st_brick_plan_bytes": (_sz, [_i64]),
    "st_brick_plan_build": (C.c_int, [_p, _i64, _p, _p, _i64, _p, _sz, _p]),
    "st_brick_plan_info": (C.c_int, [_p, _i64, _p]),
    "st_conv_brick": (C.c_int, [_p, C.c_int, _p, _i64, _p, C.c_int, C.c_int, _p, _p, _p, C.c_int, _p, C.c_int, _p, C.c_int, _p, C.c_int,
                                C.c_int, _p]),
    "st_conv_tc_weight_floats": (_i64, [C.c_int, C.c_int, C.c_int]),
    "st_conv_tc_prepare": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, _p, _p]),
    "st_conv_tc_weight_floats_fused": (_i64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "st_conv_tc_prepare_fused": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, _p, C.c_int, _p, _p, _p]),
    "st_conv_gather_tc": (C.c_int, [_p, C.c_int, _p, _i64, C.c_int, _p, C.c_int, C.c_int, _p, _p, _p, C.c_int,
                                    _p, C.c_int, _p, C.c_int, _p, C.c_int, C.c_int, _p]),
    "st_inverse_plan_workspace_bytes": (_sz, [_i64]),
    "st_inverse_plan": (C.c_int, [_p, _p, _i64, C.c_int, _p, _p, _p, _p, _sz, _p]),
    "st_strided_maps_inv_workspace_bytes": (_sz, [_i64]),
    "st_strided_maps_inv": (C.c_int, [_p, _i64, _i64, _p, _p, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "st_conv_gather_tc_inv": (C.c_int, [_p, C.c_int, _p, _p, _p, _i64, C.c_int, _p, C.c_int, C.c_int, _p, _p, _p, C.c_int, C.c_int, _p]),
    "st_conv_plan_bytes": (_sz, [_i64]),
    "st_conv_plan_build": (C.c_int, [_p, _i64, C.c_int, _i64, _p, _sz, _p]),
    "st_conv_tp_supported": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "st_conv_gather_tp": (C.c_int, [_p, C.c_int, _p, _i64, C.c_int, _p, C.c_int, C.c_int, _p, _p, _p, C.c_int,
                                    _p, C.c_int, _p, C.c_int, _p, C.c_int, C.c_int, _p]),
    "st_heads_fused": (C.c_int, [_p, C.c_int, _i64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "st_knn_workspace_bytes": (_sz, [_i64]),
    "st_knn": (C.c_int, [_p, _i64, _p, _i64, C.c_int, _f, _p, _p, _p, _p, _sz, _p]),
    "st_outlier_mask": (C.c_int, [_p, _i64, _p, _f, C.c_int, _p, _p, _sz, _p]),
    "st_edges_workspace_bytes": (_sz, [_i64, C.c_int]),
    "st_edges_from_knn": (C.c_int, [_p, _p, _i64, C.c_int, _p, _p, _p, _pi64, _p, _sz, _p]),
    "st_connected_components": (C.c_int, [_p, _i64, _i64, _p, _p, _p]),
    "st_csr_workspace_bytes": (_sz, [_i64, _i64]),
    "st_csr_build": (C.c_int, [_p, _p, _i64, _p, _i64, _p, _p, _p, _pi64, _p, _sz, _p]),
    "st_sssp": (C.c_int, [_p, _p, _p, _i64, _p, _i32, _f, _p, _p, _pi32, _p, _p, _p]),
    "st_tree_distances": (C.c_int, [_p, _p, _p, _i64, _p, _p, _p]),
    "st_sample_tree_workspace_bytes": (_sz, [_i64, _i32]),
    "st_sample_tree": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i64, _f, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "st_repair_branches": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _p]),
    "st_points_to_tubes": (C.c_int, [_p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "st_finish_skeletons_out_ints": (C.c_size_t, [_i64]),
    "st_finish_skeletons": (C.c_int, [_p, _p, _p, _i32, _i64, _p, _p, _p, _p, _p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, _p, _p, _p]),
}


class StB200Error(RuntimeError):
    pass


def load():
    """Load libst_b200.so; raises (loudly) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StB200Error(
            f"{LIB_PATH} not found: build the sm_100a kernels first (python -m smart_tree_b200.build or "
            f"__graft_entry__.build()).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().st_last_error().decode("utf-8", "replace")
        raise StB200Error(f"{what} failed ({rc}): {msg}")
