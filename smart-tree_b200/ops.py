"""Torch-tensor wrappers over the C ABI (include/st_b200.h).  PyTorch supplies device memory
and streams only; all arithmetic happens in libst_b200.so.  CUDA tensors only -- no fallback."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

I32, I64, F32, U8 = torch.int32, torch.int64, torch.float32, torch.uint8

# ---- instrumentation (bench.py): number of libst_b200 kernels launched, optional per-conv CUDA events
LAUNCHES = 0
LAST_SSSP_CTL = None
_conv_profile = None
_KERNELS_PER_CALL = {"blocks": 7, "voxelize": 3, "hash_build": 1, "subm_map": 1, "strided_coords": 2, "strided_maps": 1, "conv": 1, "stem": 1,
                     "heads": 1, "knn": 4, "outlier": 4, "edges": 2, "cc": 5, "csr": 2, "sssp": 6, "tree_dist": 3,
                     "sample_tree": 5, "tubes": 1, "repair": 1, "finish_skeletons": 1, "plan": 1, "inverse_plan": 3, "devoxelize": 1, "gather_rows": 1}


def _count(op):
    global LAUNCHES
    LAUNCHES += _KERNELS_PER_CALL[op]


def conv_profile(enable: bool):
    """Start/stop recording (cin, cout, taps, n_out, extra_bytes, start_event, end_event) per conv launch."""
    global _conv_profile
    prev = _conv_profile
    _conv_profile = [] if enable else None
    return prev


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def _req(t: torch.Tensor, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.StB200Error(f"{name}: expected a CUDA tensor (libst_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.StB200Error(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.StB200Error(f"{name}: expected a contiguous tensor")
    return t


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=U8, device=device)


# ------------------------------------------------------------------ block tiling
def block_tiling(xyz, block_size, buffer_size, min_points=20, max_blocks=1 << 16):
    """Returns block_ids [B,3] f32, point_index [T] i64, point_block [T] i32, block_lo / block_hi [B,3] f32."""
    import numpy as np
    lib = _lib.load()
    _req(xyz, F32, "xyz")
    n, dev = xyz.shape[0], xyz.device
    ids = torch.empty((max_blocks, 3), dtype=F32, device=dev)
    keys = torch.empty(max_blocks, dtype=I64, device=dev)
    ws = _ws(lib.st_block_workspace_bytes(n, n), dev)
    nb = C.c_int64(0)
    _count("blocks")
    _lib.check(lib.st_block_list(_ptr(xyz), n, float(block_size), int(min_points), _ptr(ids), _ptr(keys), max_blocks,
                                 C.byref(nb), _ptr(ws), ws.numel(), _stream()), "st_block_list")
    nb = nb.value
    ids = ids[:nb]
    # the reference evaluates these bounds in fp32 tensor arithmetic with Python-float scalars
    half_block = float(np.float32(block_size / 2))
    half_cube = float(np.float32((block_size + buffer_size * 2) / 2))
    reach = int(-(-buffer_size // block_size)) if buffer_size > 0 else 0
    offsets = torch.empty(max(n, 1), dtype=I32, device=dev)
    npairs = C.c_int64(0)
    _lib.check(lib.st_block_count(_ptr(xyz), n, _ptr(keys), nb, float(block_size), half_block, half_cube, reach, _ptr(offsets),
                                  C.byref(npairs), _ptr(ws), ws.numel(), _stream()), "st_block_count")
    t = npairs.value
    pidx = torch.empty(t, dtype=I64, device=dev)
    pblk = torch.empty(t, dtype=I32, device=dev)
    lo = torch.empty((nb, 3), dtype=F32, device=dev)
    hi = torch.empty((nb, 3), dtype=F32, device=dev)
    if t > n:
        ws = _ws(lib.st_block_workspace_bytes(n, t), dev)
    _lib.check(lib.st_block_emit(_ptr(xyz), n, _ptr(keys), nb, float(block_size), half_block, half_cube, reach, _ptr(offsets), t,
                                 _ptr(pidx), _ptr(pblk), _ptr(lo), _ptr(hi), _ptr(ws), ws.numel(), _stream()), "st_block_emit")
    return ids, pidx, pblk, lo, hi


# ------------------------------------------------------------------ voxelise
def voxelize(points, point_block, block_lo, block_grid, vsize):
    """points [n,F] f32, point_block [n] i32 or None, block_lo [B,3] f32, block_grid [B,3] i32.
    Returns pc_voxel_id[n] i32, rep_point[M] i32, coords[M,4] i32 (b,z,y,x)."""
    lib = _lib.load()
    _req(points, F32, "points"); _req(block_lo, F32, "block_lo"); _req(block_grid, I32, "block_grid")
    if point_block is not None:
        _req(point_block, I32, "point_block")
    n, ld = points.shape
    dev = points.device
    pc = torch.empty(n, dtype=I32, device=dev)
    rep = torch.empty(n, dtype=I32, device=dev)
    coords = torch.empty((n, 4), dtype=I32, device=dev)
    wsb = lib.st_voxelize_workspace_bytes(n)
    ws = _ws(wsb, dev)
    m = C.c_int64(0)
    _count("voxelize")
    _lib.check(lib.st_voxelize(_ptr(points), n, ld, _ptr(point_block), _ptr(block_lo), _ptr(block_grid),
                               block_lo.shape[0], float(vsize), _ptr(pc), _ptr(rep), _ptr(coords), C.byref(m),
                               _ptr(ws), ws.numel(), _stream()), "st_voxelize")
    return pc, rep[:m.value], coords[:m.value]


def gather_rows(src, idx):
    """src[idx] for a 2-D contiguous tensor of 4-byte elements and an int32 index vector (st_gather_rows)."""
    lib = _lib.load()
    if idx.dtype != I32 or src.dim() != 2 or not src.is_contiguous() or src.element_size() != 4 or not src.is_cuda:
        return src.index_select(0, idx)
    out = torch.empty((idx.shape[0], src.shape[1]), dtype=src.dtype, device=src.device)
    _count("gather_rows")
    _lib.check(lib.st_gather_rows(_ptr(src), _ptr(idx), idx.shape[0], src.shape[1] * 4, _ptr(out), _stream()), "st_gather_rows")
    return out


def devoxelize(xyz, pair_point, pair_block, pair_voxel, block_centres, block_size, voxel_medial, voxel_class):
    """Per-voxel predictions back to every input point (st_devoxelize).  Returns point_medial [n,3] f32,
    point_class [n] i32 (-1 = no prediction), point_voxel [n] i32 (-1 = none)."""
    lib = _lib.load()
    _req(xyz, F32, "xyz"); _req(pair_point, I64, "pair_point"); _req(pair_block, I32, "pair_block"); _req(pair_voxel, I32, "pair_voxel")
    _req(block_centres, F32, "block_centres"); _req(voxel_medial, F32, "voxel_medial"); _req(voxel_class, I32, "voxel_class")
    n, dev = xyz.shape[0], xyz.device
    pm = torch.empty((n, 3), dtype=F32, device=dev)
    pc = torch.empty(n, dtype=I32, device=dev)
    pv = torch.empty(n, dtype=I32, device=dev)
    _count("devoxelize")
    _lib.check(lib.st_devoxelize(_ptr(xyz), n, _ptr(pair_point), _ptr(pair_block), _ptr(pair_voxel), pair_point.shape[0], _ptr(block_centres),
                                 float(block_size), _ptr(voxel_medial), _ptr(voxel_class), _ptr(pm), _ptr(pc), _ptr(pv), _stream()),
               "st_devoxelize")
    return pm, pc, pv


# ------------------------------------------------------------------ coordinate table / maps
class CoordTable:
    """Open-addressing hash of (b,z,y,x) -> row.  Coordinates must satisfy 0 <= z,y,x <= 65533 and 0 <= b < 32768 (the
    64-bit keys hold 16 bits per field); `check()` raises if any row violated that (one 4-byte read-back)."""

    def __init__(self, coords):
        lib = _lib.load()
        _req(coords, I32, "coords")
        self.n = coords.shape[0]
        self.capacity = lib.st_hash_capacity(self.n)
        self.keys = torch.empty(self.capacity, dtype=I64, device=coords.device)
        self.vals = torch.empty(self.capacity, dtype=I32, device=coords.device)
        self.status = torch.empty(1, dtype=I32, device=coords.device)
        _count("hash_build")
        _lib.check(lib.st_hash_build(_ptr(coords), self.n, _ptr(self.keys), _ptr(self.vals), self.capacity, _ptr(self.status),
                                     _stream()), "st_hash_build")

    def check(self):
        if int(self.status.item()) != 0:
            raise _lib.StB200Error("voxel coordinates out of range: libst_b200 needs 0 <= z,y,x <= 65533 and 0 <= batch < 32768")
        return self


def subm_map(coords, table: CoordTable, spatial_shape=None):
    """nbr [27, n].  spatial_shape: optional int32 device tensor (z,y,x) -> strict_spconv_bounds clipping (st_b200.h)."""
    lib = _lib.load()
    n = coords.shape[0]
    nbr = torch.empty((27, n), dtype=I32, device=coords.device)
    if spatial_shape is not None:
        _req(spatial_shape, I32, "spatial_shape")
    _count("subm_map")
    _lib.check(lib.st_subm_map(_ptr(coords), n, _ptr(table.keys), _ptr(table.vals), table.capacity, _ptr(spatial_shape), _ptr(nbr),
                               _stream()), "st_subm_map")
    return nbr


def morton_perm(coords):
    """perm[k] = row of the k-th voxel in (batch, Z-order) order."""
    lib = _lib.load()
    _req(coords, I32, "coords")
    n = coords.shape[0]
    perm = torch.empty(n, dtype=I32, device=coords.device)
    ws = _ws(lib.st_morton_workspace_bytes(n), coords.device)
    _count("subm_map")
    _lib.check(lib.st_morton_perm(_ptr(coords), n, _ptr(perm), _ptr(ws), ws.numel(), _stream()), "st_morton_perm")
    return perm


def strided_coords(coords, morton=False, out_shape=None):
    lib = _lib.load()
    _req(coords, I32, "coords")
    n = coords.shape[0]
    out = torch.empty((max(8 * n, 1), 4), dtype=I32, device=coords.device)
    ws = _ws(lib.st_strided_coords_workspace_bytes(n), coords.device)
    m = C.c_int64(0)
    if out_shape is not None:
        _req(out_shape, I32, "out_shape")
    _count("strided_coords")
    _lib.check(lib.st_strided_coords(_ptr(coords), n, 1 if morton else 0, _ptr(out_shape), _ptr(out), C.byref(m), _ptr(ws), ws.numel(),
                                     _stream()), "st_strided_coords")
    return out[:m.value].clone() if m.value * 4 < out.shape[0] else out[:m.value]


def strided_maps(coords, out_coords, out_table: CoordTable, inverse_plan=False):
    """down [27, m], up [27, n]; with inverse_plan the `up` map is produced only in parity-sorted launch order:
    returns (down, None, (row_index, up_sorted, tile_mask)) as ops.inverse_plan would (st_strided_maps_inv)."""
    lib = _lib.load()
    n, m = coords.shape[0], out_coords.shape[0]
    down = torch.empty((27, m), dtype=I32, device=coords.device)
    if inverse_plan and n:
        buf = torch.empty(27 * n + n + (n + 127) // 128, dtype=I32, device=coords.device)
        up_sorted, row_index, tile_mask = buf[:27 * n].view(27, n), buf[27 * n:28 * n], buf[28 * n:]
        ws = _ws(lib.st_strided_maps_inv_workspace_bytes(n), coords.device)
        _count("strided_maps"); _count("inverse_plan")
        _lib.check(lib.st_strided_maps_inv(_ptr(coords), n, m, _ptr(out_table.keys), _ptr(out_table.vals), out_table.capacity,
                                           _ptr(down), _ptr(row_index), _ptr(up_sorted), _ptr(tile_mask), _ptr(ws), ws.numel(),
                                           _stream()), "st_strided_maps_inv")
        return down, None, (row_index, up_sorted, tile_mask)
    up = torch.empty((27, n), dtype=I32, device=coords.device)
    if inverse_plan:          # (empty level: nothing to plan)
        return down, up, None
    _count("strided_maps")
    _lib.check(lib.st_strided_maps(_ptr(coords), n, m, _ptr(out_table.keys), _ptr(out_table.vals), out_table.capacity,
                                   _ptr(down), _ptr(up), _stream()), "st_strided_maps")
    return down, up


# ------------------------------------------------------------------ convolution
def _ld(t):
    return t.stride(0) if t is not None else 0


def _req_rows(t, name):
    """2-D fp32 CUDA tensor whose rows are contiguous (a column slice of a wider buffer is fine)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != F32 or t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise _lib.StB200Error(f"{name}: expected a 2-D fp32 CUDA tensor with unit column stride")
    return t


def conv_tc_supported(ntaps, cin, cout):
    return _lib.load().st_conv_tc_weight_floats(ntaps, cin, cout) > 0


def conv_tc_prepare(weight):
    """weight [ntaps, cin, cout] -> the tensor-core path's pre-arranged (hi, lo) core-matrix tiles."""
    lib = _lib.load()
    _req(weight, F32, "weight")
    ntaps, cin, cout = weight.shape
    nfl = lib.st_conv_tc_weight_floats(ntaps, cin, cout)
    if nfl < 0:
        raise _lib.StB200Error(f"tensor-core conv does not support cin={cin}, cout={cout}")
    wprep = torch.empty(nfl, dtype=F32, device=weight.device)
    _lib.check(lib.st_conv_tc_prepare(_ptr(weight), ntaps, cin, cout, _ptr(wprep), _stream()), "st_conv_tc_prepare")
    return wprep


def conv_tc_prepare_fused(weight, w2, scale):
    """Tensor-core weights of a conv with a fused ResBlock identity: weight [ntaps, cin, cout] followed by extra K
    stages holding w2 [cin2, cout] / scale, so that conv_gather(..., impl="tc", in2=x, w2=None, weight_tc=this)
    computes act(scale * conv(in) + shift + w2 . in2) in one accumulation.  None if the shapes are unsupported."""
    lib = _lib.load()
    _req(weight, F32, "weight"); _req(w2, F32, "w2")
    ntaps, cin, cout = weight.shape
    cin2 = w2.shape[0]
    nfl = lib.st_conv_tc_weight_floats_fused(ntaps, cin, cout, cin2)
    if nfl < 0:
        return None
    if scale is not None:
        _req(scale, F32, "scale")
    wprep = torch.empty(nfl, dtype=F32, device=weight.device)
    _lib.check(lib.st_conv_tc_prepare_fused(_ptr(weight), ntaps, cin, cout, _ptr(w2), cin2, _ptr(scale), _ptr(wprep), _stream()),
               "st_conv_tc_prepare_fused")
    return wprep


def inverse_plan(coords, up):
    """Parity-sorted launch order of an inverse conv (st_inverse_plan): coords [n,4] of the fine level, up [ntaps, n].
    Returns (row_index [n] i32, up_sorted [ntaps, n] i32, tile_mask [ceil(n/128)] as int32 bit masks)."""
    lib = _lib.load()
    _req(coords, I32, "coords"); _req(up, I32, "up")
    ntaps, n = up.shape
    dev = up.device
    row_index = torch.empty(n, dtype=I32, device=dev)
    up_sorted = torch.empty_like(up)
    tile_mask = torch.empty(max((n + 127) // 128, 1), dtype=I32, device=dev)
    ws = _ws(lib.st_inverse_plan_workspace_bytes(n), dev)
    _count("inverse_plan")
    _lib.check(lib.st_inverse_plan(_ptr(coords), _ptr(up), n, ntaps, _ptr(row_index), _ptr(up_sorted), _ptr(tile_mask), _ptr(ws),
                                   ws.numel(), _stream()), "st_inverse_plan")
    return row_index, up_sorted, tile_mask


def conv_gather_tc_inv(inp, plan, weight_tc, ntaps, cin, cout, n_out, scale=None, shift=None, out=None, relu=False):
    """Inverse conv through the tensor-core kernel on parity-sorted rows; plan = inverse_plan(coords, up)."""
    lib = _lib.load()
    _req_rows(inp, "inp")
    row_index, up_sorted, tile_mask = plan
    if out is None:
        out = torch.empty((n_out, cout), dtype=F32, device=inp.device)
    _req_rows(out, "out")
    assert out.shape == (n_out, cout) and inp.shape[1] == cin and up_sorted.shape == (ntaps, n_out)
    _count("conv")
    prof = _conv_profile
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    _lib.check(lib.st_conv_gather_tc_inv(_ptr(inp), _ld(inp), _ptr(up_sorted), _ptr(row_index), _ptr(tile_mask), n_out, ntaps, _ptr(weight_tc),
                                         cin, cout, _ptr(scale), _ptr(shift), _ptr(out), _ld(out), 1 if relu else 0, _stream()),
               "st_conv_gather_tc_inv")
    if prof is not None:
        ev1.record()
        prof.append((cin, cout, ntaps, n_out, 0, ev0, ev1, "tcinv"))
    return out


def conv_gather_inv(inp, plan, weight, n_out, scale=None, shift=None, out=None, relu=False):
    """Inverse conv through the FMA kernel on parity-sorted rows (st_conv_gather_inv); plan = inverse_plan(coords, up),
    weight [ntaps, cin, cout]."""
    lib = _lib.load()
    _req_rows(inp, "inp"); _req(weight, F32, "weight")
    row_index, up_sorted, tile_mask = plan
    ntaps, cin, cout = weight.shape
    if out is None:
        out = torch.empty((n_out, cout), dtype=F32, device=inp.device)
    _req_rows(out, "out")
    assert out.shape == (n_out, cout) and inp.shape[1] == cin and up_sorted.shape == (ntaps, n_out)
    _count("conv")
    prof = _conv_profile
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    _lib.check(lib.st_conv_gather_inv(_ptr(inp), _ld(inp), _ptr(up_sorted), _ptr(row_index), _ptr(tile_mask), n_out, ntaps, _ptr(weight),
                                      cin, cout, _ptr(scale), _ptr(shift), _ptr(out), _ld(out), 1 if relu else 0, _stream()),
               "st_conv_gather_inv")
    if prof is not None:
        ev1.record()
        prof.append((cin, cout, ntaps, n_out, 0, ev0, ev1, "fmainv"))
    return out


def stem_conv(inp, weight, scale=None, shift=None, row_index=None, relu=True):
    """out[i] = act(scale * (W . inp[row_index[i]]) + shift) (st_stem_conv): the 1x1 input conv fused with the Z-order row
    permutation.  inp [N, >= cin] fp32 rows with unit column stride (a column slice is fine); weight [cin, cout]."""
    lib = _lib.load()
    _req_rows(inp, "inp"); _req(weight, F32, "weight")
    cin, cout = weight.shape
    n = inp.shape[0] if row_index is None else row_index.shape[0]
    if row_index is not None:
        _req(row_index, I32, "row_index")
    out = torch.empty((n, cout), dtype=F32, device=inp.device)
    _count("stem")
    _lib.check(lib.st_stem_conv(_ptr(inp), _ld(inp), _ptr(row_index), n, _ptr(weight), cin, cout, _ptr(scale), _ptr(shift), _ptr(out),
                                _ld(out), 1 if relu else 0, _stream()), "st_stem_conv")
    return out


def conv_tp_supported(ntaps, cin, cout):
    return bool(_lib.load().st_conv_tp_supported(ntaps, cin, cout))


def conv_plan_build(nbr_map, n_in):
    """Tile plan of a gather map [ntaps, n_out] into a source array of n_in rows (st_conv_plan_build):
    distinct source rows + 16-bit local map per 128-row tile, for conv_gather(..., impl="tp")."""
    lib = _lib.load()
    _req(nbr_map, I32, "map")
    ntaps, n_out = nbr_map.shape
    plan = torch.empty(max(lib.st_conv_plan_bytes(n_out), 256), dtype=U8, device=nbr_map.device)
    _count("plan")
    _lib.check(lib.st_conv_plan_build(_ptr(nbr_map), n_out, ntaps, n_in, _ptr(plan), plan.numel(), _stream()), "st_conv_plan_build")
    return plan


def conv_gather(inp, nbr_map, weight, n_out, scale=None, shift=None, residual=None, in2=None, w2=None,
                out=None, relu=False, impl="fma", weight_tc=None, plan=None):
    """out[i] = act(scale * sum_k W[k] . in[map[k,i]] + shift + residual[i] + w2 . in2[i]).
    weight [ntaps, cin, cout]; nbr_map [ntaps, n_out] i32 or None (identity, ntaps == 1).
    impl: "fma" | "tc" (tcgen05, per-thread gathers) | "tp" (tcgen05, shared-memory staged tiles; needs
    plan = conv_plan_build(nbr_map, inp.shape[0]))."""
    lib = _lib.load()
    _req_rows(inp, "inp"); _req(weight, F32, "weight")
    ntaps, cin, cout = weight.shape
    if nbr_map is not None:
        _req(nbr_map, I32, "map")
        assert nbr_map.shape == (ntaps, n_out), (nbr_map.shape, ntaps, n_out)
    if out is None:
        out = torch.empty((n_out, cout), dtype=F32, device=inp.device)
    _req_rows(out, "out")
    assert out.shape == (n_out, cout) and inp.shape[1] == cin
    if residual is not None:
        _req_rows(residual, "residual")
    if in2 is not None:
        _req_rows(in2, "in2")
        if w2 is not None:
            _req(w2, F32, "w2")
        elif impl != "tc" or weight_tc is None:
            raise _lib.StB200Error("in2 without w2 needs impl='tc' and weights from conv_tc_prepare_fused")
    fn, wptr = lib.st_conv_gather, weight
    mptr = nbr_map
    if impl in ("tc", "tp"):
        if weight_tc is None:
            weight_tc = conv_tc_prepare(weight)
        fn, wptr = lib.st_conv_gather_tc, weight_tc
        if impl == "tp":
            if plan is None:
                plan = conv_plan_build(nbr_map, inp.shape[0])
            _req(plan, U8, "plan")
            fn, mptr = lib.st_conv_gather_tp, plan
    _count("conv")
    prof = _conv_profile
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    _lib.check(fn(_ptr(inp), _ld(inp), _ptr(mptr), n_out, ntaps, _ptr(wptr), cin, cout, _ptr(scale), _ptr(shift),
                  _ptr(residual), _ld(residual), _ptr(in2), _ld(in2), _ptr(w2), (in2.shape[1] if in2 is not None else 0),
                  _ptr(out), _ld(out), 1 if relu else 0, _stream()), "st_conv_gather")
    if prof is not None:
        ev1.record()
        extra = (cout if residual is not None else 0) + (in2.shape[1] if in2 is not None else 0)
        prof.append((cin, cout, ntaps, n_out, extra, ev0, ev1, impl))
    return out


def brick_plan(coords, table: CoordTable):
    """Plan of the brick conv for one level (st_brick_plan_build): coords [n,4] i32 in (batch, Z-order)."""
    lib = _lib.load()
    _req(coords, I32, "coords")
    n = coords.shape[0]
    plan = torch.empty(lib.st_brick_plan_bytes(n), dtype=U8, device=coords.device)
    _count("subm_map")
    _lib.check(lib.st_brick_plan_build(_ptr(coords), n, _ptr(table.keys), _ptr(table.vals), table.capacity, _ptr(plan), plan.numel(), _stream()),
               "st_brick_plan_build")
    return plan


def brick_plan_info(plan, n):
    """(bricks, halo entries, status): status != 0 -> the rows were not strictly increasing in (batch, Z-order)."""
    info = (C.c_int32 * 3)()
    _lib.check(_lib.load().st_brick_plan_info(_ptr(plan), n, info), "st_brick_plan_info")
    return int(info[0]), int(info[1]), int(info[2])


def conv_brick_supported(ntaps, cin, cout):
    return ntaps == 27 and cin in (8, 16) and cout == 8


def conv_brick(inp, plan, weight, n_out, scale=None, shift=None, residual=None, in2=None, w2=None, out=None, relu=False):
    """Sub-manifold 3x3x3 conv on 4x4x4 bricks (st_conv_brick): same result as conv_gather with the level's subm_map."""
    lib = _lib.load()
    _req_rows(inp, "inp"); _req(weight, F32, "weight"); _req(plan, U8, "plan")
    ntaps, cin, cout = weight.shape
    assert conv_brick_supported(ntaps, cin, cout) and inp.shape == (n_out, cin)
    if out is None:
        out = torch.empty((n_out, cout), dtype=F32, device=inp.device)
    _req_rows(out, "out")
    if residual is not None:
        _req_rows(residual, "residual")
    if in2 is not None:
        _req_rows(in2, "in2"); _req(w2, F32, "w2")
    _count("conv")
    prof = _conv_profile
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    _lib.check(lib.st_conv_brick(_ptr(inp), _ld(inp), _ptr(plan), n_out, _ptr(weight), cin, cout, _ptr(scale), _ptr(shift), _ptr(residual),
                                 _ld(residual), _ptr(in2), _ld(in2), _ptr(w2), (in2.shape[1] if in2 is not None else 0), _ptr(out), _ld(out),
                                 1 if relu else 0, _stream()), "st_conv_brick")
    if prof is not None:
        ev1.record()
        extra = (cout if residual is not None else 0) + (in2.shape[1] if in2 is not None else 0)
        prof.append((cin, cout, ntaps, n_out, extra, ev0, ev1, "brick"))
    return out


def heads_fused(feat, params, want_logits=True, out_index=None):
    lib = _lib.load()
    _req_rows(feat, "feat"); _req(params, F32, "params")
    n, dev = feat.shape[0], feat.device
    radius = torch.empty((n, 1), dtype=F32, device=dev)
    direction = torch.empty((n, 3), dtype=F32, device=dev)
    logits = torch.empty((n, 2), dtype=F32, device=dev) if want_logits else None
    medial = torch.empty((n, 3), dtype=F32, device=dev)
    cls = torch.empty(n, dtype=I32, device=dev)
    _count("heads")
    if out_index is not None:
        _req(out_index, I32, "out_index")
    _lib.check(lib.st_heads_fused(_ptr(feat), _ld(feat), n, _ptr(params), _ptr(out_index), _ptr(radius), _ptr(direction), _ptr(logits),
                                  _ptr(medial), _ptr(cls), _stream()), "st_heads_fused")
    return radius, direction, logits, medial, cls


# ------------------------------------------------------------------ kNN / graph
_KNN_KS = (1, 2, 4, 8, 16, 24, 32, 48, 64)


def knn(src, dst, K, r, query_radius=None):
    """idx[n,K] i32 (-1 padded), d2[n,K] f32 squared distances (-1 padded)."""
    lib = _lib.load()
    _req(src, F32, "src"); _req(dst, F32, "dst")
    n, m = src.shape[0], dst.shape[0]
    k_req = int(K)
    if k_req < 1 or k_req > 64:
        raise _lib.StB200Error(f"knn: K={k_req} outside 1..64 (the register top-K kernels; the reference runs K=16)")
    K = next(k for k in _KNN_KS if k >= k_req)          # the kernel is instantiated for these; the K nearest are a prefix
    if K != k_req:
        idx, d2 = knn(src, dst, K, r, query_radius)
        return idx[:, :k_req].contiguous(), d2[:, :k_req].contiguous()
    idx = torch.empty((n, K), dtype=I32, device=src.device)
    d2 = torch.empty((n, K), dtype=F32, device=src.device)
    ws = _ws(lib.st_knn_workspace_bytes(m), src.device)
    if query_radius is not None:
        _req(query_radius, F32, "query_radius")
    _count("knn")
    _lib.check(lib.st_knn(_ptr(src), n, _ptr(dst), m, K, float(r), _ptr(query_radius), _ptr(idx), _ptr(d2), _ptr(ws),
                          ws.numel(), _stream()), "st_knn")
    return idx, d2


def outlier_mask(points, radii, r_max, nb=8):
    lib = _lib.load()
    _req(points, F32, "points"); _req(radii, F32, "radii")
    n = points.shape[0]
    keep = torch.empty(n, dtype=U8, device=points.device)
    ws = _ws(lib.st_knn_workspace_bytes(n), points.device)
    _count("outlier")
    _lib.check(lib.st_outlier_mask(_ptr(points), n, _ptr(radii), float(r_max), nb, _ptr(keep), _ptr(ws), ws.numel(), _stream()),
               "st_outlier_mask")
    return keep.bool()


def edges_from_knn(idx, d2, radii=None):
    lib = _lib.load()
    _req(idx, I32, "idx"); _req(d2, F32, "d2")
    n, K = idx.shape
    edges = torch.empty((max(n * K, 1), 2), dtype=I32, device=idx.device)
    weights = torch.empty(max(n * K, 1), dtype=F32, device=idx.device)
    ws = _ws(lib.st_edges_workspace_bytes(n, K), idx.device)
    ne = C.c_int64(0)
    _count("edges")
    _lib.check(lib.st_edges_from_knn(_ptr(idx), _ptr(d2), n, K, _ptr(radii), _ptr(edges), _ptr(weights), C.byref(ne), _ptr(ws),
                                     ws.numel(), _stream()), "st_edges_from_knn")
    return edges[:ne.value], weights[:ne.value]


def connected_components(edges, n):
    lib = _lib.load()
    _req(edges, I32, "edges")
    label = torch.empty(n, dtype=I32, device=edges.device)
    size = torch.empty(n, dtype=I32, device=edges.device)
    _count("cc")
    _lib.check(lib.st_connected_components(_ptr(edges), edges.shape[0], n, _ptr(label), _ptr(size), _stream()),
               "st_connected_components")
    return label, size


def csr_build(edges, weights, n, vertex_map=None):
    """CSR over both directions of every edge (self loops dropped).  With vertex_map [n_old] i32 the endpoints
    are renumbered on the fly (entries < 0 drop the edge) and n is the number of NEW vertices."""
    lib = _lib.load()
    _req(edges, I32, "edges"); _req(weights, F32, "weights")
    if vertex_map is not None:
        _req(vertex_map, I32, "vertex_map")
    ne, dev = edges.shape[0], edges.device
    row_ptr = torch.empty(n + 1, dtype=I32, device=dev)
    col = torch.empty(max(2 * ne, 1), dtype=I32, device=dev)
    w = torch.empty(max(2 * ne, 1), dtype=F32, device=dev)
    ws = _ws(lib.st_csr_workspace_bytes(n, ne), dev)
    na = C.c_int64(0)
    _count("csr")
    _lib.check(lib.st_csr_build(_ptr(edges), _ptr(weights), ne, _ptr(vertex_map), n, _ptr(row_ptr), _ptr(col), _ptr(w), C.byref(na), _ptr(ws),
                                ws.numel(), _stream()), "st_csr_build")
    return row_ptr, col[:na.value], w[:na.value]


def sssp(row_ptr, col, w, n, sources, want_sweeps=False, delta=0.0, orig_id=None):
    """dist [n] f32, pred [n] i32.  orig_id (optional, i32 [n]): the graph is numbered in another order than the caller's
    (spatial order, see spatial_order); results come back in the caller's numbering (include/st_b200.h)."""
    lib = _lib.load()
    _req(sources, I32, "sources")
    if orig_id is not None:
        _req(orig_id, I32, "orig_id")
    dev = row_ptr.device
    dist = torch.empty(n, dtype=F32, device=dev)
    pred = torch.empty(n, dtype=I32, device=dev)
    ctl = torch.empty(64 + 4 * n + 1088, dtype=I32, device=dev)      # control block + dirty[n] (+ seen[n], pend[n] for large graphs) + internal dist[n] + idle flags
    sweeps = C.c_int32(0)
    _count("sssp")
    _lib.check(lib.st_sssp(_ptr(row_ptr), _ptr(col), _ptr(w), n, _ptr(sources), sources.shape[0], float(delta), _ptr(dist), _ptr(pred),
                           C.byref(sweeps) if want_sweeps else None, _ptr(ctl), _ptr(orig_id), _stream()), "st_sssp")
    global LAST_SSSP_CTL
    LAST_SSSP_CTL = ctl
    return (dist, pred, sweeps.value) if want_sweeps else (dist, pred)


def spatial_order(points, group, cell=0.02):
    """Z-order of points [m,3] inside their group (int tensor [m], e.g. the component index, < 32768): returns
    perm i32 [m] (k-th vertex in spatial order) and rank i32 [m] (spatial position of each vertex).  Cells of `cell` metres;
    any order is valid for st_sssp -- this one keeps the vertex range of a CTA compact."""
    lo = points.min(0).values
    q = ((points - lo) / cell).floor().clamp_(0, 65533).int()
    coords = torch.stack([group.int(), q[:, 2], q[:, 1], q[:, 0]], 1).contiguous()
    perm = morton_perm(coords)
    rank = torch.empty_like(perm)
    rank[perm.long()] = torch.arange(perm.shape[0], dtype=I32, device=perm.device)
    return perm, rank


def tree_distances(points, pred, is_root):
    lib = _lib.load()
    _req(points, F32, "points"); _req(pred, I32, "pred"); _req(is_root, U8, "is_root")
    n = pred.shape[0]
    td = torch.empty(n, dtype=F32, device=pred.device)
    ctl = torch.empty(64 + 2 * n, dtype=I32, device=pred.device)       # control block + (pred, edge length) per vertex
    _count("tree_dist")
    _lib.check(lib.st_tree_distances(_ptr(points), _ptr(pred), _ptr(is_root), n, _ptr(td), _ptr(ctl), _stream()),
               "st_tree_distances")
    return td


def sample_tree(medial_pts, radii, pred, tree_dist, comp_off, cell_size):
    """Returns segment-local path_vertices[n], branch_len[n], branch_parent[n], comp_nb[C], comp_np[C]."""
    lib = _lib.load()
    _req(medial_pts, F32, "medial_pts"); _req(radii, F32, "radii"); _req(pred, I32, "pred")
    _req(tree_dist, F32, "tree_dist"); _req(comp_off, I32, "comp_off")
    n, dev = pred.shape[0], pred.device
    nc = comp_off.shape[0] - 1
    path = torch.empty(max(n, 1), dtype=I32, device=dev)
    blen = torch.empty(max(n, 1), dtype=I32, device=dev)
    bpar = torch.empty(max(n, 1), dtype=I32, device=dev)
    cnb = torch.zeros(max(nc, 1), dtype=I32, device=dev)
    cnp = torch.zeros(max(nc, 1), dtype=I32, device=dev)
    ws = _ws(lib.st_sample_tree_workspace_bytes(n, nc), dev)
    _count("sample_tree")
    _lib.check(lib.st_sample_tree(_ptr(medial_pts), _ptr(radii), _ptr(pred), _ptr(tree_dist), _ptr(comp_off), nc, n,
                                  float(cell_size), _ptr(path), _ptr(blen), _ptr(bpar), _ptr(cnb), _ptr(cnp), _ptr(ws),
                                  ws.numel(), _stream()), "st_sample_tree")
    return path, blen, bpar, cnb, cnp


def points_to_tubes(pts, a, b, r1, r2, tube_off):
    lib = _lib.load()
    for t, nme in ((pts, "pts"), (a, "a"), (b, "b"), (r1, "r1"), (r2, "r2")):
        _req(t, F32, nme)
    _req(tube_off, I32, "tube_off")
    nq, dev = pts.shape[0], pts.device
    vec = torch.empty((nq, 3), dtype=F32, device=dev)
    idx = torch.empty(nq, dtype=I32, device=dev)
    rr = torch.empty(nq, dtype=F32, device=dev)
    _count("tubes")
    _lib.check(lib.st_points_to_tubes(_ptr(pts), nq, _ptr(a), _ptr(b), _ptr(r1), _ptr(r2), _ptr(tube_off), _ptr(vec),
                                      _ptr(idx), _ptr(rr), _stream()), "st_points_to_tubes")
    return vec, idx, rr


def finish_skeletons(medial_pts, radii, comp_off, path, blen, bpar, cnb, cnp, prune_first=False, min_radius=0.0, min_length=0.0,
                     repair=False, smooth_kernel=0):
    """Node gather + prune + repair + smooth for every component in one launch.  Returns the packed int32
    device buffer described in include/st_b200.h (header | bmeta | nodes | smooth)."""
    lib = _lib.load()
    _req(medial_pts, F32, "medial_pts"); _req(radii, F32, "radii"); _req(comp_off, I32, "comp_off")
    for t, nm in ((path, "path"), (blen, "blen"), (bpar, "bpar"), (cnb, "cnb"), (cnp, "cnp")):
        _req(t, I32, nm)
    n = medial_pts.shape[0]
    out = torch.empty(lib.st_finish_skeletons_out_ints(n), dtype=torch.int32, device=medial_pts.device)
    depth = torch.empty(max(n, 1), dtype=torch.int32, device=medial_pts.device)
    _count("finish_skeletons")
    _lib.check(lib.st_finish_skeletons(_ptr(medial_pts), _ptr(radii), _ptr(comp_off), comp_off.shape[0] - 1, n, _ptr(path), _ptr(blen),
                                       _ptr(bpar), _ptr(cnb), _ptr(cnp), int(bool(prune_first)), float(min_radius), float(min_length),
                                       int(bool(repair)), int(smooth_kernel), _ptr(depth), _ptr(out), _stream()), "st_finish_skeletons")
    return out


def repair_branches(nodes, row, length, parent, parent_repaired, level_off):
    """In-place on nodes [R,4]: writes every listed branch's connection point into its spare row."""
    lib = _lib.load()
    _req(nodes, F32, "nodes"); _req(row, I32, "row"); _req(length, I32, "len"); _req(parent, I32, "parent")
    _req(parent_repaired, U8, "parent_repaired"); _req(level_off, I32, "level_off")
    _count("repair")
    _lib.check(lib.st_repair_branches(_ptr(nodes), _ptr(row), _ptr(length), _ptr(parent), _ptr(parent_repaired), _ptr(level_off),
                                      level_off.shape[0] - 1, _stream()), "st_repair_branches")
