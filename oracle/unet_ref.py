"""Oracle: spconv semantics + Smart_Tree.forward on the CPU (numpy, fp32 or fp64).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Parity unpinned at the spconv boundary.

Follows
  /root/reference/smart_tree/model/model.py:77-87          Smart_Tree.forward
  /root/reference/smart_tree/model/model_blocks.py:8-38    SubMConvBlock
  /root/reference/smart_tree/model/model_blocks.py:41-104  Encoder/DecoderBlock
  /root/reference/smart_tree/model/model_blocks.py:107-156 ResBlock
  /root/reference/smart_tree/model/model_blocks.py:159-243 UBlock
  /root/reference/smart_tree/model/model_blocks.py:246-320 SparseFC / MLP heads
  /root/reference/smart_tree/dataset/dataset.py:199-216    PointToVoxel call site
and SURVEY.md Appendix B (B1-B6) for the third-party semantics.
"""
from __future__ import annotations

import numpy as np

OFFSETS = np.array([(kz - 1, ky - 1, kx - 1) for kz in range(3) for ky in range(3) for kx in range(3)],
                   dtype=np.int64)  # k = (kz*3+ky)*3+kx  <->  weight[:, kz, ky, kx, :]


# ----------------------------------------------------------------------------- keys
def pack_keys(coords: np.ndarray) -> np.ndarray:
    """(b,z,y,x) int -> one int64 key, 16 bits per field, +1 bias so that -1 is legal."""
    c = coords.astype(np.int64)
    return (c[:, 0] << 48) | ((c[:, 1] + 1) << 32) | ((c[:, 2] + 1) << 16) | (c[:, 3] + 1)


def _lookup(sorted_keys, order, q):
    pos = np.searchsorted(sorted_keys, q)
    pos[pos >= len(sorted_keys)] = 0
    hit = sorted_keys[pos] == q if len(sorted_keys) else np.zeros(len(q), bool)
    return np.where(hit, order[pos] if len(sorted_keys) else 0, -1)


# ----------------------------------------------------------------------------- B6
def point_to_voxel(points: np.ndarray, vsize: float, lo: np.ndarray, hi: np.ndarray):
    """spconv PointToVoxel (CPU), max_num_points_per_voxel=1.

    points [N,F] f32 (xyz first), lo/hi = coors_range_xyz.  Returns
    voxels[M,F] (first point of each voxel, verbatim), indices_zyx[M,3] i32 in
    first-appearance order, pc_voxel_id[N] (-1 = dropped)."""
    points = np.ascontiguousarray(points, np.float32)
    vs = np.float32(vsize)
    lo = np.asarray(lo, np.float32)
    hi = np.asarray(hi, np.float32)
    # std::round on float: half away from zero (q >= 0 here); evaluated exactly via trunc + fraction
    q = (hi - lo) / vs
    t = np.trunc(q)
    grid = (t + ((q - t) >= np.float32(0.5))).astype(np.int64)
    c = np.floor((points[:, :3] - lo) / vs).astype(np.int64)          # fp32 sub, fp32 div, floor
    ok = np.all((c >= 0) & (c < grid), axis=1)
    lin = (c[:, 2] * (grid[1] + 1) + c[:, 1]) * (grid[0] + 1) + c[:, 0]
    lin_ok = lin[ok]
    idx_ok = np.nonzero(ok)[0]
    uniq, first, inv = np.unique(lin_ok, return_index=True, return_inverse=True)
    appearance = np.argsort(first, kind="stable")                       # voxel id = order of first appearance
    rank = np.empty(len(uniq), np.int64)
    rank[appearance] = np.arange(len(uniq))
    pc_voxel_id = np.full(len(points), -1, np.int64)
    pc_voxel_id[idx_ok] = rank[inv]
    rep = idx_ok[first[appearance]]
    voxels = points[rep]
    indices_zyx = c[rep][:, ::-1].astype(np.int32)
    return voxels, indices_zyx, pc_voxel_id


# ----------------------------------------------------------------------------- B2-B4
def declared_spatial_shape(coords: np.ndarray) -> np.ndarray:
    """spatial_shape as the reference declares it: max(coords) per axis, NOT max + 1
    (/root/reference/smart_tree/model/sparse.py:15-19)."""
    return coords[:, 1:].max(0).astype(np.int64) if len(coords) else np.zeros(3, np.int64)


def subm_map(coords: np.ndarray, spatial_shape=None) -> np.ndarray:
    """nbr[27,N]: row of the active voxel at coords[i] + OFFSETS[k], or -1.
    spatial_shape = None: unbounded grid (the default).  Otherwise (strict_spconv_bounds) a
    neighbour LOCATION with any coordinate >= spatial_shape is not queried, as spconv bound-checks it -- voxels on
    the max faces are invisible as neighbours; the centre tap is never clipped.  Only the clip is modelled, not the
    index aliasing spconv would suffer for those voxels (implementation detail, unverifiable here)."""
    keys = pack_keys(coords)
    order = np.argsort(keys, kind="stable")
    sk = keys[order]
    out = np.empty((27, len(coords)), np.int64)
    for k, (dz, dy, dx) in enumerate(OFFSETS):
        q = coords.astype(np.int64).copy()
        q[:, 1] += dz
        q[:, 2] += dy
        q[:, 3] += dx
        out[k] = _lookup(sk, order, pack_keys(q))
        if spatial_shape is not None and k != 13:
            out[k][np.any(q[:, 1:] >= np.asarray(spatial_shape, np.int64), axis=1)] = -1
    return out


def strided_maps(coords: np.ndarray, out_shape=None):
    """SparseConv3d(k=3,s=2,p=1) index generation.

    Returns out_coords [M,4] sorted by (b,z,y,x), down[27,M] (input row feeding output
    row o through tap k, p = 2o-1+k) and up[27,N] (output row fed by input row p through
    tap k, o = (p+1-k)/2) -- the latter is what SparseInverseConv3d gathers from (B4).
    out_shape (strict_spconv_bounds): outputs with any coordinate >= out_shape = (in_shape - 1) // 2 + 1 are not
    created (B3: 0 <= o < out_shape)."""
    c = coords.astype(np.int64)
    n = len(c)
    cand = []
    up_o = np.empty((27, n, 4), np.int64)
    up_ok = np.empty((27, n), bool)
    for k, (dz, dy, dx) in enumerate(OFFSETS):  # tap index per axis = d+1
        t = c[:, 1:] + 1 - (np.array([dz, dy, dx]) + 1)
        ok = np.all((t % 2) == 0, axis=1)
        if out_shape is not None:
            ok &= np.all(t // 2 < np.asarray(out_shape, np.int64), axis=1)
        o = np.concatenate([c[:, :1], t // 2], axis=1)
        up_o[k], up_ok[k] = o, ok
        cand.append(o[ok])
    cand = np.concatenate(cand) if n else np.zeros((0, 4), np.int64)
    ck = pack_keys(cand)
    uk, first = np.unique(ck, return_index=True)
    out_coords = cand[first]
    m = len(out_coords)
    ordr = np.arange(m)
    up = np.full((27, n), -1, np.int64)
    down = np.full((27, m), -1, np.int64)
    for k in range(27):
        rows = np.nonzero(up_ok[k])[0]
        o_rows = _lookup(uk, ordr, pack_keys(up_o[k][rows]))
        up[k, rows] = o_rows
        down[k, o_rows] = rows
    return out_coords.astype(np.int32), down, up


def gather_conv(feat: np.ndarray, weight: np.ndarray, nbr: np.ndarray, n_out: int) -> np.ndarray:
    """out[i] = sum_k W[:,k,:] . feat[nbr[k,i]]   (taps accumulated k = 0..26 in order).
    weight is the spconv layout [Cout,3,3,3,Cin]."""
    cout, cin = weight.shape[0], weight.shape[-1]
    w = weight.reshape(cout, 27, cin).astype(feat.dtype)
    out = np.zeros((n_out, cout), feat.dtype)
    for k in range(27):
        rows = np.nonzero(nbr[k] >= 0)[0]
        if len(rows):
            out[rows] += feat[nbr[k, rows]] @ w[:, k, :].T
    return out


def linear(feat, weight):
    """1x1 sub-manifold conv fast path / nn.Linear: features @ W.view(out,in).T"""
    w = weight.reshape(weight.shape[0], -1).astype(feat.dtype)
    return feat @ w.T


def batchnorm(x, p, prefix, eps):
    w = p[prefix + ".weight"].astype(x.dtype)
    b = p[prefix + ".bias"].astype(x.dtype)
    m = p[prefix + ".running_mean"].astype(x.dtype)
    v = p[prefix + ".running_var"].astype(x.dtype)
    return (x - m) / np.sqrt(v + x.dtype.type(eps)) * w + b


def relu(x):
    return np.maximum(x, 0)


# ----------------------------------------------------------------------------- network
def unet_depth(params) -> int:
    d, pre = 0, "UNet."
    while pre + "Head.sequence.0.weight" in params:
        d, pre = d + 1, pre + "U."
    return d


class LevelMaps:
    """Index structures of one UNet level, built once and shared by all its convs."""

    def __init__(self, coords, spatial_shape=None):
        self.coords = coords
        self.spatial_shape = spatial_shape
        self.nbr = subm_map(coords, spatial_shape)
        self.down = self.up = None
        self.child = None


def build_levels(coords: np.ndarray, depth: int, spatial_shape=None):
    """spatial_shape = declared_spatial_shape(coords) reproduces spconv's bound checks (strict_spconv_bounds); each level
    hands out_shape = (shape - 1) // 2 + 1 to the next, as spconv does.  None = unbounded."""
    levels = [LevelMaps(coords, spatial_shape)]
    for _ in range(depth - 1):
        cur = levels[-1]
        out_shape = None if cur.spatial_shape is None else (np.asarray(cur.spatial_shape, np.int64) - 1) // 2 + 1
        oc, down, up = strided_maps(cur.coords, out_shape)
        cur.down, cur.up = down, up
        levels.append(LevelMaps(oc, out_shape))
    return levels


def _resblock(x, p, pre, nbr, eps):
    n = len(x)
    y = gather_conv(x, p[pre + "sequence.0.weight"], nbr, n)
    y = relu(batchnorm(y, p, pre + "sequence.1", eps))
    y = gather_conv(y, p[pre + "sequence.3.weight"], nbr, n)
    y = batchnorm(y, p, pre + "sequence.4", eps)
    idk = pre + "identity.0.weight"
    ident = linear(x, p[idk]) if idk in p else x
    return relu(y + ident)


def _ublock(x, p, pre, levels, li, eps, trace):
    lv = levels[li]
    x = _resblock(x, p, pre + "Head.", lv.nbr, eps)
    if trace is not None:
        trace[pre + "Head"] = x
    if pre + "Encode.sequence.0.weight" not in p:
        return x
    skip = x
    nxt = levels[li + 1]
    y = gather_conv(x, p[pre + "Encode.sequence.0.weight"], lv.down, len(nxt.coords))
    y = relu(batchnorm(y, p, pre + "Encode.sequence.1", eps))
    if trace is not None:
        trace[pre + "Encode"] = y
    y = _ublock(y, p, pre + "U.", levels, li + 1, eps, trace)
    y = gather_conv(y, p[pre + "Decode.sequence.0.weight"], lv.up, len(lv.coords))
    y = relu(batchnorm(y, p, pre + "Decode.sequence.1", eps))
    if trace is not None:
        trace[pre + "Decode"] = y
    y = np.concatenate([skip, y], axis=1)
    y = _resblock(y, p, pre + "Tail.", lv.nbr, eps)
    if trace is not None:
        trace[pre + "Tail"] = y
    return y


def _head(x, p, pre, eps):
    i = 0
    while True:
        wk = f"{pre}sequence.{i}.weight"
        x = linear(x, p[wk])
        bk = f"{pre}sequence.{i}.bias"
        if bk in p and p[wk].ndim == 2:          # HEAD-code MLP: nn.Linear(bias=True)
            x = x + p[bk].astype(x.dtype)
        if f"{pre}sequence.{i + 1}.weight" not in p:
            return x
        x = relu(batchnorm(x, p, f"{pre}sequence.{i + 1}", eps))
        i += 3


def to_numpy_params(state_dict) -> dict:
    out = {}
    for k, v in state_dict.items():
        if k.endswith("num_batches_tracked"):
            continue
        out[k] = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
    return out


def forward(params: dict, features: np.ndarray, coords: np.ndarray, eps: float = 1e-4,
            dtype=np.float32, trace: dict | None = None, levels=None):
    """Smart_Tree.forward (model.py:77-87).  features [N,Cin], coords [N,4]=(b,z,y,x).
    Returns dict radius[N,1] (log radius), direction[N,3] (unit), class_l[N,k] logits."""
    p = params
    x = np.ascontiguousarray(features, dtype)
    if levels is None:
        levels = build_levels(np.asarray(coords), unet_depth(p))
    x = relu(batchnorm(linear(x, p["input_conv.sequence.0.weight"]), p, "input_conv.sequence.1", eps))
    if trace is not None:
        trace["input_conv"] = x
    x = _ublock(x, p, "UNet.", levels, 0, eps, trace)
    radius = _head(x, p, "radius_head.", eps)
    direction = _head(x, p, "direction_head.", eps)
    class_l = _head(x, p, "class_head.", eps)
    # F.normalize(p=2, dim=1, eps=1e-12):  v / max(||v||, eps)
    nrm = np.sqrt((direction * direction).sum(1, keepdims=True))
    direction = direction / np.maximum(nrm, dtype(1e-12))
    return {"radius": radius, "direction": direction, "class_l": class_l}


def algorithmic_bytes(level_sizes) -> int:
    """UNet compulsory fp32 feature traffic = 964 N0 + 1152 N1 + 2304 N2 + 1792 N3
    (generalised: heads counted layer-wise as in the survey)."""
    n = list(level_sizes)
    if len(n) == 4:
        return 964 * n[0] + 1152 * n[1] + 2304 * n[2] + 1792 * n[3]
    raise ValueError("closed form only given for depth 4")
