"""Pin the oracle's spconv restatement (SURVEY §8c "Oracle self-validation" (1),(2)):
sub-manifold / strided / inverse sparse conv == dense torch conv sampled at active sites."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import unet_ref as U


def _random_sparse(rng, n, extent, batch=2, cin=5):
    c = np.unique(np.concatenate([rng.integers(0, batch, (n, 1)), rng.integers(0, extent, (n, 3))], 1), axis=0)
    rng.shuffle(c)
    f = rng.standard_normal((len(c), cin)).astype(np.float32)
    return c.astype(np.int32), f


def _dense(c, f, batch, shape):
    d = torch.zeros(batch, f.shape[1], *shape, dtype=torch.float64)
    d[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]] = torch.from_numpy(f).double()
    return d


@pytest.mark.parametrize("seed", [0, 1])
def test_subm_conv_matches_dense(seed):
    rng = np.random.default_rng(seed)
    c, f = _random_sparse(rng, 400, 9)
    w = rng.standard_normal((7, 3, 3, 3, 5)).astype(np.float32)
    out = U.gather_conv(f.astype(np.float64), w, U.subm_map(c), len(c))
    dense = F.conv3d(_dense(c, f, 2, (9, 9, 9)), torch.from_numpy(w).double().permute(0, 4, 1, 2, 3), padding=1)
    ref = dense[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]].numpy()
    np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("extent", [8, 9])
def test_strided_and_inverse_conv_match_dense(extent):
    rng = np.random.default_rng(extent)
    c, f = _random_sparse(rng, 300, extent)
    oc, down, up = U.strided_maps(c)
    w = rng.standard_normal((6, 3, 3, 3, 5)).astype(np.float32)
    y = U.gather_conv(f.astype(np.float64), w, down, len(oc))
    # dense grid one cell larger than the data so that no output is clipped (oracle = unbounded coords, C-3)
    E = extent + 1
    dense = F.conv3d(_dense(c, f, 2, (E,) * 3), torch.from_numpy(w).double().permute(0, 4, 1, 2, 3),
                     stride=2, padding=1)
    ref = dense[oc[:, 0], :, oc[:, 1], oc[:, 2], oc[:, 3]].numpy()
    np.testing.assert_allclose(y, ref, rtol=1e-10, atol=1e-10)
    # active outputs == every dense output cell touched by an active input
    occ = F.conv3d(_dense(c, np.ones((len(c), 1), np.float32), 2, (E,) * 3),
                   torch.ones(1, 1, 3, 3, 3, dtype=torch.float64), stride=2, padding=1)
    assert int((occ > 0).sum()) == len(oc)
    assert np.all(np.diff(U.pack_keys(oc)) > 0)          # sorted, unique
    # inverse conv == conv_transpose3d sampled at the encoder-input sites (B4)
    w2 = rng.standard_normal((4, 3, 3, 3, 6)).astype(np.float32)   # [out=4, k, in=6]
    z = U.gather_conv(y, w2, up, len(c))
    od = tuple(dense.shape[2:])
    yd = _dense(oc, y.astype(np.float32), 2, od)
    yd[oc[:, 0], :, oc[:, 1], oc[:, 2], oc[:, 3]] = torch.from_numpy(y)
    opad = [E - ((int(o) - 1) * 2 - 2 + 3) for o in od]
    zt = F.conv_transpose3d(yd, torch.from_numpy(w2).double().permute(4, 0, 1, 2, 3), stride=2, padding=1,
                            output_padding=opad)
    ref2 = zt[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]].numpy()
    np.testing.assert_allclose(z, ref2, rtol=1e-9, atol=1e-9)


def test_batchnorm_closed_form():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((50, 6)).astype(np.float32)
    bn = torch.nn.BatchNorm1d(6, eps=1e-4).eval()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(6) + 0.5); bn.bias.copy_(torch.randn(6))
        bn.running_mean.copy_(torch.randn(6)); bn.running_var.copy_(torch.rand(6) + 0.1)
    p = {"b." + k: v.numpy() for k, v in bn.state_dict().items() if v.ndim}
    np.testing.assert_allclose(U.batchnorm(x, p, "b", 1e-4), bn(torch.from_numpy(x)).detach().numpy(), rtol=2e-6, atol=2e-6)


def test_point_to_voxel_first_point_wins():
    pts = np.array([[0.005, 0.0, 0.0, 1], [0.006, 0.001, 0.0, 2], [0.031, 0.0, 0.0, 3], [0.0349, 0.0, 0.0, 4],
                    [0.0, 0.02, 0.021, 5]], np.float32)
    lo, hi = pts[:, :3].min(0), pts[:, :3].max(0)
    vox, idx, pid = U.point_to_voxel(pts, 0.01, lo, hi)
    # grid = round((hi-lo)/vs) = (3,2,2): x=0.0349 -> c=3 >= 3 dropped; z=0.021 -> c=2 >= 2 dropped
    assert pid.tolist() == [0, 0, -1, -1, -1] or pid.tolist()[0:2] == [0, 0]
    assert vox[0, 3] == 1
    assert idx.shape[1] == 3
    rng = np.random.default_rng(0)
    p = rng.uniform(0, 1, (2000, 6)).astype(np.float32)
    vox, idx, pid = U.point_to_voxel(p, 0.1, p[:, :3].min(0), p[:, :3].max(0))
    # python restatement of the sequential loop
    seen, order = {}, []
    lo32, vs = p[:, :3].min(0), np.float32(0.1)
    grid = np.round((p[:, :3].max(0) - lo32) / vs)
    for i in range(len(p)):
        c = np.floor((p[i, :3] - lo32) / vs)
        if np.any(c < 0) or np.any(c >= grid):
            assert pid[i] == -1
            continue
        key = tuple(int(v) for v in c)
        if key not in seen:
            seen[key] = len(order); order.append(i)
        assert pid[i] == seen[key]
    assert np.array_equal(vox, p[order])
    assert np.array_equal(idx[:, ::-1], np.floor((p[order, :3] - lo32) / vs).astype(np.int32))


def test_inverse_conv_taps_are_fixed_by_parity_class():
    """What the parity-sorted decoder conv (st_inverse_plan / k_conv_tc<.., MASKED>) relies on: a fine voxel p receives
    from a coarse voxel through tap k only if p = 2o - 1 + k, so per axis k = 1 for even p and k in {0, 2} for odd p.
    Hence at most 8 of the 27 taps of the oracle's `up` map are non-empty for any voxel, they depend on the parity class
    only, and the all-even class uses the centre tap alone."""
    rng = np.random.default_rng(5)
    c, _ = _random_sparse(rng, 3000, 20)
    _, _, up = U.strided_maps(c)
    assert (up >= 0).any(axis=0).all()                                   # every fine voxel has at least one parent
    cls = ((c[:, 1] & 1) << 2) | ((c[:, 2] & 1) << 1) | (c[:, 3] & 1)
    allowed = np.zeros((8, 27), bool)
    for m in range(8):
        for k in range(27):
            kz, ky, kx = k // 9, (k // 3) % 3, k % 3
            ok = all((kk == 1) if not (m >> sh) & 1 else (kk in (0, 2)) for kk, sh in ((kz, 2), (ky, 1), (kx, 0)))
            allowed[m, k] = ok
    assert allowed.sum(1).tolist() == [1, 2, 2, 4, 2, 4, 4, 8] and allowed[0, 13]
    used = up >= 0
    assert not (used & ~allowed[cls].T).any()
    assert used.sum(0).max() <= 8 and abs(allowed.sum() / 8 - 27 / 8) < 1e-12
