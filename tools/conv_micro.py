"""Micro-benchmark of the gather-conv kernels on the level shapes of the bench workload
(one synthetic 1M-point tree -> ~495k voxels).  Usage: python tools/conv_micro.py [fma|tc|tp] [cin cout level]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from smart_tree_b200 import ops, synth
from smart_tree_b200.data_types.cloud import Cloud
from smart_tree_b200.dataset.augmentations import CentreCloud
from smart_tree_b200.dataset.dataset import SingleTreeInference
from smart_tree_b200.engine import build_levels

impl = sys.argv[1] if len(sys.argv) > 1 else "tc"
only = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else None
dev = torch.device("cuda:0")
tr = synth.make_tree(0, 1_000_000)
cloud = CentreCloud()(Cloud(xyz=torch.from_numpy(tr.xyz).to(dev), rgb=torch.from_numpy(tr.rgb).to(dev)))
bb = SingleTreeInference(cloud, 0.01, 4, 0.4).voxelize_all()
MORTON = bool(int(os.environ.get("ST_MORTON", "1")))
coords = bb.coords.contiguous()
if MORTON:
    coords = coords[ops.morton_perm(coords).long()].contiguous()
levels = build_levels(coords, 4, morton=MORTON, inverse_plan=(impl in ("tcinv", "fmainv")))
print("levels", [l.n for l in levels], "z-order" if MORTON else "input order")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cases = [(8, 8, 0), (16, 8, 0), (8, 16, 0), (16, 16, 1), (32, 16, 1), (16, 32, 1), (32, 32, 2), (64, 32, 2), (64, 64, 3)]
if only:
    cases = [only]
if os.environ.get("ST_TC_DBG"):
    from smart_tree_b200 import _lib as _l
    _l.load().st_debug_tc_set(int(os.environ["ST_TC_DBG"]))
if impl in ("tcinv", "fmainv"):
    cases = [c for c in [(16, 8, 0), (32, 16, 1), (64, 32, 2)] if only is None or c == only]
for cin, cout, li in cases:
    lv = levels[li]
    if impl in ("tcinv", "fmainv"):
        x = torch.randn(levels[li + 1].n, cin, device=dev)
        w = torch.randn(27, cin, cout, device=dev) / (4 * cin) ** 0.5
        wtc = ops.conv_tc_prepare(w)
        out = torch.empty(lv.n, cout, device=dev)
        if impl == "tcinv":
            run = lambda: ops.conv_gather_tc_inv(x, lv.inverse_plan(), wtc, 27, cin, cout, lv.n, out=out, relu=True)
        else:
            run = lambda: ops.conv_gather_inv(x, lv.inverse_plan(), w, lv.n, out=out, relu=True)
        for _ in range(3):
            run()
        ts = []
        for _ in range(10):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        us = float(np.median(ts))
        byt = 4 * (lv.n * cout + levels[li + 1].n * cin)
        print(f"{impl} {cin:3d}->{cout:3d} L{li} n={lv.n:7d}  {us:8.1f} us  {byt / us / 1e3:7.1f} GB/s")
        continue
    if impl == "brick":
        if not ops.conv_brick_supported(27, cin, cout):
            continue
        x = torch.randn(lv.n, cin, device=dev)
        w = torch.randn(27, cin, cout, device=dev) / (27 * cin) ** 0.5
        out = torch.empty(lv.n, cout, device=dev)
        ts = []
        for _ in range(5):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); plan = ops.brick_plan(lv.coords, lv.table); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        nb, nh, st = ops.brick_plan_info(plan, lv.n)
        print(f"  brick plan: {nb} bricks ({lv.n / nb:.1f} voxels each), {nh} halo entries ({nh / lv.n:.2f} per voxel), status {st}, build {float(np.median(ts)):.1f} us")
        ref = ops.conv_gather(x, lv.nbr, w, lv.n, relu=True, impl="fma")
        got = ops.conv_brick(x, plan, w, lv.n, relu=True)
        print(f"  max |brick - map kernel| = {float((ref - got).abs().max()):.3e}")
        for _ in range(3):
            ops.conv_brick(x, plan, w, lv.n, out=out, relu=True)
        ts = []
        for _ in range(10):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.conv_brick(x, plan, w, lv.n, out=out, relu=True); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        us = float(np.median(ts))
        byt = 4 * lv.n * (cin + cout)
        print(f"{impl} {cin:3d}->{cout:3d} L{li} n={lv.n:7d}  {us:8.1f} us  {byt / us / 1e3:7.1f} GB/s  ({byt / us / 1e3 / 6525.2 * 100:5.2f}% of measured HBM peak)")
        continue
    x = torch.randn(lv.n, cin, device=dev)
    w = torch.randn(27, cin, cout, device=dev) / (27 * cin) ** 0.5
    wtc = ops.conv_tc_prepare(w) if impl in ("tc", "tp") else None
    plan = ops.conv_plan_build(lv.nbr, lv.n) if impl == "tp" else None
    if plan is not None:
        import numpy as _np
        nb = (lv.n + 127) // 128
        hdr = plan[:nb * 48].view(torch.int32).reshape(nb, 12).cpu().numpy()
        print(f"  plan: {nb} tiles, {int(hdr[:, 0].sum())} split, distinct rows mean {hdr[hdr[:, 0] == 0, 1].mean():.0f} max {hdr[:, 1:9].max()}")
    out = torch.empty(lv.n, cout, device=dev)
    for _ in range(3):
        ops.conv_gather(x, lv.nbr, w, lv.n, out=out, relu=True, impl=impl, weight_tc=wtc, plan=plan)
    ts = []
    for _ in range(10):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.conv_gather(x, lv.nbr, w, lv.n, out=out, relu=True, impl=impl, weight_tc=wtc, plan=plan)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    us = float(np.median(ts))
    byt = 4 * lv.n * (cin + cout)
    print(f"{impl} {cin:3d}->{cout:3d} L{li} n={lv.n:7d}  {us:8.1f} us  {byt / us / 1e3:7.1f} GB/s  ({byt / us / 1e3 / 6525.2 * 100:5.2f}% of measured HBM peak)")

if only and impl == "tcinv":
    import ctypes as C
    from smart_tree_b200 import _lib
    lib = _lib.load()
    buf = (C.c_longlong * 1024)()
    lib.st_debug_tc_trace.argtypes = [C.c_void_p]
    lib.st_debug_tc_trace(buf)
    t = np.array(buf[:]).reshape(8, 128)
    t0 = t[5, 0]
    nt = int((t[5] > 0).sum())
    print("tile starts (cycles):", (t[5, :nt] - t0).tolist())
    names = {6: "P.start", 7: "P.loads", 0: "P.slot_ok", 1: "P.arrived", 4: "M.b_full", 2: "M.a_full", 3: "M.commit"}
    for g in range(0, 24):
        print(g, "  ".join(f"{names[k]}={t[k, g] - t0:6d}" for k in (6, 7, 0, 1, 4, 2, 3)))
if only and impl in ("tc", "tp"):
    import ctypes as C
    from smart_tree_b200 import _lib
    lib = _lib.load()
    buf = (C.c_longlong * 1024)()
    lib.st_debug_tc_trace.argtypes = [C.c_void_p]
    lib.st_debug_tc_trace(buf)
    t = np.array(buf[:]).reshape(8, 128)
    t0 = t[6, 0]
    ns = -(-27 * only[0] // 32)
    print("traced CTA", os.environ.get("ST_TC_DBG", "0"), "stages per tile", ns)
    print(f"kernel span of this CTA: {t[5, 122] - t[5, 120]} ns = {t[5, 123] - t[5, 121]} cycles; entry -> first P.start {t0 - t[5, 121]} cycles; entry -> after TMEM alloc {t[5, 124] - t[5, 121]}; last arrive -> exit {t[5, 123] - t[1, :].max()}")
    for tile in range(0, 128 // ns):
        g0, g1 = tile * ns, tile * ns + ns - 1
        if t[6, g0] == 0:
            break
        st = np.diff(t[6, g0:g1 + 1])
        print(f"tile {tile}: start {t[6, g0] - t0:7d}  f_wait {t[5, g0] - t0:7d} f_ok {t[5, g0 + 1] - t0:7d}  last arrive {t[1, g1] - t0:7d}  "
              f"mma first b_full {t[4, g0] - t0:7d} a_full {t[2, g0] - t0:7d} last commit {t[3, g1] - t0:7d}  stage min/med/max {st.min()}/{int(np.median(st))}/{st.max()}")
