"""Micro-benchmark of the gather-conv kernels on the level shapes of the bench workload
(one synthetic 1M-point tree -> ~495k voxels).  Usage: python tools/conv_micro.py [fma|tc] [cin cout level]"""
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
levels = build_levels(coords, 4, morton=MORTON)
print("levels", [l.n for l in levels], "z-order" if MORTON else "input order")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cases = [(8, 8, 0), (16, 8, 0), (16, 16, 1), (32, 16, 1), (32, 32, 2), (64, 32, 2), (64, 64, 3)]
if only:
    cases = [only]
for cin, cout, li in cases:
    lv = levels[li]
    x = torch.randn(lv.n, cin, device=dev)
    w = torch.randn(27, cin, cout, device=dev) / (27 * cin) ** 0.5
    wtc = ops.conv_tc_prepare(w) if impl == "tc" else None
    out = torch.empty(lv.n, cout, device=dev)
    for _ in range(3):
        ops.conv_gather(x, lv.nbr, w, lv.n, out=out, relu=True, impl=impl, weight_tc=wtc)
    ts = []
    for _ in range(10):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.conv_gather(x, lv.nbr, w, lv.n, out=out, relu=True, impl=impl, weight_tc=wtc)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    us = float(np.median(ts))
    byt = 4 * lv.n * (cin + cout)
    print(f"{impl} {cin:3d}->{cout:3d} L{li} n={lv.n:7d}  {us:8.1f} us  {byt / us / 1e3:7.1f} GB/s  ({byt / us / 1e3 / 6525.2 * 100:5.2f}% of measured HBM peak)")

if only and impl == "tc":
    import ctypes as C
    from smart_tree_b200 import _lib
    lib = _lib.load()
    buf = (C.c_longlong * 320)()
    lib.st_debug_tc_trace.argtypes = [C.c_void_p]
    lib.st_debug_tc_trace(buf)
    t = np.array(buf[:]).reshape(5, 64)
    t0 = t[0, 0]
    names = ["prod: a_empty ok", "prod: arrived   ", "mma : a_full ok ", "mma : committed ", "mma : b_full ok "]
    for g in range(0, 24):
        print(g, "  ".join(f"{names[k].strip()}={t[k, g] - t0:7d}" for k in (0, 1, 4, 2, 3)))
