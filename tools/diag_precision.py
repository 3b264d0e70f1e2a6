import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import test_gpu_parity as T
from oracle import unet_ref as U
from smart_tree_b200.engine import SmartTreeEngine
for seed in (3, 7):
    sd = T._randomise(T._load("noble-elevator-58"), seed)
    feats, coords = T._cloud_inputs(2, 20000, 0.02)
    feats = np.random.default_rng(1).standard_normal(feats.shape).astype(np.float32)
    p = U.to_numpy_params(sd)
    tr = {}
    ref32 = U.forward(p, feats, coords, trace=tr)
    ref64 = U.forward(p, feats, coords, dtype=np.float64)
    # un-normalised direction head from the oracle
    for impl in ("fma", "tc", "auto"):
        eng = SmartTreeEngine(sd, device="cuda", conv_impl=impl)
        out = eng.forward(T._t(feats), T._t(coords))
        msg = []
        for k in ("radius", "direction", "class_l"):
            g = out[k].cpu().numpy()
            e = np.abs(g - ref64[k]).max(1)
            i = int(e.argmax())
            msg.append(f"{k}: max {e.max()/np.abs(ref64[k]).max():.2e} p99.9 {np.quantile(e,0.999)/np.abs(ref64[k]).max():.2e}")
        e32 = {k: np.abs(ref32[k]-ref64[k]).max()/np.abs(ref64[k]).max() for k in ("radius","direction","class_l")}
        print(seed, impl, " | ".join(msg), " oracle32-vs-64:", {k: f"{v:.1e}" for k,v in e32.items()})
