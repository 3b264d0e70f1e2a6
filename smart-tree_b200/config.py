"""Tiny recursive `_target_` instantiator: hydra / omegaconf are not installed in the target
image, and the reference only uses `hydra.utils.instantiate(cfg.pipeline)` plus `+path=` /
`+directory=` overrides (/root/reference/smart_tree/cli.py:10-26)."""
from __future__ import annotations

import importlib
import os

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))

# the reference's own class paths resolve to the B200 implementations
_ALIASES = {"smart_tree.": "smart_tree_b200."}


def _resolve(target: str):
    for old, new in _ALIASES.items():
        if target.startswith(old):
            target = new + target[len(old):]
    mod, _, name = target.rpartition(".")
    return getattr(importlib.import_module(mod), name)


def instantiate(node):
    if isinstance(node, dict):
        kw = {k: instantiate(v) for k, v in node.items() if k not in ("_target_", "_partial_")}
        if "_target_" in node:
            return _resolve(node["_target_"])(**kw)
        return kw
    if isinstance(node, (list, tuple)):
        return [instantiate(v) for v in node]
    return node


def load_config(path=None, overrides=()):
    with open(path or os.path.join(HERE, "conf", "pipeline.yaml")) as f:
        cfg = yaml.safe_load(f)
    for ov in overrides:
        key, _, val = ov.lstrip("+").partition("=")
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = yaml.safe_load(val)
    mi = cfg.get("pipeline", {}).get("model_inference", {})
    for k in ("model_path", "weights_path"):
        if k in mi and not os.path.isabs(str(mi[k])):
            cand = os.path.join(HERE, str(mi[k]).replace("smart_tree/", ""))
            if os.path.exists(cand) or k == "model_path":
                mi[k] = cand
    return cfg
