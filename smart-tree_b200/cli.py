"""`run-smart-tree` entry point (/root/reference/smart_tree/cli.py:10-26, pyproject.toml:36):
    python -m smart_tree_b200.cli +path=tree.npz        | +directory=clouds/
Accepts hydra-style `key=value` / `+key=value` overrides."""
import os
import sys
from pathlib import Path

from .config import instantiate, load_config


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    cfg = load_config(overrides=[a for a in argv if "=" in a])
    pipeline = instantiate(cfg["pipeline"])
    if "path" in cfg:
        pipeline.process_cloud(Path(cfg["path"]))
    elif "directory" in cfg:
        for p in sorted(os.listdir(cfg["directory"])):
            pipeline.process_cloud(Path(f"{cfg['directory']}/{p}"))
    else:
        print("Please supply a path or directory to point clouds.")
    return pipeline


if __name__ == "__main__":
    main()
