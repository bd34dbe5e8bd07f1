"""Shared plumbing of the applications: config composition and the data-directory convention."""
from __future__ import annotations

import argparse
import importlib
import os
from pathlib import Path
from typing import Callable, List, Optional, Tuple

from ..config import Config, compose


def load_callable(spec: Optional[str]) -> Optional[Callable]:
    """`package.module:attribute` -> the object (the pixel / text encoders are the caller's models)."""
    if not spec:
        return None
    if ":" not in spec:
        raise SystemExit(f"expected module:attribute, got {spec!r}")
    mod, attr = spec.split(":", 1)
    obj = importlib.import_module(mod)
    for part in attr.split("."):
        obj = getattr(obj, part)
    return obj


def base_parser(prog: str, config_name: str, description: str = "") -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog=prog, description=description, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config-dir", default=os.environ.get("AVL_CONFIG_DIR", "config"),
                    help="the reference's config/ directory (default: $AVL_CONFIG_DIR or ./config)")
    ap.add_argument("--config-name", default=config_name)
    ap.add_argument("overrides", nargs="*", help="Hydra-style overrides: key.path=value, +key=value, group=option, ~key")
    return ap


def scene_dirs(config: Config) -> List[Path]:
    """`sorted(dirs of <avlmaps_data_dir>/avlmaps_dataset)` (reference create_map.py:14-15, index_map.py:24-25)."""
    data_dir = Path(config.data_paths.avlmaps_data_dir) / "avlmaps_dataset"
    if not data_dir.is_dir():
        raise SystemExit(f"{data_dir} is not a directory (set data_paths.avlmaps_data_dir=...)")
    return sorted(x for x in data_dir.iterdir() if x.is_dir())


def compose_from_args(args) -> Tuple[Config, Path]:
    config = compose(args.config_dir, args.config_name, args.overrides)
    dirs = scene_dirs(config)
    sid = int(config.scene_id)
    if not -len(dirs) <= sid < len(dirs):
        raise SystemExit(f"scene_id {sid} out of range: {len(dirs)} scene(s) under the data directory")
    return config, dirs[sid]
