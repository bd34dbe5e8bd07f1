"""Top-down obstacle maps of a built scene: the reference's application/generate_obstacle_map.py:19-33 on the B200
engine.

    python -m avlmaps_b200.application.generate_obstacle_map --config-dir <reference>/config data_paths=default \\
        data_paths.avlmaps_data_dir=/data --out obstacles/ [--customize --text-encoder my_clip:encode --clip-dim 512]

Like `LangRobot.load_scene_map` (lang_robot.py:31-34) it creates the map class `map_config.map_type` names, loads the
scene and derives the occupancy obstacle map (map.py:79-95); `--customize` then runs `customize_obstacle_map`
(vlmap.py:127-156: open-vocabulary classes on the tensor cores, argmax, 2-D scatter, dilation) with the class lists of
`map_config`.  The reference shows both maps in cv2 windows; here they are written as PNG / .npy under `--out`.
The scenes live under `<avlmaps_data_dir>/vlmaps_dataset` in this one application (generate_obstacle_map.py:20)."""
from __future__ import annotations

import sys
from pathlib import Path
from typing import List, Optional

import numpy as np

from ..config import compose
from ..map import Map
from ._common import base_parser, load_callable


def main(argv: Optional[List[str]] = None) -> int:
    ap = base_parser("avlmaps_b200.application.generate_obstacle_map", "map_indexing_cfg.yaml", __doc__)
    ap.add_argument("--customize", action="store_true", help="also run customize_obstacle_map (needs a text encoder)")
    ap.add_argument("--text-encoder", default=None, help="module:function, list[str] -> (len, D) array")
    ap.add_argument("--clip-dim", type=int, default=512)
    ap.add_argument("--dataset-dir-name", default="vlmaps_dataset", help="generate_obstacle_map.py:20 reads vlmaps_dataset")
    ap.add_argument("--out", default=None, help="directory for obstacles.png / obstacles_custom.png (+ .npy)")
    args = ap.parse_intermixed_args(argv)
    config = compose(args.config_dir, args.config_name, args.overrides)
    data_dir = Path(config.data_paths.avlmaps_data_dir) / args.dataset_dir_name
    if not data_dir.is_dir():
        raise SystemExit(f"{data_dir} is not a directory (set data_paths.avlmaps_data_dir=... / --dataset-dir-name)")
    data_dirs = sorted(x for x in data_dir.iterdir() if x.is_dir())
    sid = int(config.scene_id)
    if not -len(data_dirs) <= sid < len(data_dirs):
        raise SystemExit(f"scene_id {sid} out of range: {len(data_dirs)} scene(s)")
    m = Map.create(config.map_config)                 # lang_robot.py:32
    if m.load_map(data_dirs[sid]) is False:           # lang_robot.py:33
        return 1
    m.generate_obstacle_map()                         # lang_robot.py:34
    maps = {"obstacles": m.obstacles_cropped}
    print(f"obstacle map: rows {m.rmin}..{m.rmax}, cols {m.cmin}..{m.cmax}, {int((m.obstacles_cropped == 0).sum())} occupied cells")
    if args.customize:
        enc = load_callable(args.text_encoder)
        if enc is not None:
            m.set_text_encoder(enc, args.clip_dim)
        m.customize_obstacle_map(config.map_config.potential_obstacle_names, config.map_config.obstacle_names, vis=False)
        maps["obstacles_custom"] = m.obstacles_new_cropped
        print(f"customised obstacle map: {int((m.obstacles_new_cropped == 0).sum())} occupied cells")
    if args.out:
        import cv2

        out = Path(args.out)
        out.mkdir(parents=True, exist_ok=True)
        for name, arr in maps.items():
            np.save(out / f"{name}.npy", np.asarray(arr))
            cv2.imwrite(str(out / f"{name}.png"), np.asarray(arr).astype(np.uint8) * 255)   # generate_obstacle_map.py:26
    return 0


if __name__ == "__main__":
    sys.exit(main())
