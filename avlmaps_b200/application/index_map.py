"""Index a built map: the object branch of the reference's application/index_map.py:23-149 on the B200 engine.

    python -m avlmaps_b200.application.index_map --config-dir <reference>/config data_paths=default \\
        data_paths.avlmaps_data_dir=/data --object sofa --object "potted plant"

Without `--object` it runs the reference's prompt loop (`1. object ... 6. exit`).  The text tower is CLIP ViT-B/32
through `VLMap._init_clip()` (index_map.py:28) when the `clip` package is installed, or `--text-encoder
module:function` (`list[str] -> (len, D)`) with `--clip-dim D`.  The sound / area / image branches need the reference's
AudioCLIP, CLIP ViT-L/14 and HLoc wrappers (attach them as `avlmap.sound_map / area_map / visual_map`, see
avlmaps_b200/map/avlmap.py); the open3d / matplotlib views of the reference are not reproduced -- the heat is printed
as a goal voxel and optionally saved with `--out`."""
from __future__ import annotations

import sys
from pathlib import Path
from typing import List, Optional

import numpy as np

from ..map import AVLMap
from ._common import base_parser, compose_from_args, load_callable

PROMPT = ("What do you want to index? (1. object, 2. sound, 3. area, 4. image, 5. show rgb point cloud, or 6. exit)\nInput: ")


def _report(avlmap: AVLMap, name: str, heat: np.ndarray, out_dir: Optional[Path]) -> None:
    goal = avlmap.get_max_pos_3d(heat)        # grid_pos[np.argmax(heat)], habitat_lang_robot.py:427-430
    print(f"{name!r}: {int((heat >= 1).sum())} target voxels, {int((heat > 0).sum())} with heat > 0, "
          f"goal voxel (row, col, height) = {np.asarray(goal).tolist()}")
    if out_dir is not None:
        out_dir.mkdir(parents=True, exist_ok=True)
        np.save(out_dir / f"heat_{name.replace(' ', '_')}.npy", heat)


def main(argv: Optional[List[str]] = None, input_fn=input) -> int:
    ap = base_parser("avlmaps_b200.application.index_map", "map_indexing_cfg.yaml", __doc__)
    ap.add_argument("--object", action="append", default=[], help="object name to index (repeatable); none = prompt loop")
    ap.add_argument("--text-encoder", default=None, help="module:function, list[str] -> (len, D) array")
    ap.add_argument("--clip-dim", type=int, default=512)
    ap.add_argument("--out", default=None, help="directory for heat_<name>.npy")
    args = ap.parse_args(argv)
    config, scene = compose_from_args(args)
    avlmap = AVLMap(config, data_dir=scene)                               # index_map.py:26
    if not avlmap.vlmap.load_map(scene):                                  # index_map.py:27 (prints and returns False)
        return 1
    enc = load_callable(args.text_encoder)
    if enc is not None:
        avlmap.vlmap.set_text_encoder(enc, args.clip_dim)
    else:
        avlmap.vlmap._init_clip()                                         # index_map.py:28
    decay = float(config.get("decay_rate", 0.01))                         # index_map.py:38 uses 0.01
    out_dir = Path(args.out) if args.out else None
    if args.object:
        for name in args.object:
            _report(avlmap, name, avlmap.index_object(name, decay_rate=decay), out_dir)
        return 0
    while True:
        choice = input_fn(PROMPT).strip()
        if choice == "1":
            name = input_fn("What is the object name you want to index?\nInput: ")
            _report(avlmap, name, avlmap.index_object(name, decay_rate=decay), out_dir)
        elif choice in ("2", "3", "4"):
            print("this modality needs the reference's SoundMap / AreaMap / VisualMap model wrappers attached to the AVLMap "
                  "(avlmaps_b200/map/avlmap.py); only the object modality runs from this command line")
        elif choice == "5":
            print(f"{avlmap.vlmap.grid_pos.shape[0]} voxels; the open3d viewer of the reference is not reproduced")
        elif choice == "6":
            return 0


if __name__ == "__main__":
    sys.exit(main())
