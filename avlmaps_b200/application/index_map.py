"""Index a built map: the object branch of the reference's application/index_map.py:23-149 on the B200 engine.

    python -m avlmaps_b200.application.index_map --config-dir <reference>/config data_paths=default \\
        data_paths.avlmaps_data_dir=/data --object sofa --object "potted plant"

Without `--object / --area / --sound` it runs the reference's prompt loop (`1. object ... 6. exit`).  The text tower is
CLIP ViT-B/32 through `VLMap._init_clip()` (index_map.py:28) when the `clip` package is installed, or `--text-encoder
module:function` (`list[str] -> (len, D)`) with `--clip-dim D`.  The area branch (index_map.py:88-90) runs once
`--area-text-encoder module:function` names the CLIP ViT-L/14 text tower (the scene's `area_map/clip_sparse_map.h5df`
is loaded); the sound branch (index_map.py:41-87) once `--sound-text-encoder`, `--sound-categories a,b,c` and
`--sound-logit-scale` describe the AudioCLIP side (the scene's `audio_video/audio_data_<level>.pkl` is loaded).  The
image branch needs the reference's HLoc wrapper attached as `avlmap.visual_map`.  The open3d / matplotlib views of the
reference are not reproduced -- the heat is printed as a goal voxel and optionally saved with `--out`."""
from __future__ import annotations

import sys
from pathlib import Path
from typing import List, Optional

import numpy as np

from ..map import AreaMap, AVLMap, SoundMap
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
    ap.add_argument("--area", action="append", default=[], help="area name to index (needs --area-text-encoder)")
    ap.add_argument("--sound", action="append", default=[], help="sound name to index (needs the --sound-* options)")
    ap.add_argument("--text-encoder", default=None, help="module:function, list[str] -> (len, D) array")
    ap.add_argument("--clip-dim", type=int, default=512)
    ap.add_argument("--area-text-encoder", default=None, help="module:function of the CLIP ViT-L/14 text tower")
    ap.add_argument("--area-clip-dim", type=int, default=768)
    ap.add_argument("--sound-text-encoder", default=None, help="module:function, names -> (C, 1024) AudioCLIP text features")
    ap.add_argument("--sound-categories", default="", help="comma-separated sound category names")
    ap.add_argument("--sound-logit-scale", type=float, default=float(np.log(100.0)), help="AudioCLIP logit_scale_at (log)")
    ap.add_argument("--out", default=None, help="directory for heat_<name>.npy")
    args = ap.parse_intermixed_args(argv)
    config, scene = compose_from_args(args)
    avlmap = AVLMap(config, data_dir=scene)                               # index_map.py:26
    if not avlmap.vlmap.load_map(scene):                                  # index_map.py:27 (prints and returns False)
        return 1
    enc = load_callable(args.text_encoder)
    if enc is not None:
        avlmap.vlmap.set_text_encoder(enc, args.clip_dim)
    else:
        avlmap.vlmap._init_clip()                                         # index_map.py:28
    if args.area_text_encoder:
        avlmap.area_map = AreaMap(str(scene), text_encoder=load_callable(args.area_text_encoder), clip_feat_dim=args.area_clip_dim)
        avlmap.area_map.load_map(scene)                                   # avlmap.py:51
    if args.sound_text_encoder:
        level = 1
        if "sound_data_collect_params" in config and "difficulty" in config.sound_data_collect_params:
            level = config.sound_data_collect_params.difficulty           # index_map.py:49
        cats = [c.strip() for c in args.sound_categories.split(",") if c.strip()]
        avlmap.sound_map = SoundMap(cats, load_callable(args.sound_text_encoder), args.sound_logit_scale, difficulty_level=level)
        avlmap.sound_map.load_sound_map(scene)                            # avlmap.py:53
    decay = float(config.get("decay_rate", 0.01))                         # index_map.py:38 uses 0.01
    out_dir = Path(args.out) if args.out else None

    def run(kind: str, name: str) -> None:
        if kind == "object":
            heat = avlmap.index_object(name, decay_rate=decay)           # index_map.py:38
        elif kind == "sound":
            if avlmap.sound_map is None:
                print("the sound modality needs --sound-text-encoder / --sound-categories (AudioCLIP is outside this engine)")
                return
            heat = avlmap.index_sound(name, decay_rate=decay)            # index_map.py:86
        else:
            if avlmap.area_map is None:
                print("the area modality needs --area-text-encoder (CLIP ViT-L/14 is outside this engine)")
                return
            heat = avlmap.index_area(name, decay_rate=decay)             # index_map.py:90
        _report(avlmap, name, heat, out_dir)

    if args.object or args.area or args.sound:
        for kind, names in (("object", args.object), ("area", args.area), ("sound", args.sound)):
            for name in names:
                run(kind, name)
        return 0
    while True:
        choice = input_fn(PROMPT).strip()
        if choice == "1":
            run("object", input_fn("What is the object name you want to index?\nInput: "))
        elif choice == "2":
            run("sound", input_fn("What is the sound name you want to index?\nInput: "))
        elif choice == "3":
            run("area", input_fn("What is the area name you want to index?\nInput: "))
        elif choice == "4":
            print("the image modality needs the reference's HLoc wrapper attached as avlmap.visual_map "
                  "(avlmaps_b200/map/avlmap.py); it does not run from this command line")
        elif choice == "5":
            print(f"{avlmap.vlmap.grid_pos.shape[0]} voxels; the open3d viewer of the reference is not reproduced")
        elif choice == "6":
            return 0


if __name__ == "__main__":
    sys.exit(main())
