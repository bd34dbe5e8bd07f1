"""Build the map of one scene: the reference's application/create_map.py:12-17 on the B200 engine.

    python -m avlmaps_b200.application.create_map --config-dir <reference>/config \\
        --feature-fn my_lseg:get_lseg_feat data_paths.avlmaps_data_dir=/data scene_id=0 map_config.depth_sample_rate=1

`--feature-fn module:function` is the pixel encoder, `rgb (H, W, 3) uint8 -> (1, D, FH, FW) float32`, i.e. what
`avlmaps.utils.lseg_utils.get_lseg_feat` returns (lseg_utils.py:20-119); LSeg itself is outside this engine."""
from __future__ import annotations

import sys
from typing import List, Optional

from ..map import AVLMap
from ._common import base_parser, compose_from_args, load_callable


def main(argv: Optional[List[str]] = None) -> int:
    ap = base_parser("avlmaps_b200.application.create_map", "map_creation_cfg.yaml", __doc__)
    ap.add_argument("--feature-fn", default=None, help="module:function of the pixel encoder")
    args = ap.parse_intermixed_args(argv)
    config, scene = compose_from_args(args)
    avlmap = AVLMap(config, feature_fn=load_callable(args.feature_fn))   # create_map.py:13
    avlmap.create_map(scene)                                             # create_map.py:17
    return 0


if __name__ == "__main__":
    sys.exit(main())
