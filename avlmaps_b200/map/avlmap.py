"""`AVLMap` facade (reference avlmaps/map/avlmap.py:18-163), object modality.

The sound / area / image modalities of the reference wrap AudioCLIP, CLIP ViT-L/14 and HLoc models
(SoundMap, AreaMap, VisualMap) that are outside this engine; their similarity call sites are served by
engine.DeviceMap.scores / topk and their fusion by engine.fuse_topk.  `index_object` is the modality
whose cost sits on the accelerated path: per-voxel argmax (tcgen05) + nearest-target distance decay."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from ..engine import heat_from_mask_3d, topk_vector
from .map import cfg_get
from .vlmap import VLMap


class AVLMap:
    def __init__(self, config, data_dir: str = "", feature_fn=None):
        self.config = config
        self.vlmap = VLMap(cfg_get(config, "map_config"), data_dir=data_dir, feature_fn=feature_fn)

    def create_map(self, data_dir) -> bool:
        self.vlmap.create_map(data_dir)   # avlmap.py:39
        return True

    def load_map(self, data_dir: str) -> bool:
        self.vlmap.load_map(data_dir)     # avlmap.py:50
        return True

    def index_object(self, object_name: str, init_categories: Optional[List[str]] = None, decay_rate: float = 0.1) -> np.ndarray:
        """Reference avlmap.py:67-76, including its quirk of dropping the first and last entry of
        `init_categories` (`init_categories[1:-1]`)."""
        if init_categories is not None:
            self.vlmap.init_categories(init_categories[1:-1], return_scores=False)
            mask = self.vlmap.index_map(object_name, with_init_cat=True)
        else:
            mask = self.vlmap.index_map(object_name, with_init_cat=False)
        cs = cfg_get(cfg_get(self.config, "params"), "cs")
        return heat_from_mask_3d(self.vlmap.grid_pos, mask, cell_size=cs, decay_rate=decay_rate)

    def get_max_pos_3d(self, heat: np.ndarray) -> np.ndarray:
        """HabitatLanguageRobot.get_max_pos_3d (habitat_lang_robot.py:427-430): grid_pos[argmax(heat)]."""
        idx, _ = topk_vector(heat, 1)
        return self.vlmap.grid_pos[int(idx[0])]
