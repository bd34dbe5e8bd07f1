"""`AVLMap` facade (reference avlmaps/map/avlmap.py:18-163).

`index_object` is the modality whose cost sits on the accelerated path: per-voxel argmax (tcgen05) + nearest-target
distance decay.  The sound / area / image modalities wrap AudioCLIP, CLIP ViT-L/14 and HLoc models in the reference;
here `avlmaps_b200.map.SoundMap` / `AreaMap` take those encoders as callables (or attach the reference's own
objects -- only the methods listed below are used), and the heat maps run on the device.  Cross-modal goal selection
over dense per-voxel modalities is `engine.fuse_topk`."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from ..engine import heat2d_normalize_lift, heat2d_sources, heat_from_mask_3d, heat_planar, topk_vector
from .map import cfg_get
from .vlmap import VLMap


class AVLMap:
    def __init__(self, config, data_dir: str = "", feature_fn=None):
        self.config = config
        self.vlmap = VLMap(cfg_get(config, "map_config"), data_dir=data_dir, feature_fn=feature_fn)

    def create_map(self, data_dir) -> bool:
        """Reference avlmap.py:38-47; the area / visual / sound maps are built too when such objects are attached."""
        self.vlmap.create_map(data_dir)   # avlmap.py:39
        if self.area_map is not None:
            self.area_map.create_map(data_dir)
        if self.visual_map is not None:
            self.visual_map.create_and_load_map(data_dir)
        if self.sound_map is not None and hasattr(self.sound_map, "create_sound_map"):
            self.sound_map.create_sound_map(data_dir)
        return True

    def load_map(self, data_dir: str) -> bool:
        """Reference avlmap.py:49-55."""
        self.vlmap.load_map(data_dir)     # avlmap.py:50
        if self.area_map is not None:
            self.area_map.load_map(data_dir)
        if self.visual_map is not None:
            self.visual_map.create_and_load_map(data_dir)
        if self.sound_map is not None:
            self.sound_map.load_sound_map(data_dir)
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

    # ------------------------------------------------------------------ area / sound modalities
    # AreaMap (CLIP ViT-L/14 per-frame embeddings) and SoundMap (AudioCLIP segment database) are third-party
    # model wrappers outside this engine.  Attach objects with the reference's interface
    #   area_map.index_map(name, with_init_cat=False) -> (F,) scores,  area_map.robot_pose_list
    #   sound_map.get_distribution_and_locations(name) -> (probabilities, locations_list)
    #   dataloader.from_habitat_tf(tf); dataloader.to_full_map_pose() -> (row, col, deg)
    # (e.g. the reference's own AreaMap / SoundMap / VLMapsDataloaderHabitat) and the heat maps below run on
    # the device: the F (or M) full-grid distance_transform_edt calls become one kernel.
    area_map = None
    sound_map = None
    dataloader = None

    @staticmethod
    def area_heat_2d(shape, cells: List, scores: np.ndarray, decay_rate: float = 0.1) -> np.ndarray:
        """Numeric core of index_area_2d (avlmap.py:78-98): `cells[i]` is the (row, col) of frame i or None when
        it falls outside the grid; `scores` are the min-max normalised frame scores."""
        groups = [np.zeros((0, 2), np.int32) if c is None else np.asarray(c, np.int32).reshape(1, 2) for c in cells]
        dist_map = heat2d_sources(shape, groups, np.asarray(scores, np.float32), decay_rate, "area")
        return heat2d_normalize_lift(dist_map)[0]   # (dist_map - min) / (max - min), avlmap.py:97

    @staticmethod
    def sound_heat_2d(shape, cells_per_segment: List, probabilities: np.ndarray, decay_rate: float = 0.01) -> np.ndarray:
        """Numeric core of index_sound_2d (avlmap.py:111-133)."""
        rows, cols = shape
        groups = []
        for seg in cells_per_segment:
            seg = np.asarray(seg, np.int64).reshape(-1, 2).copy()
            seg[seg[:, 0] < 0, 0] += rows    # the reference indexes tmp_dist_map[row, col] unchecked: negatives wrap
            seg[seg[:, 1] < 0, 1] += cols
            if seg.size and (seg.min() < 0 or seg[:, 0].max() >= rows or seg[:, 1].max() >= cols):
                raise IndexError("sound location outside the map (the reference raises here too)")
            groups.append(seg.astype(np.int32))
        dist_map = heat2d_sources(shape, groups, np.asarray(probabilities, np.float32), decay_rate, "sound")
        return heat2d_normalize_lift(dist_map)[0]   # avlmap.py:131

    def lift_heat_2d_to_3d(self, heatmap_2d: np.ndarray) -> np.ndarray:
        """avlmap.py:100-109 / 135-144: heatmap_3d[id] = heatmap_2d[row, col] for every occupied cell -- the
        Python loop over np.where(occupied_ids != -1) is a gather through grid_pos."""
        return heat2d_normalize_lift(np.asarray(heatmap_2d), self.vlmap.grid_pos, normalize=False)[1]

    def _ensure_dataloader(self):
        """The reference builds its pose converter in load_map (avlmap.py:54); here it is made on first use, so that
        maps without a poses.txt (or with an attached converter) keep working."""
        if self.dataloader is None:
            from ..dataloader import VLMapsDataloaderHabitat

            self.dataloader = VLMapsDataloaderHabitat(self.vlmap.data_dir, cfg_get(self.config, "map_config"), self.vlmap)
        return self.dataloader

    def index_area_2d(self, area_name: str, decay_rate: float = 0.1) -> np.ndarray:
        self._ensure_dataloader()
        scores = self.area_map.index_map(area_name, with_init_cat=False)
        scores = (scores - np.min(scores)) / (np.max(scores) - np.min(scores))          # avlmap.py:81
        shape = self.vlmap.occupied_ids.shape[:2]
        cells = []
        for tf_hab in self.area_map.robot_pose_list:
            self.dataloader.from_habitat_tf(tf_hab)
            row, col, _ = self.dataloader.to_full_map_pose()
            cells.append(None if (row < 0 or row >= shape[0] or col < 0 or col >= shape[1]) else (row, col))
        return self.area_heat_2d(shape, cells, scores, decay_rate)

    def index_area(self, area_name: str, decay_rate: float = 0.1) -> np.ndarray:
        return self.lift_heat_2d_to_3d(self.index_area_2d(area_name, decay_rate))

    def index_sound_2d(self, sound_name: str, decay_rate: float = 0.01) -> np.ndarray:
        self._ensure_dataloader()
        probabilities, locations_list = self.sound_map.get_distribution_and_locations(sound_name)
        shape = self.vlmap.occupied_ids.shape[:2]
        segs = []
        for locations in locations_list:
            cells = []
            for location in locations:
                tf_hab = np.eye(4)
                tf_hab[:3, 3] = location
                self.dataloader.from_habitat_tf(tf_hab)
                row, col, _ = self.dataloader.to_full_map_pose()
                cells.append((row, col))
            segs.append(cells)
        return self.sound_heat_2d(shape, segs, probabilities, decay_rate)

    def index_sound(self, sound_name: str, decay_rate: float = 0.01) -> np.ndarray:
        return self.lift_heat_2d_to_3d(self.index_sound_2d(sound_name, decay_rate))

    visual_map = None

    def index_image(self, image: np.ndarray, query_cam_intrinsics: np.ndarray = None, decay_rate: float = 0.01) -> np.ndarray:
        """Reference avlmap.py:146-163.  `visual_map.localize_image` (HLoc: NetVLAD + SuperPoint/SuperGlue + PnP) is
        outside this engine -- attach the reference's VisualMap; the per-voxel heat runs on the device."""
        self._ensure_dataloader()
        _, query_base_tf = self.visual_map.localize_image(image, query_cam_intrinsic_mat=query_cam_intrinsics)
        self.dataloader.from_habitat_tf(query_base_tf)
        row, col, _ = self.dataloader.to_full_map_pose()
        return self.image_heat(row, col, decay_rate)

    def image_heat(self, row: float, col: float, decay_rate: float = 0.01) -> np.ndarray:
        """Numeric core of index_image (avlmap.py:156-162): clip(1 - decay * planar distance to (row, col), 0, 1)."""
        return heat_planar(self.vlmap.grid_pos, row, col, 1.0, decay_rate)

    def get_max_pos_3d(self, heat: np.ndarray) -> np.ndarray:
        """HabitatLanguageRobot.get_max_pos_3d (habitat_lang_robot.py:427-430): grid_pos[argmax(heat)]."""
        idx, _ = topk_vector(heat, 1)
        return self.vlmap.grid_pos[int(idx[0])]
