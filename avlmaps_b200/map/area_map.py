"""`AreaMap` (reference avlmaps/map/area_map.py:19-119): one CLIP image embedding per frame and a true-cosine
similarity against text embeddings.  The encoders (CLIP ViT-L/14 image / text towers) are the caller's --
`image_encoder(rgb) -> (D,)`, `text_encoder(texts) -> (len, D)` (or an openai-CLIP model for the text side,
like the reference); the similarity `clip_sparse_map @ text_feats.T` (:102, :118) runs through the same
exact kernel as VLMap's scores (engine.DeviceMap.scores), so AVLMap.index_area_2d gets the reference's numbers.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Callable, List, Optional, Union

import numpy as np

from ..engine import DeviceMap
from ..utils.clip_utils import get_text_feats
from ..utils.mapping_utils import cvt_pose_vec2tf, load_clip_sparse_map, save_clip_sparse_map
from .vlmap import find_similar_category_id


class AreaMap:
    def __init__(self, data_dir: str = "", image_encoder: Optional[Callable] = None, text_encoder=None,
                 clip_feat_dim: int = 768) -> None:
        self.clip_sparse_map = None
        self.robot_pose_list = None
        self.scores_mat = None
        self.categories = None
        self.image_encoder = image_encoder
        self.clip_model = text_encoder          # the reference keeps both towers in self.clip_model (:50)
        self.clip_feat_dim = clip_feat_dim      # ViT-L/14 (:36-46)
        self._device_map: Optional[DeviceMap] = None
        if data_dir:
            self._setup_paths(data_dir)

    def _setup_paths(self, data_dir: Union[Path, str]) -> None:
        """Reference area_map.py:53-63."""
        self.data_dir = Path(data_dir)
        self.rgb_dir = self.data_dir / "rgb"
        self.depth_dir = self.data_dir / "depth"
        self.pose_path = self.data_dir / "poses.txt"
        self.map_save_dir = self.data_dir / "area_map"
        os.makedirs(self.map_save_dir, exist_ok=True)
        try:
            self.rgb_paths = sorted(self.rgb_dir.glob("*.png"))
            self.depth_paths = sorted(self.depth_dir.glob("*.npy"))
        except FileNotFoundError as e:
            print(e)

    def set_sparse_map(self, clip_sparse_map: np.ndarray, robot_pose_list) -> None:
        self.clip_sparse_map = np.ascontiguousarray(clip_sparse_map, np.float32)
        self.robot_pose_list = robot_pose_list
        if self._device_map is not None:
            self._device_map.close()
        self._device_map = DeviceMap(self.clip_sparse_map)
        self.scores_mat, self.categories = None, None

    def create_map(self, data_dir: Union[Path, str]) -> None:
        """Reference area_map.py:65-92: embed every frame, L2-normalised (clip_utils.get_img_feats :99-103)."""
        import cv2

        if self.image_encoder is None:
            raise RuntimeError("AreaMap.create_map needs image_encoder=... (CLIP ViT-L/14 is outside this engine)")
        self._setup_paths(Path(data_dir))
        self.base_poses = np.loadtxt(self.pose_path)
        clip_sparse_map = np.zeros((len(self.rgb_paths), self.clip_feat_dim), dtype=np.float32)
        robot_pose_list = []
        for iter_i, (rgb_path, base_posevec) in enumerate(zip(self.rgb_paths, self.base_poses)):
            rgb = cv2.cvtColor(cv2.imread(rgb_path.as_posix()), cv2.COLOR_BGR2RGB)
            feats = np.asarray(self.image_encoder(rgb), np.float32).reshape(-1)
            clip_sparse_map[iter_i] = feats / np.linalg.norm(feats)
            robot_pose_list.append(cvt_pose_vec2tf(base_posevec))
        save_clip_sparse_map(self.map_save_dir / "clip_sparse_map.h5df", clip_sparse_map, robot_pose_list)
        self.set_sparse_map(clip_sparse_map, robot_pose_list)

    def load_map(self, data_dir: Union[Path, str]) -> None:
        self._setup_paths(data_dir)
        self.set_sparse_map(*load_clip_sparse_map(self.map_save_dir / "clip_sparse_map.h5df"))

    def init_categories(self, categories: List[str]) -> np.ndarray:
        """Reference area_map.py:99-103."""
        self.categories = categories
        text_feats = get_text_feats(categories, self.clip_model, self.clip_feat_dim)
        self.scores_mat = self._device_map.scores(text_feats)   # clip_sparse_map @ text_feats.T
        return self.scores_mat

    def index_map(self, language_desc: str, with_init_cat: bool = True):
        """Reference area_map.py:105-119: (F,) scores of every frame for one description."""
        if with_init_cat and self.scores_mat is not None and self.categories is not None:
            cat_id = find_similar_category_id(language_desc, self.categories)
            return self.scores_mat[:, cat_id].flatten()
        if with_init_cat:
            raise Exception(
                "Categories are not preloaded. Call init_categories(categories: List[str]) to initialize categories.")
        text_feats = get_text_feats([language_desc], self.clip_model, self.clip_feat_dim)
        return self._device_map.scores(text_feats).flatten()
