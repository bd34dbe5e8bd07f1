"""`Map` base class with the reference's state, attribute names and conventions
(reference avlmaps/map/map.py:18-129).  The 2-D spatial-relation helpers of the reference
(map.py:183-485, scalar contour geometry) are outside the accelerated path and not reproduced."""
from __future__ import annotations

from pathlib import Path
from typing import List, Union

import numpy as np


def cfg_get(cfg, key):
    """map_config is read both as attribute and as key in the reference (map.py:23-24,60)."""
    if isinstance(cfg, dict):
        return cfg[key]
    try:
        return cfg[key]
    except Exception:  # noqa: BLE001
        return getattr(cfg, key)


class Map:
    def __init__(self, map_config, data_dir: str = ""):
        self.map_config = map_config
        self.gs = cfg_get(map_config, "grid_size")
        self.cs = cfg_get(map_config, "cell_size")

        self.mapped_iter_list = None
        self.grid_feat = None
        self.grid_pos = None
        self.weight = None
        self.occupied_ids = None
        self.grid_rgb = None

        self.obstacles_map = None
        self.obstacles_cropped = None

        self._setup_transforms()
        if data_dir:
            self._setup_paths(data_dir)

    def _setup_paths(self, data_dir: Union[Path, str]) -> None:
        """Reference map.py:41-52."""
        self.data_dir = Path(data_dir)
        self.rgb_dir = self.data_dir / "rgb"
        self.depth_dir = self.data_dir / "depth"
        self.semantic_dir = self.data_dir / "semantic"
        self.pose_path = self.data_dir / "poses.txt"
        try:
            self.rgb_paths = sorted(self.rgb_dir.glob("*.png"))
            self.depth_paths = sorted(self.depth_dir.glob("*.npy"))
            self.semantic_paths = sorted(self.semantic_dir.glob("*.npy"))
        except FileNotFoundError as e:
            print(e)

    def _setup_transforms(self):
        """base2cam_tf and base_transform (reference map.py:54-68)."""
        pose_info = cfg_get(self.map_config, "pose_info")
        self.base2cam_tf = np.eye(4)
        self.base2cam_tf[:3, :3] = np.array([cfg_get(pose_info, "base2cam_rot")]).reshape((3, 3))
        self.base2cam_tf[1, 3] = cfg_get(pose_info, "camera_height")
        self.base_transform = np.eye(4)
        self.base_transform[0, :3] = cfg_get(pose_info, "base_forward_axis")
        self.base_transform[1, :3] = cfg_get(pose_info, "base_left_axis")
        self.base_transform[2, :3] = cfg_get(pose_info, "base_up_axis")
        return self.base2cam_tf, self.base_transform

    # the reference's base-class stubs RETURN (not raise) NotImplementedError (map.py:70-77)
    def create_map(self, data_dir: Union[Path, str]):
        return NotImplementedError

    def load_map(self, map_dir: str):
        return NotImplementedError

    def index_map(self, language_desc: str, with_init_cat: bool = True):
        return NotImplementedError

    def init_categories(self, categories: List[str]) -> np.ndarray:
        return NotImplementedError

    def generate_obstacle_map(self, h_min: float = 0, h_max: float = 1.5) -> np.ndarray:
        """Reference map.py:79-95 (including its `occupied_ids > 0` quirk: voxel id 0 counts as free)."""
        assert self.occupied_ids is not None, "map not loaded"
        heights = np.arange(0, self.occupied_ids.shape[-1]) * self.cs
        height_mask = np.logical_and(heights > h_min, heights < h_max)
        self.obstacles_map = np.sum(self.occupied_ids[..., height_mask] > 0, axis=2) == 0
        self.generate_cropped_obstacle_map(self.obstacles_map)
        return self.obstacles_map

    def generate_cropped_obstacle_map(self, obstacle_map: np.ndarray) -> np.ndarray:
        """Reference map.py:97-104."""
        x_indices, y_indices = np.where(obstacle_map == 0)
        self.rmin, self.rmax = np.min(x_indices), np.max(x_indices)
        self.cmin, self.cmax = np.min(y_indices), np.max(y_indices)
        self.obstacles_cropped = obstacle_map[self.rmin:self.rmax + 1, self.cmin:self.cmax + 1]
        return self.obstacles_cropped

    def generate_rgb_topdown_map(self) -> np.ndarray:
        """Reference map.py:106-113 (a per-voxel loop: later voxels overwrite earlier ones of the same column)."""
        assert self.grid_rgb is not None, "map not loaded"
        assert self.grid_pos is not None
        rgb_topdown = np.zeros((self.gs, self.gs, 3))
        # numpy's fancy assignment keeps the LAST value written for repeated indices, like the loop
        rgb_topdown[self.grid_pos[:, 0], self.grid_pos[:, 1], :] = np.asarray(self.grid_rgb).reshape(-1, 3)
        return rgb_topdown.astype(np.uint8)

    @staticmethod
    def _dilate_map(binary_map: np.ndarray, dilate_iter: int = 0, gaussian_sigma: float = 1.0):
        """Reference map.py:170-181 (host-side 2-D morphology on the small cropped obstacle map)."""
        import cv2
        from scipy.ndimage import binary_dilation, gaussian_filter

        h, w = binary_map.shape
        binary_map = cv2.resize(binary_map.astype(float), (w * 2, h * 2))
        binary_map = gaussian_filter((binary_map).astype(float), sigma=gaussian_sigma, truncate=3)
        binary_map = (binary_map > 0.5).astype(np.uint8)
        binary_map = binary_dilation(binary_map, structure=np.ones((3, 3)), iterations=dilate_iter * 2)
        return cv2.resize(binary_map.astype(float), (w, h))

    @staticmethod
    def create(map_config) -> "Map":
        """Reference map.py:121-129."""
        from . import VLMap, VLMapMultiFloor

        if cfg_get(map_config, "map_type") == "vlmap":
            return VLMap(map_config)
        if cfg_get(map_config, "map_type") == "vlmap_openmap":  # reference map.py:128-129
            return VLMapMultiFloor(map_config)
        raise NotImplementedError("map_type must be 'vlmap' or 'vlmap_openmap'")
