"""`VLMapsDataloaderHabitat` with the reference's method surface (reference
avlmaps/dataloader/habitat_dataloader.py:21-148): converts a base pose in the habitat frame to the map's
(row, col, angle_deg) and back.  `AVLMap.index_area / index_sound / index_image` route every frame, sound-segment and
localisation pose through it (avlmap.py:84-90, 117-123, 152-154), so it sits on the caller side of the heat kernels.
Scalar float64 pose algebra, kept on the host exactly as the reference evaluates it.

The full-map pose is (row, col, angle_deg), angle 0 pointing towards negative rows; the cropped-map pose subtracts
the obstacle map's (rmin, cmin)."""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional, Union

import numpy as np

from ..map.map import Map, cfg_get
from ..utils.mapping_utils import base_pos2grid_id_3d, cvt_pose_vec2tf, grid_id2base_pos_3d


def base_rot_mat2theta(rot_mat: np.ndarray) -> float:
    """Yaw of a base rotation with x forward, y left, z up (reference mapping_utils.py:379-389)."""
    return np.arctan2(rot_mat[1, 0], rot_mat[0, 0])


class VLMapsDataloaderHabitat:
    def __init__(self, data_dir: Union[Path, str], map_config, map: Optional[Map] = None, load_gt_map: bool = False):
        """Reference habitat_dataloader.py:29-85: with `map=None` the map named by `map_config.map_type` is created and
        loaded from `data_dir`; either way its obstacle map is (re)generated for the crop offsets."""
        self.data_dir = data_dir
        self.map_config = map_config
        self.map = map
        self.cs = cfg_get(map_config, "cell_size")
        self.gs = cfg_get(map_config, "grid_size")
        self.camera_height = cfg_get(cfg_get(map_config, "pose_info"), "camera_height")
        if map is None:
            self.map_dir = str(Path(data_dir) / cfg_get(map_config, "map_type"))
            self.map = Map.create(map_config)
            load_success = self.map.load_map(data_dir)
            assert load_success is True, f"Map loading fails. It could be because the map hasn't been created at {self.map_dir}."
        self.map.generate_obstacle_map()
        self.obstacles = self.map.obstacles_map
        self.obstacles_cropped = self.map.obstacles_cropped
        self.rmin, self.rmax, self.cmin, self.cmax = self.map.rmin, self.map.rmax, self.map.cmin, self.map.cmax
        self.base2cam_tf, self.base_transform = self.map.base2cam_tf, self.map.base_transform
        self.base_poses = np.loadtxt(self.map.pose_path)
        first = self.base_poses[0] if self.base_poses.ndim == 2 else self.base_poses   # a one-line poses.txt loads as 1-D
        self.init_base_tf = self.base_transform @ cvt_pose_vec2tf(first) @ np.linalg.inv(self.base_transform)
        self.inv_init_base_tf = np.linalg.inv(self.init_base_tf)
        self.full_map_pose = None  # (row, col, theta_deg)

    def get_obstacles_cropped(self) -> np.ndarray:
        return self.obstacles_cropped

    def get_color_topdown_bgr_cropped(self, times: int = 1) -> np.ndarray:
        """Reference habitat_dataloader.py:97-105 (including its use of shape[1] for both resize extents)."""
        import cv2

        top = self.map.generate_rgb_topdown_map()[self.rmin:self.rmax + 1, self.cmin:self.cmax + 1]
        bgr = cv2.cvtColor(top, cv2.COLOR_RGB2BGR)
        return cv2.resize(bgr, (bgr.shape[1] * times, bgr.shape[1] * times))

    # ---- setters
    def from_cropped_map_pose(self, row: int, col: int, theta_deg: float):
        self.full_map_pose = [row + self.rmin, col + self.cmin, theta_deg]

    def from_full_map_pose(self, row: int, col: int, theta_deg: float):
        self.full_map_pose = [row, col, theta_deg]

    def from_habitat_tf(self, tf_hab: np.ndarray):
        """Reference habitat_dataloader.py:115-121."""
        tf = self.inv_init_base_tf @ self.base_transform @ tf_hab @ np.linalg.inv(self.base_transform)
        theta_deg = np.rad2deg(base_rot_mat2theta(tf[:3, :3]))
        x, y, z = tf[:3, 3]
        row, col, _ = base_pos2grid_id_3d(self.gs, self.cs, x, y, z)
        self.full_map_pose = [row, col, theta_deg]

    def from_camera_tf(self, tf_cam: np.ndarray):
        """Reference habitat_dataloader.py:123-125."""
        self.from_habitat_tf(self.base_transform @ self.inv_init_base_tf @ self.base2cam_tf @ tf_cam)

    # ---- getters
    def to_cropped_map_pose(self) -> List:
        assert self.full_map_pose is not None, "Please call from_xx() first."
        return [self.full_map_pose[0] - self.rmin, self.full_map_pose[1] - self.cmin, self.full_map_pose[2]]

    def to_full_map_pose(self) -> List:
        assert self.full_map_pose is not None, "Please call from_xx() first."
        return self.full_map_pose

    def to_habitat_tf(self) -> np.ndarray:
        """Reference habitat_dataloader.py:135-148 (the cell centre at height 0, yaw about z in the base frame)."""
        assert self.full_map_pose is not None, "Please call from_xx() first."
        row, col, theta_deg = self.full_map_pose
        x, y, z = grid_id2base_pos_3d(row, col, 0, self.cs, self.gs)
        theta = np.deg2rad(theta_deg)
        tf = np.eye(4)
        tf[:3, 3] = [x, y, z]
        tf[0, 0] = tf[1, 1] = np.cos(theta)
        tf[0, 1] = -np.sin(theta)
        tf[1, 0] = np.sin(theta)
        return np.linalg.inv(self.base_transform) @ self.init_base_tf @ tf @ self.base_transform
