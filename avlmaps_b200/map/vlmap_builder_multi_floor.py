"""`VLMapBuilderMultiFloor` with the reference's constructor and `create_global_map()` (reference
avlmaps/map/vlmap_builder_multi_floor.py:41-199).  Both passes over the frames run on the B200:

* pass 1 (:97-118) -- the reference accumulates every back-projected point of every frame in an
  open3d cloud only to take its min / max; here engine.FrameBounds keeps a running min / max on the
  device (exact, order-free) -> `pcd_min`, `pcd_max`, grid size `ceil((max - min) / cs + 1)` (:222);
* pass 2 (:124-199) -- the fusion loop, through engine.DeviceBuilder.global_grid: cells are
  `np.round((p - pcd_min) / cs)` as (row, height, col), negative indices wrap like numpy's.

Kept on the host exactly as the reference computes it: `pose @ diag(1, -1, -1, 1)` (:77-79,105,141),
the calibration matrices, and the sample permutations, drawn from numpy's GLOBAL RNG once per frame
in EACH pass (:368-370) -- so `np.random.seed(s)` before the call reproduces the reference's voxel ids.

Where the reference dies with an IndexError (a second-pass sample above the first-pass bounds in
height, or below -size) the point is skipped and counted (`device_builder.num_rejected_oob`).
`create_mobile_base_map` / `create_camera_map` RETURN NotImplementedError like the reference (:201-215)."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Callable, List, Optional

import numpy as np

from .. import _lib as L
from ..engine import DeviceBuilder, FrameBounds
from ..utils.mapping_utils import get_sim_cam_mat, load_depth_img, map_file_exists, save_3d_map_multi_floor
from .map import cfg_get
from .vlmap_builder import VLMapBuilder, _default_feature_fn


class VLMapBuilderMultiFloor:
    def __init__(self, data_dir: Path, map_config, pose_paths: List[Path], rgb_paths: List[Path],
                 depth_paths: List[Path], base2cam_tf: np.ndarray, base_transform: np.ndarray,
                 feature_fn: Optional[Callable] = None):
        self.data_dir = Path(data_dir)
        self.pose_paths = pose_paths
        self.rgb_paths = rgb_paths
        self.depth_paths = depth_paths
        self.map_config = map_config
        self.base2cam_tf = base2cam_tf
        self.base_transform = base_transform
        self.feature_fn = feature_fn
        self.device_builder: Optional[DeviceBuilder] = None

    @staticmethod
    def _load_depth_mm(depth_path) -> np.ndarray:
        """uint16 millimetres; the `/ 1000.0` of the reference (:103,128) happens in the kernel, in float64."""
        depth = load_depth_img(str(depth_path))
        if depth is None:
            raise FileNotFoundError(depth_path)
        if depth.dtype != np.uint16:
            # the reference divides whatever cv2 returns by 1000.0; only 16-bit PNGs keep that exact on the device
            raise ValueError(f"{depth_path}: expected a 16-bit depth PNG (millimetres), got {depth.dtype}")
        return depth

    def create_global_map(self):
        """Build the map centred at the global origin (reference :60-199) and save it to
        <data_dir>/vlmap_multi_floor/vlmaps_multi_floor.h5df."""
        cs = cfg_get(self.map_config, "cell_size")
        depth_sample_rate = cfg_get(self.map_config, "depth_sample_rate")
        skip_frame = cfg_get(self.map_config, "skip_frame")
        self.camera_pose_tfs = [np.loadtxt(p).reshape((4, 4)) for p in self.pose_paths]
        self.init_cam_tf = self.camera_pose_tfs[0]
        self.inv_init_cam_tf = np.linalg.inv(self.init_cam_tf)
        self.habitat2cam_rot_tf = np.eye(4)
        self.habitat2cam_rot_tf[1, 1] = -1
        self.habitat2cam_rot_tf[2, 2] = -1

        self.map_save_dir = self.data_dir / "vlmap_multi_floor"
        os.makedirs(self.map_save_dir, exist_ok=True)
        self.map_save_path = self.map_save_dir / "vlmaps_multi_floor.h5df"
        if map_file_exists(self.map_save_path):
            # the reference's reload path unpacks 8 names from the 9-tuple load_3d_map returns (:236 vs :257)
            # and cannot run; we refuse explicitly instead of re-fusing on top of a saved map
            raise NotImplementedError(f"{self.map_save_path} exists: resume of a saved map is not supported")

        calib_mat = np.array(cfg_get(self.map_config, "cam_calib_mat"), dtype=np.float64).reshape((3, 3))
        calib_inv = np.linalg.inv(calib_mat)  # depth2pc, mapping_utils.py:237
        feature_fn = self.feature_fn or _default_feature_fn
        frames = [i for i in range(len(self.rgb_paths)) if i % skip_frame == 0]

        # ---- pass 1: bounds of the global cloud (:97-118)
        bounds = FrameBounds()
        for frame_i in frames:
            depth_mm = self._load_depth_mm(self.depth_paths[frame_i])
            sample_idx = VLMapBuilder._sample_order(depth_mm.shape[0] * depth_mm.shape[1], depth_sample_rate)
            transform_tf = self.camera_pose_tfs[frame_i] @ self.habitat2cam_rot_tf
            bounds.add_frame(depth_mm, calib_inv, transform_tf, sample_idx=sample_idx, min_depth=0.1, max_depth=100)
        self.pcd_min, self.pcd_max, n_points = bounds.get()
        bounds.close()
        if n_points == 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")  # np.min, :117
        grid_size = np.ceil((self.pcd_max - self.pcd_min) / cs + 1).astype(int)  # (x, y, z) = col?, height, row?  (:222)
        self.grid_size = grid_size

        # ---- pass 2: fusion (:124-199)
        import cv2

        mapped_iter_set = set()
        builder = None
        for frame_i in frames:
            bgr = cv2.imread(str(self.rgb_paths[frame_i]))
            rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
            depth_mm = self._load_depth_mm(self.depth_paths[frame_i])
            pix_feats = feature_fn(rgb)  # (1, D, FH, FW), like get_lseg_feat (:131-133)
            if builder is None:
                self.clip_feat_dim = int(pix_feats.shape[1])
                builder = DeviceBuilder.global_grid(int(grid_size[0]), int(grid_size[2]), int(grid_size[1]), cs,
                                                    self.pcd_min, self.clip_feat_dim)  # occupied_ids: grid_size[[0, 2, 1]]
                self.device_builder = builder
            pix_feats_intr = get_sim_cam_mat(pix_feats.shape[2], pix_feats.shape[3])  # :135
            sample_idx = VLMapBuilder._sample_order(depth_mm.shape[0] * depth_mm.shape[1], depth_sample_rate)
            transform_tf = self.camera_pose_tfs[frame_i] @ self.habitat2cam_rot_tf  # :141
            on_device = type(pix_feats).__module__.startswith("torch") and pix_feats.is_cuda
            if on_device:
                import torch

                depth_a = torch.from_numpy(depth_mm).cuda()
                rgb_a = torch.from_numpy(np.ascontiguousarray(rgb)).cuda()
                sidx_a = torch.from_numpy(sample_idx).cuda()
            else:
                depth_a, rgb_a, sidx_a = depth_mm, rgb, sample_idx
            builder.add_frame(depth_a, pix_feats, calib_inv, calib_mat, pix_feats_intr, transform_tf, rgb=rgb_a,
                              sample_idx=sidx_a, feat_layout=L.FEAT_CHW, min_depth=0.1, max_depth=100)
            mapped_iter_set.add(frame_i)
            if frame_i % (skip_frame * 100) == skip_frame * 99:  # :195-197
                print(f"Temporarily saving {builder.num_voxels} features at iter {frame_i}...")
                self._save(builder, mapped_iter_set)
        if builder is not None:
            self._save(builder, mapped_iter_set)

    def create_mobile_base_map(self):
        return NotImplementedError

    def create_camera_map(self):
        return NotImplementedError

    def _save(self, builder: DeviceBuilder, mapped_iter_set) -> None:
        out = builder.export()
        save_3d_map_multi_floor(self.map_save_path, out["grid_feat"], out["grid_pos"], out["weight"], out["grid_rgb"],
                                out["occupied_ids"], mapped_iter_set, self.pcd_min, self.pcd_max,
                                cfg_get(self.map_config, "cell_size"))
