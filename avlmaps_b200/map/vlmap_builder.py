"""`VLMapBuilder` with the reference's constructor and `create_mobile_base_map()` (reference
avlmaps/map/vlmap_builder.py:35-185); the per-point Python loop is replaced by the kernels of
csrc/build_path.cu through engine.DeviceBuilder.

What stays on the host, computed exactly as the reference does: the pose chain (:67-74,106-108,133),
the camera matrices (:98,126) and the sample permutation, which the reference draws from numpy's
GLOBAL RNG (:275-277) -- this class draws it the same way, so `np.random.seed(s)` before the call
gives the reference's voxel ids.

The pixel encoder (LSeg, reference lseg_utils.get_lseg_feat) is not part of this engine: pass
`feature_fn(rgb) -> (1, D, FH, FW) float32` (numpy, or a torch CUDA tensor to stay in HBM)."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Callable, List, Optional

import numpy as np

from .. import _lib as L
from ..engine import DeviceBuilder
from ..utils.mapping_utils import cvt_pose_vec2tf, get_sim_cam_mat, load_3d_map, load_depth_npy, map_file_exists, save_3d_map
from .map import cfg_get


def _default_feature_fn():
    raise RuntimeError(
        "VLMapBuilder needs a pixel encoder: pass feature_fn=... returning the (1, D, FH, FW) array that "
        "avlmaps.utils.lseg_utils.get_lseg_feat returns (LSeg itself is outside this engine)")


class VLMapBuilder:
    def __init__(self, data_dir: Path, map_config, pose_path: Path, rgb_paths: List[Path], depth_paths: List[Path],
                 base2cam_tf: np.ndarray, base_transform: np.ndarray, feature_fn: Optional[Callable] = None,
                 save_every: int = 100):
        self.data_dir = Path(data_dir)
        self.pose_path = pose_path
        self.rgb_paths = rgb_paths
        self.depth_paths = depth_paths
        self.map_config = map_config
        self.base2cam_tf = base2cam_tf
        self.base_transform = base_transform
        self.feature_fn = feature_fn
        self.save_every = save_every
        self.device_builder: Optional[DeviceBuilder] = None

    # ------------------------------------------------------------------ host geometry (reference arithmetic)
    def _frame_transforms(self, base_poses: np.ndarray) -> List[np.ndarray]:
        self.init_base_tf = self.base_transform @ cvt_pose_vec2tf(base_poses[0]) @ np.linalg.inv(self.base_transform)
        self.inv_init_base_tf = np.linalg.inv(self.init_base_tf)
        self.init_cam_tf = self.init_base_tf @ self.base2cam_tf
        self.inv_init_cam_tf = np.linalg.inv(self.init_cam_tf)
        out = []
        for base_posevec in base_poses:
            habitat_base_pose = cvt_pose_vec2tf(base_posevec)
            base_pose = self.base_transform @ habitat_base_pose @ np.linalg.inv(self.base_transform)
            tf = self.inv_init_base_tf @ base_pose
            out.append(tf @ self.base_transform @ self.base2cam_tf)  # pc_transform, :133
        return out

    @staticmethod
    def _sample_order(n_pixels: int, depth_sample_rate: int) -> np.ndarray:
        """shuffle_mask[::rate] of _backproject_depth (:275-277), drawn from numpy's global RNG."""
        shuffle_mask = np.arange(n_pixels)
        np.random.shuffle(shuffle_mask)
        return np.ascontiguousarray(shuffle_mask[::depth_sample_rate], dtype=np.int32)

    def _load_frame(self, rgb_path, depth_path):
        import cv2

        bgr = cv2.imread(str(rgb_path))
        rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
        return rgb, load_depth_npy(depth_path)

    # ------------------------------------------------------------------ the build
    def create_mobile_base_map(self):
        """Build the 3-D map centred at the first base frame (reference :54-185) and save it to
        <data_dir>/vlmap/vlmaps.h5df every `save_every` frames and at the end."""
        pose_info = cfg_get(self.map_config, "pose_info")
        camera_height = cfg_get(pose_info, "camera_height")
        cs = cfg_get(self.map_config, "cell_size")
        gs = cfg_get(self.map_config, "grid_size")
        depth_sample_rate = cfg_get(self.map_config, "depth_sample_rate")

        self.base_poses = np.loadtxt(self.pose_path)
        if self.base_poses.ndim == 1:  # a single pose line; the reference fails here (:65), we accept it
            self.base_poses = self.base_poses[None, :]
        pc_transforms = self._frame_transforms(self.base_poses)

        self.map_save_dir = self.data_dir / "vlmap"
        os.makedirs(self.map_save_dir, exist_ok=True)
        self.map_save_path = self.map_save_dir / "vlmaps.h5df"

        calib_mat = np.array(cfg_get(self.map_config, "cam_calib_mat"), dtype=np.float64).reshape((3, 3))
        calib_inv = np.linalg.inv(calib_mat)  # depth2pc, mapping_utils.py:237
        vh = int(camera_height / cs)          # _init_map, :201
        feature_fn = self.feature_fn or _default_feature_fn
        resume = None
        if map_file_exists(self.map_save_path):
            # _init_map (:212-222): the saved map becomes the initial state (max_id = grid_feat.shape[0]) and the
            # loop below fuses EVERY frame on top of it -- the reference never consults mapped_iter_set (:102-180)
            resume = load_3d_map(self.map_save_path)

        mapped_iter_set = set(resume[0]) if resume is not None else set()
        builder = None
        for frame_i, (rgb_path, depth_path) in enumerate(zip(self.rgb_paths, self.depth_paths)):
            rgb, depth = self._load_frame(rgb_path, depth_path)
            pix_feats = feature_fn(rgb)  # (1, D, FH, FW), like get_lseg_feat (:123-125)
            if builder is None:
                self.clip_feat_dim = int(pix_feats.shape[1])
                builder = DeviceBuilder(gs, vh, cs, self.clip_feat_dim, capacity=gs * gs)  # :202
                self.device_builder = builder
                if resume is not None:
                    builder.import_state(resume[1], resume[2], resume[3], resume[5])
            pix_feats_intr = get_sim_cam_mat(pix_feats.shape[2], pix_feats.shape[3])  # :126
            sample_idx = self._sample_order(depth.shape[0] * depth.shape[1], depth_sample_rate)
            on_device = type(pix_feats).__module__.startswith("torch") and pix_feats.is_cuda
            if on_device:
                import torch

                depth_a = torch.from_numpy(np.ascontiguousarray(depth, np.float32)).cuda()
                rgb_a = torch.from_numpy(np.ascontiguousarray(rgb)).cuda()
                sidx_a = torch.from_numpy(sample_idx).cuda()
            else:
                depth_a, rgb_a, sidx_a = depth, rgb, sample_idx
            builder.add_frame(depth_a, pix_feats, calib_inv, calib_mat, pix_feats_intr, pc_transforms[frame_i],
                              rgb=rgb_a, sample_idx=sidx_a, feat_layout=L.FEAT_CHW, min_depth=0.1, max_depth=6)
            mapped_iter_set.add(frame_i)
            if self.save_every and frame_i % self.save_every == self.save_every - 1:
                print(f"Temporarily saving {builder.num_voxels} features at iter {frame_i}...")
                self._save_3d_map(builder, mapped_iter_set)
        if builder is not None:
            self._save_3d_map(builder, mapped_iter_set)

    def create_camera_map(self):
        """Not implemented in the reference either, which RETURNS the exception class (:187-193)."""
        return NotImplementedError

    def _save_3d_map(self, builder: DeviceBuilder, mapped_iter_set) -> None:
        out = builder.export()
        save_3d_map(self.map_save_path, out["grid_feat"], out["grid_pos"], out["weight"], out["occupied_ids"],
                    list(mapped_iter_set), out["grid_rgb"])
