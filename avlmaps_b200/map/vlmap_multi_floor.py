"""`VLMapMultiFloor` (reference avlmaps/map/vlmap_multi_floor.py:27-144): the same index surface as VLMap
(init_categories / index_map run on the resident DeviceMap), global-frame build through
VLMapBuilderMultiFloor, and the extra `pcd_min, pcd_max, cs` attributes after load_map (:66-84)."""
from __future__ import annotations

from pathlib import Path
from typing import Union

from ..utils.mapping_utils import load_3d_map_multi_floor, map_file_exists
from .map import cfg_get
from .vlmap import VLMap
from .vlmap_builder_multi_floor import VLMapBuilderMultiFloor


class VLMapMultiFloor(VLMap):
    def _setup_paths(self, data_dir: Union[Path, str]) -> None:
        """Reference vlmap_multi_floor.py:33-46: PNG depth, one pose file per frame."""
        self.data_dir = Path(data_dir)
        self.rgb_dir = self.data_dir / "rgb"
        self.depth_dir = self.data_dir / "depth"
        self.semantic_dir = self.data_dir / "semantic"
        self.pose_dir = self.data_dir / "pose"
        try:
            self.rgb_paths = sorted(self.rgb_dir.glob("*.png"))
            self.depth_paths = sorted(self.depth_dir.glob("*.png"))
            self.semantic_paths = sorted(self.semantic_dir.glob("*.npy"))
            self.pose_paths = sorted(self.pose_dir.glob("*.txt"))
        except FileNotFoundError as e:
            print(e)

    def create_map(self, data_dir: Union[Path, str]) -> None:
        """Reference vlmap_multi_floor.py:48-65."""
        print("Creating map for scene at: ", data_dir)
        self._setup_paths(data_dir)
        self.map_builder = VLMapBuilderMultiFloor(self.data_dir, self.map_config, self.pose_paths, self.rgb_paths,
                                                  self.depth_paths, self.base2cam_tf, self.base_transform,
                                                  feature_fn=self.feature_fn)
        pose_type = cfg_get(cfg_get(self.map_config, "pose_info"), "pose_type")
        if pose_type == "mobile_base":
            self.map_builder.create_mobile_base_map()
        elif pose_type == "camera":
            self.map_builder.create_camera_map()
        elif pose_type == "global":
            self.map_builder.create_global_map()

    def load_map(self, data_dir: str) -> bool:
        """Reference vlmap_multi_floor.py:66-84."""
        self._setup_paths(data_dir)
        self.map_save_path = Path(data_dir) / "vlmap_multi_floor" / "vlmaps_multi_floor.h5df"
        if not map_file_exists(self.map_save_path):
            print("Loading VLMap failed because the file doesn't exist.")
            return False
        (self.mapped_iter_list, self.grid_feat, self.grid_pos, self.weight, self.occupied_ids, self.grid_rgb,
         self.pcd_min, self.pcd_max, self.cs) = load_3d_map_multi_floor(self.map_save_path)
        self.set_map_arrays(self.grid_feat)
        return True
