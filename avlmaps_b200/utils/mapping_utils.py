"""Host-side geometry and map I/O with the reference's names and signatures
(reference avlmaps/utils/mapping_utils.py).  Scalar fp64 pose/camera math stays on the host exactly as
the reference computes it (numpy + scipy), because the kernels take these matrices as inputs
(SURVEY.md section 0.6); per-point work happens in csrc/build_path.cu."""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Set, Tuple

import numpy as np


def cvt_pose_vec2tf(pos_quat_vec: np.ndarray) -> np.ndarray:
    """(px, py, pz, qx, qy, qz, qw) -> 4x4.  Reference mapping_utils.py:18-26."""
    from scipy.spatial.transform import Rotation as R

    pose_tf = np.eye(4)
    pose_tf[:3, 3] = pos_quat_vec[:3].flatten()
    pose_tf[:3, :3] = R.from_quat(pos_quat_vec[3:].flatten()).as_matrix()
    return pose_tf


def load_depth_npy(depth_filepath) -> np.ndarray:
    """Reference mapping_utils.py (load_depth_npy): depth in metres, (H, W) float32."""
    with open(depth_filepath, "rb") as f:
        return np.load(f)


def load_depth_img(depth_filepath):
    """Reference mapping_utils.py:93-94 (16-bit PNG, millimetres)."""
    import cv2

    return cv2.imread(depth_filepath, cv2.IMREAD_UNCHANGED)


def get_sim_cam_mat(h: int, w: int) -> np.ndarray:
    """Reference mapping_utils.py:591-596."""
    cam_mat = np.eye(3)
    cam_mat[0, 0] = cam_mat[1, 1] = w / 2.0
    cam_mat[0, 2] = w / 2.0
    cam_mat[1, 2] = h / 2.0
    return cam_mat


def base_pos2grid_id_3d(gs, cs, x_base, y_base, z_base):
    """Reference mapping_utils.py:345-349 (scalar helper; the kernels do this per point)."""
    row = int(gs / 2 - int(x_base / cs))
    col = int(gs / 2 - int(y_base / cs))
    h = int(z_base / cs)
    return [row, col, h]


def grid_id2base_pos_3d(row, col, height, cs, gs):
    base_x = (gs / 2 - row) * cs
    base_y = (gs / 2 - col) * cs
    base_z = height * cs
    return [base_x, base_y, base_z]


# ---------------------------------------------------------------------------------------------- map files
# The reference writes HDF5 through h5py (mapping_utils.py:469-505).  h5py is used when it is importable;
# otherwise `h5lite` (this package's dependency-free reader / writer of the same file structure) reads and writes
# the very same datasets, so `vlmaps.h5df` files interchange with the reference either way.  Maps saved by earlier
# versions of this package as an `.npz` twin (`<path>.npz`) still load.
from . import h5lite

_FIELDS = ("mapped_iter_list", "grid_feat", "grid_pos", "weight", "occupied_ids", "grid_rgb", "init_height_id")


def _have_h5py() -> bool:
    try:
        import h5py  # noqa: F401

        return hasattr(h5py, "File")
    except Exception:  # noqa: BLE001
        return False


def _npz_twin(path) -> Path:
    return Path(str(path) + ".npz")


def map_file_exists(path) -> bool:
    return Path(path).exists() or _npz_twin(path).exists()


def _write_datasets(save_path, data: dict) -> None:
    if _have_h5py():
        import h5py

        with h5py.File(save_path, "w") as f:
            for k, v in data.items():
                f.create_dataset(k, data=v)
    else:
        h5lite.write_file(save_path, data)


def _read_datasets(map_path, fields) -> dict:
    """{name: array} of the datasets of `fields` present in the file (scalars as 0-d arrays)."""
    if Path(map_path).exists():
        if _have_h5py():
            import h5py

            with h5py.File(map_path, "r") as f:
                return {k: np.asarray(f[k][()]) for k in fields if k in f}
        return h5lite.read_file(map_path, list(fields))
    twin = _npz_twin(map_path)
    if not twin.exists():
        raise FileNotFoundError(f"{map_path} does not exist")
    with np.load(twin) as z:
        return {k: z[k] for k in fields if k in z.files}


def save_3d_map(save_path, grid_feat: np.ndarray, grid_pos: np.ndarray, weight: np.ndarray, occupied_ids: np.ndarray,
                mapped_iter_list: Set[int], grid_rgb: Optional[np.ndarray] = None, init_height_id: Optional[int] = None) -> None:
    """Reference mapping_utils.py:469-505: same arguments, same dataset names."""
    data = {
        "mapped_iter_list": np.array(list(mapped_iter_list), dtype=np.int32),
        "grid_feat": grid_feat, "grid_pos": grid_pos, "weight": weight, "occupied_ids": occupied_ids,
    }
    if init_height_id is not None:
        data["init_height_id"] = np.array(init_height_id, dtype=np.int32)
    if grid_rgb is not None:
        data["grid_rgb"] = grid_rgb
    _write_datasets(save_path, data)


def load_3d_map(map_path, mmap_feat: bool = False) -> Tuple:
    """Reference mapping_utils.py:508-541: returns the 6-tuple (7 with init_height_id).

    mmap_feat=True returns `grid_feat` as a read-only memory map of the file instead of a copy (contiguous float32
    storage only, which is what the reference and this package write): the device upload then reads the page cache
    directly and host RAM holds no second copy of a multi-GB map.  The map file must not be rewritten in place while
    the array is alive (this package's writer renames a temporary, which is safe; h5py's "w" mode truncates)."""
    d = _read_datasets(map_path, _FIELDS if not mmap_feat else tuple(k for k in _FIELDS if k != "grid_feat"))
    if mmap_feat:
        feat = None
        if Path(map_path).exists():
            with h5lite.File(map_path) as f:
                ds = f["grid_feat"]
                if ds.offset is not None and ds.size and ds.dtype == np.dtype("<f4"):
                    feat = ds.memmap()
        d["grid_feat"] = feat if feat is not None else _read_datasets(map_path, ("grid_feat",))["grid_feat"]
    mapped_iter_list = d["mapped_iter_list"].tolist()
    out = (mapped_iter_list, d["grid_feat"], d["grid_pos"], d["weight"], d["occupied_ids"], d.get("grid_rgb"))
    if "init_height_id" in d:
        return out + (d["init_height_id"],)
    return out


_FIELDS_MF = ("mapped_iter_list", "grid_feat", "grid_pos", "weight", "occupied_ids", "grid_rgb", "pcd_min", "pcd_max", "cs")


def save_3d_map_multi_floor(save_path, grid_feat, grid_pos, weight, grid_rgb, occupied_ids, mapped_iter_set, pcd_min,
                            pcd_max, cs) -> None:
    """VLMapBuilderMultiFloor.save_3d_map (reference vlmap_builder_multi_floor.py:370-393): same dataset names."""
    _write_datasets(save_path, {
        "mapped_iter_list": np.array(list(mapped_iter_set), dtype=np.int32),
        "grid_feat": grid_feat, "grid_pos": grid_pos, "weight": weight, "occupied_ids": occupied_ids,
        "grid_rgb": grid_rgb, "pcd_min": np.asarray(pcd_min), "pcd_max": np.asarray(pcd_max), "cs": np.asarray(cs),
    })


def load_3d_map_multi_floor(map_path) -> Tuple:
    """VLMapBuilderMultiFloor.load_3d_map (reference vlmap_builder_multi_floor.py:243-255): the 9-tuple."""
    d = _read_datasets(map_path, _FIELDS_MF)
    missing = [k for k in _FIELDS_MF if k not in d]
    if missing:
        raise KeyError(f"{map_path} lacks the multi-floor datasets {missing}")
    return (d["mapped_iter_list"].tolist(), d["grid_feat"], d["grid_pos"], d["weight"], d["occupied_ids"], d["grid_rgb"],
            d["pcd_min"], d["pcd_max"], d["cs"][()])


def save_clip_sparse_map(save_path, clip_sparse_map: np.ndarray, robot_pose_list) -> None:
    """Reference mapping_utils.py:637-640."""
    _write_datasets(save_path, {"clip_sparse_map": np.asarray(clip_sparse_map), "robot_pose_list": np.asarray(robot_pose_list)})


def load_clip_sparse_map(load_path):
    """Reference mapping_utils.py:643-647."""
    d = _read_datasets(load_path, ("clip_sparse_map", "robot_pose_list"))
    return d["clip_sparse_map"], d["robot_pose_list"]
