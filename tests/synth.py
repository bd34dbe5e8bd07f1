"""Deterministic synthetic inputs shared by the golden generator, the tests and bench.py
(SURVEY.md section 8d).  Everything is a function of integer seeds."""
from __future__ import annotations

import numpy as np

DEFAULT_POSE_INFO = {  # /root/reference/config/map_config/vlmaps.yaml:2-9
    "pose_type": "mobile_base",
    "camera_height": 1.5,
    "base2cam_rot": [1, 0, 0, 0, -1, 0, 0, 0, -1],
    "base_forward_axis": [0, 0, -1],
    "base_left_axis": [-1, 0, 0],
    "base_up_axis": [0, 1, 0],
}


def index_inputs(n: int, d: int, nq: int, seed: int = 0, unit_rows: bool = False):
    """LSeg-like map: rows of norm ~ 14.29 * alpha (un-normalised, like real fused maps) and unit queries."""
    feat = np.random.default_rng(seed).standard_normal((n, d), dtype=np.float32)
    if unit_rows:
        feat /= np.linalg.norm(feat, axis=1, keepdims=True)
    else:
        s = (14.2857 * np.random.default_rng(seed + 1).uniform(0.05, 1.0, n)).astype(np.float32)
        feat *= s[:, None]
    q = np.random.default_rng(seed + 2).standard_normal((nq, d))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.ascontiguousarray(feat, np.float32), np.ascontiguousarray(q, np.float32)


def circle_poses(n_frames: int, radius: float = 2.0, height: float = 0.0) -> np.ndarray:
    """Habitat base poses (px,py,pz,qx,qy,qz,qw) on a circle in the x-z plane, heading tangent."""
    from scipy.spatial.transform import Rotation as R

    out = np.zeros((n_frames, 7))
    for i in range(n_frames):
        th = 2 * np.pi * i / max(n_frames, 1)
        out[i, 0] = radius * np.cos(th) - radius
        out[i, 1] = height
        out[i, 2] = radius * np.sin(th)
        out[i, 3:] = R.from_euler("y", -th).as_quat()
    return out


def map_config(gs: int, cs: float, camera_height: float, calib, rate: int) -> dict:
    pi = dict(DEFAULT_POSE_INFO)
    pi["camera_height"] = camera_height
    return {"map_type": "vlmap", "pose_info": pi, "cam_calib_mat": [float(x) for x in np.asarray(calib).flatten()],
            "grid_size": gs, "cell_size": cs, "depth_sample_rate": rate}


def build_inputs(n_frames: int, h: int, w: int, fh: int, fw: int, d: int, seed: int = 0, pool: int = 0,
                 depth_lo: float = 0.05, depth_hi: float = 6.5):
    """Per-frame depth (some pixels outside [0.1, 6]), RGB, and (1,D,FH,FW) features of LSeg-like norm."""
    depths, rgbs, feats = [], [], []
    npool = pool if pool > 0 else n_frames
    fpool = []
    for i in range(npool):
        f = np.random.default_rng(200 + seed * 1000 + i).standard_normal((1, d, fh, fw), dtype=np.float32)
        f *= np.float32(14.2857 / np.sqrt(d))
        fpool.append(f)
    for i in range(n_frames):
        rng = np.random.default_rng(100 + seed * 1000 + i)
        depths.append(rng.uniform(depth_lo, depth_hi, (h, w)).astype(np.float32))
        rgbs.append(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        feats.append(fpool[i % npool])
    return depths, rgbs, feats


def global_cam_poses(n_frames: int, radius: float = 0.5, floors: int = 2, floor_height: float = 1.2):
    """Camera poses (4, 4) in the global habitat frame (y up, camera looks down -z) for the multi-floor
    builder: a circle per floor, yaw tangent, a small pitch so rows of the grid are not axis-aligned."""
    from scipy.spatial.transform import Rotation as R

    out = []
    for i in range(n_frames):
        th = 2 * np.pi * i / max(n_frames, 1)
        tf = np.eye(4)
        tf[:3, :3] = (R.from_euler("y", th) * R.from_euler("x", 0.1 * np.sin(3 * th))).as_matrix()
        tf[:3, 3] = [radius * np.cos(th), 1.5 + floor_height * (i * floors // max(n_frames, 1)), radius * np.sin(th)]
        out.append(tf)
    return out


def multi_floor_config(cs: float, calib, rate: int, skip_frame: int = 1) -> dict:
    pi = dict(DEFAULT_POSE_INFO)
    pi["pose_type"] = "global"
    pi["building_init_height"] = 0.0
    return {"map_type": "vlmap_openmap", "pose_info": pi, "cam_calib_mat": [float(x) for x in np.asarray(calib).flatten()],
            "grid_size": 1000, "cell_size": cs, "depth_sample_rate": rate, "skip_frame": skip_frame}


def multi_floor_inputs(n_frames: int, h: int, w: int, fh: int, fw: int, d: int, seed: int = 0,
                       depth_lo_mm: int = 50, depth_hi_mm: int = 3000):
    """uint16 millimetre depth (some pixels below min_depth = 0.1 m), RGB, (1, D, FH, FW) features."""
    depths, rgbs, feats = [], [], []
    for i in range(n_frames):
        rng = np.random.default_rng(300 + seed * 1000 + i)
        depths.append(rng.integers(depth_lo_mm, depth_hi_mm, (h, w)).astype(np.uint16))
        rgbs.append(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        f = np.random.default_rng(400 + seed * 1000 + i).standard_normal((1, d, fh, fw), dtype=np.float32)
        feats.append(f * np.float32(14.2857 / np.sqrt(d)))
    return depths, rgbs, feats


def crc_text_encoder(dim: int):
    """Deterministic stand-in for a CLIP text tower: features are a function of the text's CRC32 (Python's hash() is
    salted per process, so it cannot pin golden vectors).  Rows are NOT normalised, like encode_text's output."""
    import zlib

    def enc(texts):
        return np.stack([np.random.default_rng(zlib.crc32(t.encode())).standard_normal(dim) for t in texts]).astype(np.float32)

    return enc
