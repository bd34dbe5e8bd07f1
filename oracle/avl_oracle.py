"""ORACLE -- test infrastructure only.

CPU restatement of the reference's two hot paths.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs may import this module, and only as the checker
or the timed CPU baseline; nothing under avlmaps_b200/ imports it.

Pinned: tests/golden/*.npz were produced by running the UNMODIFIED reference through
oracle/ref_shim.py in the build container (tests/golden/gen_golden.py); tests/test_oracle_golden.py
checks every function here against those vectors.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_DIR = HERE / "_lib"
LIB = LIB_DIR / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """gcc-compile oracle/build_oracle.c (no FMA contraction: the geometry must round like numpy)."""
    src = HERE / "build_oracle.c"
    if LIB.exists() and not force and LIB.stat().st_mtime >= src.stat().st_mtime:
        return LIB
    LIB_DIR.mkdir(exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", str(src), "-o", str(LIB), "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.oracle_builder_create.restype = vp
        L.oracle_builder_create.argtypes = [i32, i32, dbl, i32, i64]
        L.oracle_builder_destroy.argtypes = [vp]
        L.oracle_builder_num_voxels.restype = i64
        L.oracle_builder_num_voxels.argtypes = [vp]
        L.oracle_builder_num_accepted.restype = i64
        L.oracle_builder_num_accepted.argtypes = [vp]
        for nm in ("grid_feat", "grid_pos", "weight", "occupied_ids", "grid_rgb"):
            f = getattr(L, "oracle_builder_" + nm)
            f.restype = vp
            f.argtypes = [vp]
        L.oracle_builder_add_frame.restype = C.c_int
        L.oracle_builder_add_frame.argtypes = [vp, vp, i32, i32, vp, i32, i32, vp, vp, i32, vp, vp, vp, vp, dbl, dbl]
        L.oracle_builder_add_frame_u16.restype = C.c_int
        L.oracle_builder_add_frame_u16.argtypes = L.oracle_builder_add_frame.argtypes
        L.oracle_builder_create_global.restype = vp
        L.oracle_builder_create_global.argtypes = [i32, i32, i32, dbl, vp, i32, i64]
        L.oracle_builder_num_oob.restype = i64
        L.oracle_builder_num_oob.argtypes = [vp]
        L.oracle_builder_import.restype = C.c_int
        L.oracle_builder_import.argtypes = [vp, vp, vp, vp, vp, vp, i64]
        L.oracle_frame_bounds.restype = i64
        L.oracle_frame_bounds.argtypes = [vp, C.c_int, i32, i32, vp, i32, vp, vp, dbl, dbl, vp]
        L.oracle_scores.argtypes = [vp, i64, i32, vp, i32, vp, C.c_int, vp]
        L.oracle_argmax.argtypes = [vp, i64, i32, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# =============================================================================== index path
def ref_scores_fp32(feat: np.ndarray, q: np.ndarray) -> np.ndarray:
    """The reference's literal operation: `map_feats @ text_feats.T` in float32 through BLAS
    (avlmaps/utils/clip_utils.py:227-229, 236-240).  Summation order is OpenBLAS'."""
    return feat.reshape((-1, feat.shape[-1])) @ q.T


def scores(feat: np.ndarray, q: np.ndarray, scale=None, normalize: bool = False, use_c: bool = False) -> np.ndarray:
    """Canonical scores: fp64-accumulated dot rounded once to fp32, then the fp32 multiplies by
    1/||a_i|| (normalize) and scale_q.  Restates clip_utils.py:229 (scale=None, normalize=False), the
    cosine of avlmaps/map/area_map.py:102,118 (normalize=True) and the scaled audio logits of
    avlmaps/map/sound_map.py:108-109.  use_c=True runs the plain-C loop (k ascending)."""
    feat = np.ascontiguousarray(feat, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    sc = None if scale is None else np.ascontiguousarray(scale, np.float32)
    if use_c:
        out = np.empty((feat.shape[0], q.shape[0]), np.float32)
        lib().oracle_scores(_p(feat), feat.shape[0], feat.shape[1], _p(q), q.shape[0], _p(sc), int(normalize), _p(out))
        return out
    s = (feat.astype(np.float64) @ q.astype(np.float64).T).astype(np.float32)
    if normalize:
        nrm = np.sqrt(np.einsum("ij,ij->i", feat.astype(np.float64), feat.astype(np.float64))).astype(np.float32)
        with np.errstate(divide="ignore"):
            inv = np.where(nrm > 0, np.float32(1.0) / nrm, np.float32(0.0)).astype(np.float32)
        s = s * inv[:, None]
    if sc is not None:
        s = s * sc[None, :]
    return s.astype(np.float32)


def argmax(sc: np.ndarray) -> np.ndarray:
    """avlmaps/map/vlmap.py:123 -- np.argmax(scores_mat, axis=1): first maximum wins."""
    return np.argmax(sc, axis=1).astype(np.int32)


def index_mask(sc: np.ndarray, cat_id: int) -> np.ndarray:
    """avlmaps/map/vlmap.py:123-124"""
    return np.argmax(sc, axis=1) == cat_id


def topk_vector(v: np.ndarray, k: int):
    """k best entries by (value desc, index asc); k=1 is `np.argmax(heat)`
    (avlmaps/robot/habitat_lang_robot.py:427-430), k>1 the argsort retrieval of
    avlmaps/utils/clip_utils.py:86-93 with a defined tie rule."""
    n = v.shape[0]
    kk = min(k, n)
    idx = np.full(k, -1, np.int64)
    val = np.full(k, -np.inf, np.float32)
    if kk:
        if n > 4 * kk:
            part = np.argpartition(-v.astype(np.float64), kk - 1)[:kk]
            thr = v[part].min()
            cand = np.nonzero(v >= thr)[0]
        else:
            cand = np.arange(n)
        order = cand[np.lexsort((cand, -v[cand].astype(np.float64)))][:kk]
        idx[:kk] = order
        val[:kk] = v[order]
    return idx, val


def topk(sc: np.ndarray, k: int):
    nq = sc.shape[1]
    idx = np.empty((nq, k), np.int64)
    val = np.empty((nq, k), np.float32)
    for j in range(nq):
        idx[j], val[j] = topk_vector(np.ascontiguousarray(sc[:, j]), k)
    return idx, val


def minmax(v: np.ndarray) -> np.ndarray:
    """(x - min) / (max - min) in float32: avlmaps/map/sound_map.py:151-152,
    avlmaps/robot/habitat_lang_robot.py:213-214, avlmaps/map/avlmap.py:81."""
    v = v.astype(np.float32)
    return (v - np.min(v)) / (np.max(v) - np.min(v))


FUSE_PRODUCT, FUSE_MAX, FUSE_SUM = 0, 1, 2


def fuse_heat(sa: np.ndarray, sb: np.ndarray, combine: int = FUSE_PRODUCT) -> np.ndarray:
    """Per-voxel fused heat of one query pair: min-max each modality, then combine (product per the
    paper; max = habitat_lang_robot.py:223 / avlmap.py:95; sum = avlmap.py:129)."""
    ha, hb = minmax(sa), minmax(sb)
    if combine == FUSE_PRODUCT:
        return (ha * hb).astype(np.float32)
    if combine == FUSE_MAX:
        return np.maximum(ha, hb).astype(np.float32)
    return (ha + hb).astype(np.float32)


def fuse_topk(sa: np.ndarray, sb: np.ndarray, combine: int, k: int):
    n_pairs = sa.shape[1]
    idx = np.empty((n_pairs, k), np.int64)
    val = np.empty((n_pairs, k), np.float32)
    for j in range(n_pairs):
        idx[j], val[j] = topk_vector(fuse_heat(sa[:, j], sb[:, j], combine), k)
    return idx, val


# =============================================================================== build path: host prep
def cvt_pose_vec2tf(pos_quat_vec: np.ndarray) -> np.ndarray:
    """avlmaps/utils/mapping_utils.py:18-26 (scipy Rotation.from_quat, xyzw)."""
    from scipy.spatial.transform import Rotation as R

    pose_tf = np.eye(4)
    pose_tf[:3, 3] = pos_quat_vec[:3].flatten()
    pose_tf[:3, :3] = R.from_quat(pos_quat_vec[3:].flatten()).as_matrix()
    return pose_tf


def setup_transforms(pose_info: dict):
    """avlmaps/map/map.py:54-68"""
    base2cam_tf = np.eye(4)
    base2cam_tf[:3, :3] = np.array([pose_info["base2cam_rot"]]).reshape((3, 3))
    base2cam_tf[1, 3] = pose_info["camera_height"]
    base_transform = np.eye(4)
    base_transform[0, :3] = pose_info["base_forward_axis"]
    base_transform[1, :3] = pose_info["base_left_axis"]
    base_transform[2, :3] = pose_info["base_up_axis"]
    return base2cam_tf, base_transform


def frame_transforms(poses: np.ndarray, base2cam_tf: np.ndarray, base_transform: np.ndarray):
    """pc_transform of every frame: avlmaps/map/vlmap_builder.py:67-74 (init), :106-108, :133."""
    init_base_tf = base_transform @ cvt_pose_vec2tf(poses[0]) @ np.linalg.inv(base_transform)
    inv_init_base_tf = np.linalg.inv(init_base_tf)
    out = []
    for pv in poses:
        base_pose = base_transform @ cvt_pose_vec2tf(pv) @ np.linalg.inv(base_transform)
        tf = inv_init_base_tf @ base_pose
        out.append(tf @ base_transform @ base2cam_tf)
    return out


def get_sim_cam_mat(h: int, w: int) -> np.ndarray:
    """avlmaps/utils/mapping_utils.py:591-596"""
    cam_mat = np.eye(3)
    cam_mat[0, 0] = cam_mat[1, 1] = w / 2.0
    cam_mat[0, 2] = w / 2.0
    cam_mat[1, 2] = h / 2.0
    return cam_mat


def sample_order(n_pixels: int, rate: int) -> np.ndarray:
    """avlmaps/map/vlmap_builder.py:275-277: global-RNG shuffle of arange, then [::rate].
    Call np.random.seed(...) once before the first frame, like a user of the reference would."""
    shuffle_mask = np.arange(n_pixels)
    np.random.shuffle(shuffle_mask)
    return shuffle_mask[::rate].astype(np.int32)


class BuildOracle:
    """Sequential restatement of the fusion loop (avlmaps/map/vlmap_builder.py:129-178) in C."""

    def __init__(self, gs: int, vh: int, cs: float, dim: int, capacity: int | None = None):
        self.gs, self.vh, self.cs, self.dim = gs, vh, cs, dim
        self.shape = (gs, gs, vh)
        self.capacity = int(capacity if capacity is not None else gs * gs)  # vlmap_builder.py:202
        self._h = lib().oracle_builder_create(gs, vh, cs, dim, self.capacity)
        self._keep = []

    @classmethod
    def global_grid(cls, n_row: int, n_col: int, n_height: int, cs: float, pcd_min, dim: int,
                    capacity: int | None = None):
        """Grid of VLMapBuilderMultiFloor._init_map (vlmap_builder_multi_floor.py:217-241)."""
        self = cls.__new__(cls)
        self.gs, self.vh, self.cs, self.dim = n_row, n_height, cs, dim
        self.shape = (n_row, n_col, n_height)
        self.capacity = int(capacity if capacity is not None else n_row * n_col)  # :225
        pm = np.ascontiguousarray(pcd_min, np.float64)
        self._h = lib().oracle_builder_create_global(n_row, n_col, n_height, cs, _p(pm), dim, self.capacity)
        self._keep = []
        return self

    def import_state(self, grid_feat, grid_pos, weight, grid_rgb, occupied_ids):
        """_init_map's reload (vlmap_builder.py:212-222)."""
        f, p = np.ascontiguousarray(grid_feat, np.float32), np.ascontiguousarray(grid_pos, np.int32)
        w, o = np.ascontiguousarray(weight, np.float32), np.ascontiguousarray(occupied_ids, np.int32)
        c = None if grid_rgb is None else np.ascontiguousarray(grid_rgb, np.uint8)
        assert o.shape == tuple(self.shape)
        if lib().oracle_builder_import(self._h, _p(f), _p(p), _p(w), _p(c), _p(o), f.shape[0]) != 0:
            raise RuntimeError("oracle import: capacity too small or builder not fresh")

    @property
    def num_oob(self) -> int:
        return int(lib().oracle_builder_num_oob(self._h))

    def add_frame(self, depth, feat_chw, rgb, sample_idx, kinv, k, kfeat, tf, min_depth=0.1, max_depth=6.0):
        u16 = np.asarray(depth).dtype == np.uint16  # multi-floor: millimetres, / 1000.0 inside
        depth = np.ascontiguousarray(depth, np.uint16 if u16 else np.float32)
        feat_chw = np.ascontiguousarray(feat_chw, np.float32)
        assert feat_chw.ndim == 4 and feat_chw.shape[0] == 1 and feat_chw.shape[1] == self.dim
        rgb = None if rgb is None else np.ascontiguousarray(rgb, np.uint8)
        sidx = None if sample_idx is None else np.ascontiguousarray(sample_idx, np.int32)
        n = depth.size if sidx is None else sidx.size
        mats = [np.ascontiguousarray(m, np.float64) for m in (kinv, k, kfeat, tf)]
        fn = lib().oracle_builder_add_frame_u16 if u16 else lib().oracle_builder_add_frame
        rc = fn(self._h, _p(depth), depth.shape[0], depth.shape[1], _p(feat_chw),
                feat_chw.shape[2], feat_chw.shape[3], _p(rgb), _p(sidx), n,
                _p(mats[0]), _p(mats[1]), _p(mats[2]), _p(mats[3]), float(min_depth), float(max_depth))
        if rc != 0:
            raise RuntimeError("oracle capacity exceeded (_reserve_map_space is not restated)")

    @property
    def num_voxels(self) -> int:
        return int(lib().oracle_builder_num_voxels(self._h))

    @property
    def num_accepted(self) -> int:
        return int(lib().oracle_builder_num_accepted(self._h))

    def export(self):
        """arrays[:max_id] + occupied_ids, like _save_3d_map (vlmap_builder.py:313-327)."""
        L, v = lib(), self.num_voxels

        def view(ptr, ctype, shape):
            n = int(np.prod(shape))
            if n == 0:
                return np.zeros(shape, np.dtype(ctype))
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).reshape(shape).copy()

        return dict(
            grid_feat=view(L.oracle_builder_grid_feat(self._h), C.c_float, (v, self.dim)),
            grid_pos=view(L.oracle_builder_grid_pos(self._h), C.c_int32, (v, 3)),
            weight=view(L.oracle_builder_weight(self._h), C.c_float, (v,)),
            occupied_ids=view(L.oracle_builder_occupied_ids(self._h), C.c_int32, self.shape),
            grid_rgb=view(L.oracle_builder_grid_rgb(self._h), C.c_uint8, (v, 3)),
        )

    def close(self):
        if self._h:
            lib().oracle_builder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def build_map(map_config: dict, poses: np.ndarray, depths, rgbs, feats, sample_idx, capacity=None, resume=None):
    """Whole build from the same inputs ref_shim.ref_build takes (vlmap_builder.py:54-185).
    `resume`: dict of a saved map (grid_feat, grid_pos, weight, grid_rgb, occupied_ids) reloaded first (:212-222)."""
    cs, gs = map_config["cell_size"], map_config["grid_size"]
    vh = int(map_config["pose_info"]["camera_height"] / cs)  # vlmap_builder.py:201
    base2cam_tf, base_transform = setup_transforms(map_config["pose_info"])
    tfs = frame_transforms(poses, base2cam_tf, base_transform)
    calib = np.array(map_config["cam_calib_mat"], dtype=np.float64).reshape((3, 3))  # vlmap_builder.py:98
    kinv = np.linalg.inv(calib)  # mapping_utils.py:237
    D = feats[0].shape[1]
    b = BuildOracle(gs, vh, cs, D, capacity)
    if resume is not None:
        b.import_state(resume["grid_feat"], resume["grid_pos"], resume["weight"], resume.get("grid_rgb"), resume["occupied_ids"])
    for i, tf in enumerate(tfs):
        kfeat = get_sim_cam_mat(feats[i].shape[2], feats[i].shape[3])  # vlmap_builder.py:126
        b.add_frame(depths[i], feats[i], None if rgbs is None else rgbs[i], sample_idx[i], kinv, calib, kfeat, tf)
    out = b.export()
    out["num_accepted"] = b.num_accepted
    b.close()
    return out


# =============================================================================== multi-floor build
HABITAT2CAM_ROT_TF = np.diag([1.0, -1.0, -1.0, 1.0])  # vlmap_builder_multi_floor.py:77-79


def frame_bounds(depth, sample_idx, kinv, tf, minmax=None, min_depth=0.1, max_depth=100.0):
    """One frame of the first pass of create_global_map (vlmap_builder_multi_floor.py:97-118): merges the
    min / max of the transformed valid sampled points into `minmax` (6 float64: min xyz, max xyz)."""
    u16 = np.asarray(depth).dtype == np.uint16
    depth = np.ascontiguousarray(depth, np.uint16 if u16 else np.float32)
    if minmax is None:
        minmax = np.array([np.inf] * 3 + [-np.inf] * 3)
    sidx = None if sample_idx is None else np.ascontiguousarray(sample_idx, np.int32)
    n = depth.size if sidx is None else sidx.size
    kinv, tf = np.ascontiguousarray(kinv, np.float64), np.ascontiguousarray(tf, np.float64)
    cnt = lib().oracle_frame_bounds(_p(depth), int(u16), depth.shape[0], depth.shape[1], _p(sidx), n, _p(kinv), _p(tf),
                                    float(min_depth), float(max_depth), _p(minmax))
    return minmax, int(cnt)


def global_grid_size(pcd_min, pcd_max, cs):
    """grid_size = ceil((pcd_max - pcd_min) / cs + 1) as (x, y, z) (vlmap_builder_multi_floor.py:222);
    occupied_ids is grid_size[[0, 2, 1]] = (n_row, n_col, n_height) (:224)."""
    gsz = np.ceil((np.asarray(pcd_max) - np.asarray(pcd_min)) / cs + 1).astype(int)
    return int(gsz[0]), int(gsz[2]), int(gsz[1])


def build_map_multi_floor(map_config: dict, cam_poses, depths_mm, rgbs, feats, sample_idx_pass1, sample_idx_pass2):
    """Whole create_global_map (vlmap_builder_multi_floor.py:60-199) from the inputs
    ref_shim.ref_build_multi_floor takes; sample lists are those of the frames with frame_i % skip_frame == 0."""
    cs, skip = map_config["cell_size"], map_config["skip_frame"]
    calib = np.array(map_config["cam_calib_mat"], dtype=np.float64).reshape((3, 3))
    kinv = np.linalg.inv(calib)
    used = [i for i in range(len(depths_mm)) if i % skip == 0]
    tfs = {i: np.asarray(cam_poses[i], np.float64).reshape(4, 4) @ HABITAT2CAM_ROT_TF for i in used}  # :105,141
    mm = None
    for j, i in enumerate(used):
        mm, _ = frame_bounds(depths_mm[i], sample_idx_pass1[j], kinv, tfs[i], mm)
    pcd_min, pcd_max = mm[:3].copy(), mm[3:].copy()
    n_row, n_col, n_height = global_grid_size(pcd_min, pcd_max, cs)
    D = feats[0].shape[1]
    b = BuildOracle.global_grid(n_row, n_col, n_height, cs, pcd_min, D)
    for j, i in enumerate(used):
        kfeat = get_sim_cam_mat(feats[i].shape[2], feats[i].shape[3])  # :135
        b.add_frame(depths_mm[i], feats[i], None if rgbs is None else rgbs[i], sample_idx_pass2[j], kinv, calib, kfeat,
                    tfs[i], min_depth=0.1, max_depth=100)  # :138
    out = b.export()
    out.update(num_accepted=b.num_accepted, num_oob=b.num_oob, pcd_min=pcd_min, pcd_max=pcd_max)
    b.close()
    return out


# =============================================================================== heat
def heatmap_from_mask_3d(grid_pos: np.ndarray, mask: np.ndarray, cell_size: float = 0.05, decay_rate: float = 0.01):
    """avlmaps/utils/visualize_utils.py:29-49, vectorised over the non-target voxels in blocks.
    heat = 1 on target voxels, clip(1 - min_dist/cell_size * decay, 0, 1) elsewhere."""
    pc = grid_pos.astype(np.float64)
    mask = mask.astype(bool)
    target = pc[mask]
    heat = np.ones(pc.shape[0], np.float32)
    other = np.nonzero(~mask)[0]
    if target.shape[0] == 0:
        # np.argmin of an empty array raises in the reference; nothing to restate
        raise ValueError("empty target mask")
    for s in range(0, other.size, 2048):
        ids = other[s:s + 2048]
        d = np.sqrt(((pc[ids, None, :] - target[None, :, :]) ** 2).sum(-1)) / cell_size
        heat[ids] = np.clip(1 - d.min(1) * decay_rate, 0, 1)
    return heat


def area_heat_2d(shape, cells, scores, decay_rate: float = 0.1):
    """avlmaps/map/avlmap.py:78-98 restated line by line (scipy EDT per frame); cells[i] = (row, col) or None."""
    from scipy.ndimage import distance_transform_edt

    dist_map = np.zeros(shape, dtype=np.float32)
    for i, cell in enumerate(cells):
        tmp_dist_map = np.zeros_like(dist_map, dtype=np.float32)
        if cell is None:
            continue
        row, col = cell
        s = scores[i]
        tmp_dist_map[row, col] = s
        dists = distance_transform_edt(tmp_dist_map == 0)
        tmp = np.ones_like(dists) * s - (dists * decay_rate)
        tmp_dist_map = np.clip(tmp, 0, 1)
        dist_map = np.where(dist_map > tmp_dist_map, dist_map, tmp_dist_map)
    return (dist_map - np.min(dist_map)) / (np.max(dist_map) - np.min(dist_map))


def sound_heat_2d(shape, cells_per_segment, probabilities, decay_rate: float = 0.01):
    """avlmaps/map/avlmap.py:111-133 restated line by line."""
    from scipy.ndimage import distance_transform_edt

    dist_map = np.zeros(shape, dtype=np.float32)
    for loc_i, cells in enumerate(cells_per_segment):
        tmp_dist_map = np.zeros_like(dist_map, dtype=np.float32)
        for row, col in cells:
            tmp_dist_map[row, col] = probabilities[loc_i]
        con = probabilities[loc_i]
        dists = distance_transform_edt(tmp_dist_map == 0)
        reduct = con * dists * decay_rate
        tmp = np.ones_like(tmp_dist_map) * con - reduct
        tmp_dist_map = np.where(tmp < 0, np.zeros_like(tmp), tmp)
        dist_map += tmp_dist_map
    return (dist_map - np.min(dist_map)) / (np.max(dist_map) - np.min(dist_map))


def lift_heat_2d_to_3d(heatmap_2d, occupied_ids, n_vox):
    """avlmaps/map/avlmap.py:100-109 / 135-144 (the Python loop over occupied cells)."""
    heatmap_3d = np.zeros(n_vox, dtype=np.float32)
    rows, cols, heights = np.where(occupied_ids != -1)
    for row, col, heigh in zip(rows, cols, heights):
        heatmap_3d[occupied_ids[row, col, heigh]] = heatmap_2d[row, col]
    return heatmap_3d


def image_heat(grid_pos, row, col, camera_height=1.5, cs=0.05, decay_rate=0.01):
    """avlmaps/map/avlmap.py:156-162: planar distance decay around the localised query image."""
    height = camera_height / cs
    pos = np.array([row, col, height])
    sim_mat = np.zeros((grid_pos.shape[0], 1))
    dists = np.linalg.norm((grid_pos - pos)[:, :2], axis=1)
    sim_mat[:, 0] = np.clip(1.0 - decay_rate * dists, 0, 1)
    return np.max(sim_mat, axis=1).flatten()
