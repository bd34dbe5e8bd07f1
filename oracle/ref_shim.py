"""ORACLE -- test infrastructure only (see oracle/README.md).

Loads the UNMODIFIED reference files from /root/reference and runs the reference's own hot-path
functions, so that golden vectors can be generated in the build container (the reference tree does
not exist on the GPU box; nothing that runs there imports this module).

The package `avlmaps.map` cannot be imported as a whole (its __init__ pulls hloc, librosa,
audioclip, shapely ...), so single files are loaded with importlib after empty stub modules are put
in sys.modules for the third-party imports that are absent here.
Only three names are monkey-patched, all of them I/O or model loading, never arithmetic:
  VLMapBuilder._init_lseg   (would load the LSeg checkpoint)   -> sets device / clip_feat_dim
  get_lseg_feat             (would run LSeg)                   -> returns the synthetic (1,D,FH,FW) features
  save_3d_map               (would write HDF5 through h5py)    -> captures the arrays
and `get_text_feats` (would run CLIP) for the index path -> returns the synthetic embeddings.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

REF_ROOT = Path(os.environ.get("AVL_REFERENCE_ROOT", "/root/reference"))
_STUBS = ["h5py", "matplotlib", "matplotlib.patches", "matplotlib.pyplot", "clip", "open3d", "omegaconf",
          "gdown", "timm"]
_loaded = {}


def available() -> bool:
    return (REF_ROOT / "avlmaps" / "map" / "vlmap_builder.py").exists()


def _install_stubs():
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["omegaconf"].DictConfig = dict
    sys.modules["omegaconf"].OmegaConf = object
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))


def load(name: str, rel: str):
    if name in _loaded:
        return _loaded[name]
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    _install_stubs()
    spec = importlib.util.spec_from_file_location(name, REF_ROOT / rel)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _loaded[name] = mod
    return mod


class AttrDict(dict):
    """dict with attribute access, what the reference expects from an OmegaConf node."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e
        return AttrDict(v) if isinstance(v, dict) else v


# ------------------------------------------------------------------------------------------ index path
def ref_get_lseg_score(feat: np.ndarray, text_feats: np.ndarray) -> np.ndarray:
    """The reference's get_lseg_score (avlmaps/utils/clip_utils.py:196-242) with
    use_multiple_templates=False; the CLIP text encoder is replaced by `text_feats`."""
    cu = load("ref_clip_utils", "avlmaps/utils/clip_utils.py")
    names = [f"q{i}" for i in range(text_feats.shape[0] - 1)] + ["other"]
    saved = cu.get_text_feats
    cu.get_text_feats = lambda in_text, clip_model, clip_feat_dim, batch_size=64: text_feats
    try:
        return cu.get_lseg_score(None, names, feat, feat.shape[-1], use_multiple_templates=False)
    finally:
        cu.get_text_feats = saved


def ref_index_mask(scores: np.ndarray, cat_id: int) -> np.ndarray:
    """avlmaps/map/vlmap.py:123-124"""
    max_ids = np.argmax(scores, axis=1)
    return max_ids == cat_id


def ref_heatmap_from_mask_3d(grid_pos, mask, cell_size=0.05, decay_rate=0.01):
    vu = load("ref_visualize_utils", "avlmaps/utils/visualize_utils.py")
    return vu.get_heatmap_from_mask_3d(grid_pos, mask, cell_size=cell_size, decay_rate=decay_rate)


# ------------------------------------------------------------------------------------------ build path
def ref_transforms(map_config):
    """Map._setup_transforms (avlmaps/map/map.py:54-68), restated because map.py needs shapely."""
    base2cam_tf = np.eye(4)
    base2cam_tf[:3, :3] = np.array([map_config["pose_info"]["base2cam_rot"]]).reshape((3, 3))
    base2cam_tf[1, 3] = map_config["pose_info"]["camera_height"]
    base_transform = np.eye(4)
    base_transform[0, :3] = map_config["pose_info"]["base_forward_axis"]
    base_transform[1, :3] = map_config["pose_info"]["base_left_axis"]
    base_transform[2, :3] = map_config["pose_info"]["base_up_axis"]
    return base2cam_tf, base_transform


def ref_build(map_config: dict, poses: np.ndarray, depths, rgbs, feats, seed: int, resume: dict | None = None):
    """Run the reference's VLMapBuilder.create_mobile_base_map (vlmap_builder.py:54-185) on synthetic
    frames.  depths[i] (H,W) f32, rgbs[i] (H,W,3) u8 RGB, feats[i] (1,D,FH,FW) f32.
    Returns dict(grid_feat, grid_pos, weight, occupied_ids, grid_rgb, sample_idx) where sample_idx[i]
    is the pixel order the reference's global-RNG shuffle produced for frame i.
    `resume`: a previously saved map; a placeholder vlmap/vlmaps.h5df is created so that _init_map takes its
    reload branch (:212-222) and load_3d_map (h5py) is patched to return these arrays."""
    import cv2

    vb = load("ref_vlmap_builder", "avlmaps/map/vlmap_builder.py")
    cfg = AttrDict(map_config)
    base2cam_tf, base_transform = ref_transforms(map_config)
    D = feats[0].shape[1]
    captured = {}
    frame_counter = {"i": 0}
    sample_orders = []

    def fake_init_lseg(self):
        self.device = "cpu"
        self.clip_feat_dim = D
        return None, None, 480, 520, [0.5] * 3, [0.5] * 3

    def fake_get_lseg_feat(*a, **k):
        i = frame_counter["i"]
        frame_counter["i"] += 1
        return feats[i]

    def fake_save(path, grid_feat, grid_pos, weight, occupied_ids, mapped_iter_list, grid_rgb=None,
                  init_height_id=None):
        captured.update(grid_feat=np.array(grid_feat), grid_pos=np.array(grid_pos), weight=np.array(weight),
                        occupied_ids=np.array(occupied_ids), grid_rgb=np.array(grid_rgb),
                        mapped_iter_list=list(mapped_iter_list))

    # record the permutation np.random.shuffle produces inside _backproject_depth without touching it
    orig_shuffle = np.random.shuffle

    def recording_shuffle(x):
        orig_shuffle(x)
        sample_orders.append(np.array(x))

    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "rgb").mkdir()
        (td / "depth").mkdir()
        rgb_paths, depth_paths = [], []
        for i, (d, c) in enumerate(zip(depths, rgbs)):
            rp = td / "rgb" / f"{i:06d}.png"
            dp = td / "depth" / f"{i:06d}.npy"
            cv2.imwrite(str(rp), cv2.cvtColor(c, cv2.COLOR_RGB2BGR))
            np.save(dp, d)
            rgb_paths.append(rp)
            depth_paths.append(dp)
        pose_path = td / "poses.txt"
        np.savetxt(pose_path, poses)
        saved = (vb.VLMapBuilder._init_lseg, vb.get_lseg_feat, vb.save_3d_map, vb.load_3d_map)
        vb.VLMapBuilder._init_lseg = fake_init_lseg
        vb.get_lseg_feat = fake_get_lseg_feat
        vb.save_3d_map = fake_save
        if resume is not None:
            (td / "vlmap").mkdir()
            (td / "vlmap" / "vlmaps.h5df").write_bytes(b"placeholder")
            vb.load_3d_map = lambda path: (list(resume["mapped_iter_list"]), resume["grid_feat"].copy(),
                                           resume["grid_pos"].copy(), resume["weight"].copy(),
                                           resume["occupied_ids"].copy(), resume["grid_rgb"].copy())
        np.random.shuffle = recording_shuffle
        try:
            np.random.seed(seed)
            b = vb.VLMapBuilder(td, cfg, pose_path, rgb_paths, depth_paths, base2cam_tf, base_transform)
            import time as _time

            _t0 = _time.perf_counter()
            b.create_mobile_base_map()
            captured["create_mobile_base_map_s"] = _time.perf_counter() - _t0
        finally:
            vb.VLMapBuilder._init_lseg, vb.get_lseg_feat, vb.save_3d_map, vb.load_3d_map = saved
            np.random.shuffle = orig_shuffle
    rate = map_config["depth_sample_rate"]
    captured["sample_idx"] = [s[::rate].astype(np.int32) for s in sample_orders]
    return captured


# ------------------------------------------------------------------------------------------ multi-floor build
class _FakePointCloud:
    """Stand-in for open3d.geometry.PointCloud: the reference only accumulates points and reads them
    back for np.min / np.max (vlmap_builder_multi_floor.py:104-118)."""

    def __init__(self):
        self.points = np.zeros((0, 3))

    def __iadd__(self, other):
        self.points = np.concatenate([np.asarray(self.points), np.asarray(other.points)], axis=0)
        return self


def _install_open3d_stub():
    _install_stubs()
    o3d = sys.modules["open3d"]
    geometry = types.ModuleType("open3d.geometry")
    geometry.PointCloud = _FakePointCloud
    utility = types.ModuleType("open3d.utility")
    utility.Vector3dVector = lambda a: np.asarray(a)
    o3d.geometry, o3d.utility = geometry, utility


def ref_build_multi_floor(map_config: dict, cam_poses, depths_mm, rgbs, feats, seed: int):
    """Run the reference's VLMapBuilderMultiFloor.create_global_map (vlmap_builder_multi_floor.py:60-199).
    cam_poses[i] (4,4) camera pose in the global frame, depths_mm[i] (H,W) uint16 millimetres (written as
    16-bit PNG, the reference reads them with cv2.IMREAD_UNCHANGED and divides by 1000.0), rgbs / feats as
    ref_build.  Patched: _init_lseg, get_lseg_feat, the save_3d_map method (h5py), open3d's PointCloud.
    Returns the saved arrays + pcd_min / pcd_max + the sample orders of both passes."""
    import cv2

    _install_open3d_stub()
    vb = load("ref_vlmap_builder_multi_floor", "avlmaps/map/vlmap_builder_multi_floor.py")
    cfg = AttrDict(map_config)
    base2cam_tf, base_transform = ref_transforms(map_config)
    D = feats[0].shape[1]
    skip = map_config["skip_frame"]
    used = [i for i in range(len(depths_mm)) if i % skip == 0]
    captured = {}
    frame_counter = {"i": 0}
    sample_orders = []

    def fake_init_lseg(self):
        self.device = "cpu"
        self.clip_feat_dim = D
        return None, None, 480, 520, [0.5] * 3, [0.5] * 3

    def fake_get_lseg_feat(*a, **k):
        i = used[frame_counter["i"]]
        frame_counter["i"] += 1
        return feats[i]

    def fake_save(self, grid_feat, grid_pos, weight, grid_rgb, occupied_ids, mapped_iter_set, max_id):
        captured.update(grid_feat=np.array(grid_feat[:max_id]), grid_pos=np.array(grid_pos[:max_id]),
                        weight=np.array(weight[:max_id]), grid_rgb=np.array(grid_rgb[:max_id]),
                        occupied_ids=np.array(occupied_ids), mapped_iter_list=sorted(mapped_iter_set),
                        pcd_min=np.array(self.pcd_min), pcd_max=np.array(self.pcd_max))

    orig_shuffle = np.random.shuffle

    def recording_shuffle(x):
        orig_shuffle(x)
        sample_orders.append(np.array(x))

    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for sub in ("rgb", "depth", "pose"):
            (td / sub).mkdir()
        rgb_paths, depth_paths, pose_paths = [], [], []
        for i, (d, c, t) in enumerate(zip(depths_mm, rgbs, cam_poses)):
            rp, dp, pp = td / "rgb" / f"{i:06d}.png", td / "depth" / f"{i:06d}.png", td / "pose" / f"{i:06d}.txt"
            cv2.imwrite(str(rp), cv2.cvtColor(c, cv2.COLOR_RGB2BGR))
            assert d.dtype == np.uint16
            cv2.imwrite(str(dp), d)
            np.savetxt(pp, np.asarray(t).reshape(-1))
            rgb_paths.append(rp); depth_paths.append(dp); pose_paths.append(pp)
        cls = vb.VLMapBuilderMultiFloor
        saved = (cls._init_lseg, vb.get_lseg_feat, cls.save_3d_map)
        cls._init_lseg = fake_init_lseg
        vb.get_lseg_feat = fake_get_lseg_feat
        cls.save_3d_map = fake_save
        np.random.shuffle = recording_shuffle
        try:
            np.random.seed(seed)
            b = cls(td, cfg, pose_paths, rgb_paths, depth_paths, base2cam_tf, base_transform)
            b.create_global_map()
        finally:
            cls._init_lseg, vb.get_lseg_feat, cls.save_3d_map = saved
            np.random.shuffle = orig_shuffle
    rate = map_config["depth_sample_rate"]
    orders = [s[::rate].astype(np.int32) for s in sample_orders]
    n = len(used)
    assert len(orders) == 2 * n
    captured["sample_idx_pass1"] = orders[:n]
    captured["sample_idx_pass2"] = orders[n:]
    captured["used_frames"] = used
    return captured


# ------------------------------------------------------------------------------------------ AVLMap heat methods
class _FakeDataloader:
    """Stand-in for VLMapsDataloaderHabitat: the habitat pose's translation IS (row, 0, col) here, so
    to_full_map_pose returns what the test put in (the real one converts metres to cells; scalar pose math)."""

    def from_habitat_tf(self, tf):
        self._tf = np.asarray(tf)

    def to_full_map_pose(self):
        return int(self._tf[0, 3]), int(self._tf[2, 3]), 0.0


def ref_avlmap_heats(occupied_ids, grid_pos, frame_cells, frame_scores, sound_cells, sound_probs, image_cell,
                     area_decay=0.1, sound_decay=0.01, image_decay=0.01, camera_height=1.5, cs=0.05):
    """Run the reference's own AVLMap.index_area_2d / index_area / index_sound_2d / index_sound / index_image
    (avlmaps/map/avlmap.py:78-163) on fake collaborators: avlmap.py is loaded unmodified with `avlmaps.map` and
    `avlmaps.dataloader.habitat_dataloader` stubbed (they pull hloc / librosa / habitat), an AVLMap is created
    without __init__ and given a vlmap (occupied_ids, grid_pos), an area_map (scores + poses), a sound_map
    (probabilities + locations), a visual_map (a localisation result) and the dataloader above."""
    _install_open3d_stub()
    for name in ("avlmaps.map", "avlmaps.dataloader", "avlmaps.dataloader.habitat_dataloader"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    for cls in ("VLMap", "SoundMap", "AreaMap", "VisualMap"):
        setattr(sys.modules["avlmaps.map"], cls, type(cls, (), {}))
    sys.modules["avlmaps.dataloader.habitat_dataloader"].VLMapsDataloaderHabitat = _FakeDataloader
    am = load("ref_avlmap", "avlmaps/map/avlmap.py")

    def tf_of(cell):
        tf = np.eye(4)
        tf[0, 3], tf[2, 3] = cell[0], cell[1]
        return tf

    a = am.AVLMap.__new__(am.AVLMap)
    a.config = AttrDict({"map_config": {"pose_info": {"camera_height": camera_height}}, "params": {"cs": cs}})
    a.vlmap = types.SimpleNamespace(occupied_ids=occupied_ids, grid_pos=grid_pos)
    a.dataloader = _FakeDataloader()
    a.area_map = types.SimpleNamespace(index_map=lambda name, with_init_cat=False: np.array(frame_scores, np.float32),
                                       robot_pose_list=[tf_of(c) for c in frame_cells])
    locs = [[np.array([c[0], 0.0, c[1]]) for c in seg] for seg in sound_cells]
    a.sound_map = types.SimpleNamespace(
        get_distribution_and_locations=lambda name: (np.array(sound_probs, np.float32), locs))
    a.visual_map = types.SimpleNamespace(localize_image=lambda image, query_cam_intrinsic_mat=None: (None, tf_of(image_cell)))
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):   # index_sound_2d prints the map shape
        out = dict(area_2d=a.index_area_2d("kitchen", decay_rate=area_decay), area_3d=a.index_area("kitchen", decay_rate=area_decay),
                   sound_2d=a.index_sound_2d("door", decay_rate=sound_decay), sound_3d=a.index_sound("door", decay_rate=sound_decay),
                   image_3d=a.index_image(None, decay_rate=image_decay))
    return out


# ------------------------------------------------------------------------------------------ template scoring, dynamic obstacles
def _patched_text_feats(encoder):
    """get_text_feats (clip_utils.py:133-149) with the CLIP forward replaced: encode, then the reference's own
    normalisation line (:145) -- rows divided by their L2 norm."""
    def f(in_text, clip_model, clip_feat_dim, batch_size=64):
        feats = np.asarray(encoder(in_text), np.float32)
        return feats / np.linalg.norm(feats, axis=-1, keepdims=True)
    return f


def ref_get_lseg_score_templates(feat, landmarks, encoder, avg_mode):
    """The reference's get_lseg_score with use_multiple_templates=True (clip_utils.py:216-234): 63 prompts per
    landmark (+ "other"), features (avg_mode 0) or scores (avg_mode 1) averaged over the templates."""
    cu = load("ref_clip_utils", "avlmaps/utils/clip_utils.py")
    saved = cu.get_text_feats
    cu.get_text_feats = _patched_text_feats(encoder)
    try:
        return cu.get_lseg_score(None, list(landmarks), feat, feat.shape[-1], use_multiple_templates=True, avg_mode=avg_mode)
    finally:
        cu.get_text_feats = saved


def ref_dynamic_obstacles(encoder, obstacles_cropped, potential, obstacle, grid_feat, grid_pos, rmin, cmin):
    """The reference's get_dynamic_obstacles_map_3d (index_utils.py:138-184) with its own get_lseg_score duplicate
    (:64-108); `openai` is stubbed (only find_similar_category_id's fallback uses it)."""
    if "openai" not in sys.modules:
        sys.modules["openai"] = types.ModuleType("openai")
    iu = load("ref_index_utils", "avlmaps/utils/index_utils.py")
    saved = iu.get_text_feats
    iu.get_text_feats = _patched_text_feats(encoder)
    import contextlib
    import io

    try:
        with contextlib.redirect_stdout(io.StringIO()):
            return iu.get_dynamic_obstacles_map_3d(None, obstacles_cropped, list(potential), list(obstacle), grid_feat,
                                                   grid_pos, rmin, cmin, grid_feat.shape[-1])
    finally:
        iu.get_text_feats = saved
