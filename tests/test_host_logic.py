"""CPU: host-side logic of the drop-in classes against the oracle / the reference's conventions."""
import numpy as np
import pytest

import synth
from avlmaps_b200.map.map import Map
from avlmaps_b200.map.vlmap import VLMap, find_similar_category_id
from avlmaps_b200.map.vlmap_builder import VLMapBuilder
from avlmaps_b200.utils import clip_utils, mapping_utils
from oracle import avl_oracle as O


def cfg():
    return synth.map_config(64, 0.05, 1.6, [32, 0, 32, 0, 32, 24, 0, 0, 1], 3)


def test_transforms_match_oracle():
    m = Map(cfg())
    b2c, bt = O.setup_transforms(cfg()["pose_info"])
    assert np.array_equal(m.base2cam_tf, b2c) and np.array_equal(m.base_transform, bt)
    poses = synth.circle_poses(5, 0.7)
    vb = VLMapBuilder("/tmp", cfg(), None, [], [], m.base2cam_tf, m.base_transform)
    got = vb._frame_transforms(poses)
    want = O.frame_transforms(poses, b2c, bt)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)  # same numpy operations in the same order: bit-equal
    assert np.array_equal(mapping_utils.get_sim_cam_mat(390, 520), O.get_sim_cam_mat(390, 520))
    assert np.array_equal(mapping_utils.cvt_pose_vec2tf(poses[2]), O.cvt_pose_vec2tf(poses[2]))


def test_sample_order_uses_global_rng_like_reference():
    np.random.seed(11)
    a = [VLMapBuilder._sample_order(1000, 7) for _ in range(3)]
    np.random.seed(11)
    b = [O.sample_order(1000, 7) for _ in range(3)]
    for x, y in zip(a, b):
        assert x.dtype == np.int32 and np.array_equal(x, y)
    assert len(a[0]) == len(range(0, 1000, 7))


def test_prompt_templates():
    assert len(clip_utils.multiple_templates) == 63
    assert clip_utils.multiple_templates[0] == "There is {} in the scene."
    assert clip_utils.multiple_templates[-1] == "a painting of a {}."
    enc = lambda texts: np.random.default_rng(len(texts)).standard_normal((len(texts), 16))  # noqa: E731
    f = clip_utils.get_text_feats(["a", "b", "c"], enc, 16)
    assert np.allclose(np.linalg.norm(f, axis=1), 1.0, atol=1e-6)  # rows L2-normalised (clip_utils.py:145)
    fm = clip_utils.get_text_feats_multiple_templates(["chair", "other"], enc, 16)
    assert fm.shape == (2, 16) and np.all(np.linalg.norm(fm, axis=1) < 1.0)  # mean of unit rows, not re-normalised
    tf, names, n_tmp = clip_utils.landmark_text_feats(enc, ["chair"], 16, True, 0, True)
    assert names == ["chair", "other"] and tf.shape == (2, 16) and n_tmp == 63
    tf, names, _ = clip_utils.landmark_text_feats(enc, ["chair", "other"], 16, False, 0, True)
    assert names == ["chair", "other"]  # "other" is not appended twice (clip_utils.py:213-215)


def test_save_load_roundtrip(tmp_path):
    p = tmp_path / "vlmaps.h5df"
    gf = np.arange(12, dtype=np.float32).reshape(3, 4)
    gp = np.arange(9, dtype=np.int32).reshape(3, 3)
    w = np.ones(3, np.float32)
    occ = -np.ones((2, 2, 2), np.int32)
    rgb = np.zeros((3, 3), np.uint8)
    assert not mapping_utils.map_file_exists(p)
    mapping_utils.save_3d_map(p, gf, gp, w, occ, [0, 1], rgb)
    assert mapping_utils.map_file_exists(p)
    it, gf2, gp2, w2, occ2, rgb2 = mapping_utils.load_3d_map(p)
    assert it == [0, 1] and np.array_equal(gf, gf2) and np.array_equal(gp, gp2) and np.array_equal(occ, occ2)
    assert w2.dtype == np.float32 and rgb2.dtype == np.uint8


def test_reference_conventions():
    m = VLMap(cfg())
    assert m.grid_feat is None and m.scores_mat is None and m.categories is None
    assert Map(cfg()).create_map("x") is NotImplementedError  # returned, not raised (map.py:70-77)
    assert m.load_map("/nonexistent/dir") is False            # vlmap.py:53-55
    with pytest.raises(Exception, match="Categories are not preloaded"):
        m.index_map("chair", with_init_cat=True)              # vlmap.py:109-112
    with pytest.raises(AttributeError):
        m.index_map("chair", with_init_cat=False)             # needs _init_clip first, like the reference
    assert find_similar_category_id("b", ["a", "b"]) == 1
    assert VLMapBuilder("/tmp", cfg(), None, [], [], None, None).create_camera_map() is NotImplementedError


def test_generate_obstacle_map_quirk():
    m = Map(cfg())
    occ = -np.ones((4, 4, 32), np.int32)
    occ[1, 1, 5] = 0   # voxel id 0 counts as free in the reference (occupied_ids > 0)
    occ[2, 2, 5] = 7
    m.occupied_ids = occ
    obs = m.generate_obstacle_map()
    assert obs[1, 1] and not obs[2, 2]
    assert (m.rmin, m.rmax, m.cmin, m.cmax) == (2, 2, 2, 2)


def test_get_pos_matches_the_reference_helpers():
    """VLMap.get_pos (vlmap.py:158-187): the mask comes from index_map (stubbed here, no GPU); pooling, morphology and
    island extraction must equal the reference's helpers, restated in the test: pool_3d_label_to_2d
    (visualize_utils.py:77-83, a per-voxel loop) and get_segment_islands_pos (index_utils.py:34-62)."""
    import cv2
    from scipy.ndimage import binary_closing, binary_dilation, gaussian_filter

    rng = np.random.default_rng(3)
    gs = 64
    v = VLMap(cfg())
    n = 3000
    pos = np.stack([rng.integers(8, 56, n), rng.integers(10, 50, n), rng.integers(0, 8, n)], 1).astype(np.int32)
    occ = -np.ones((gs, gs, 32), np.int32)
    occ[pos[:, 0], pos[:, 1], pos[:, 2]] = np.arange(n)
    v.grid_pos, v.occupied_ids, v.categories = pos, occ, ["chair"]
    blob = ((pos[:, 0] - 20) ** 2 + (pos[:, 1] - 22) ** 2 < 30) | ((pos[:, 0] - 44) ** 2 + (pos[:, 1] - 40) ** 2 < 16)
    v.index_map = lambda name, with_init_cat=True: blob
    contours, centers, bboxes = v.get_pos("chair")
    # reference arithmetic, loop form
    mask_2d = np.zeros((gs, gs), dtype=bool)
    for i, (row, col, _) in enumerate(pos):
        mask_2d[row, col] = blob[i] or mask_2d[row, col]
    m = Map(cfg())
    m.occupied_ids = occ
    m.generate_obstacle_map()
    crop = mask_2d[m.rmin:m.rmax + 1, m.cmin:m.cmax + 1]
    fg = binary_dilation(gaussian_filter(binary_closing(crop, iterations=3).astype(float), sigma=0.8, truncate=3) > 0.5)
    found, _ = cv2.findContours((fg == 1).astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    assert len(contours) == len(found) == 2
    for got, c, ctr, bb in zip(contours, found, centers, bboxes):
        t = c.reshape((-1, 2))
        want = np.stack([t[:, 1] + m.rmin, t[:, 0] + m.cmin], axis=1)
        assert np.array_equal(got, want)
        assert bb == [want[:, 0].min(), want[:, 0].max(), want[:, 1].min(), want[:, 1].max()]
        assert ctr == [(bb[0] + bb[1]) / 2, (bb[2] + bb[3]) / 2]


def test_frame_marshalling_without_a_device():
    """avl_frame structs as the ctypes layer fills them (no GPU needed): pointers, shapes, the four matrices as
    row-major float64, uint16 depth flag, an empty sample list told apart from 'every pixel', set_tf patching one frame."""
    import ctypes as C
    import types

    from avlmaps_b200 import _lib as L
    from avlmaps_b200 import engine

    depth = np.arange(12, dtype=np.float32).reshape(3, 4)
    feat = np.zeros((1, 5, 2, 3), np.float32)
    k = np.arange(9.0).reshape(3, 3)
    tf = np.arange(16.0).reshape(4, 4).T                    # non-contiguous on purpose
    sidx = np.array([5, 1, 7], np.int32)
    fr, flags, keep = engine._fill_frame(depth, feat, np.linalg.inv(k + np.eye(3)), k, k * 2, tf, None, sidx, L.FEAT_CHW, 0.1, 6.0, dim=5)
    assert (fr.h, fr.w, fr.fh, fr.fw, fr.feat_layout, fr.n_samples) == (3, 4, 2, 3, L.FEAT_CHW, 3)
    assert fr.depth == keep[0].keep.ctypes.data and fr.rgb is None and flags == 0
    assert list(fr.k) == k.ravel().tolist() and list(fr.kfeat) == (k * 2).ravel().tolist()
    assert list(fr.tf) == np.ascontiguousarray(tf).ravel().tolist()
    assert (fr.min_depth, fr.max_depth) == (0.1, 6.0)
    with pytest.raises(ValueError):
        engine._fill_frame(depth, feat, k, k, k, tf, None, sidx, L.FEAT_CHW, 0.1, 6.0, dim=4)   # feature dim mismatch
    # uint16 depth = millimetres -> flag for the kernel's / 1000.0
    _, flags16, _ = engine._fill_frame(depth.astype(np.uint16), feat, k, k, k, tf, None, None, L.FEAT_CHW, 0.1, 100.0, dim=5)
    assert flags16 == L.AVL_DEPTH_U16_MM
    # float16 features are handed over as they are (AVL_FEAT_F16), channel-major or pixel-major; float32 carries no flag
    feat16 = feat.astype(np.float16)
    fr16, flagsf16, keep16 = engine._fill_frame(depth, feat16, k, k, k, tf, None, sidx, L.FEAT_CHW, 0.1, 6.0, dim=5)
    assert flagsf16 == L.AVL_FEAT_F16 and keep16[1].keep.dtype == np.float16 and fr16.feat == feat16.ctypes.data
    assert (fr16.fh, fr16.fw) == (fr.fh, fr.fw) and flags == 0
    hwc16 = np.zeros((3, 4, 5), np.float16)       # pixel-major fp16 rows: the hand-off of an encoder that stays on the GPU
    frh, flagsh, keeph = engine._fill_frame(depth, hwc16, k, k, k, tf, None, sidx, L.FEAT_HWC, 0.1, 6.0, dim=5)
    assert flagsh == L.AVL_FEAT_F16 and (frh.fh, frh.fw, frh.feat_layout) == (3, 4, L.FEAT_HWC) and frh.feat == hwc16.ctypes.data
    # sample_idx=None means every pixel (NULL pointer); an EMPTY list must stay distinguishable (non-NULL, 0 samples)
    fr_all, _, _ = engine._fill_frame(depth, feat, k, k, k, tf, None, None, L.FEAT_CHW, 0.1, 6.0, dim=5)
    fr_none, _, _ = engine._fill_frame(depth, feat, k, k, k, tf, None, sidx[:0], L.FEAT_CHW, 0.1, 6.0, dim=5)
    assert fr_all.sample_idx is None and fr_all.n_samples == 0
    assert fr_none.sample_idx is not None and fr_none.n_samples == 0
    # PreparedFrames: marshalled once, pose patched per frame
    frames = [dict(depth=depth, feat=feat, kinv=k, k=k, kfeat=k, tf=np.eye(4) * (i + 1), sample_idx=sidx) for i in range(3)]
    prep = engine.PreparedFrames(types.SimpleNamespace(dim=5), frames)
    assert prep.n == 3 and [prep.arr[i].tf[0] for i in range(3)] == [1.0, 2.0, 3.0]
    prep.set_tf(1, tf)
    assert list(prep.arr[1].tf) == np.ascontiguousarray(tf).ravel().tolist() and prep.arr[2].tf[0] == 3.0


def test_ctypes_structs_match_the_header(tmp_path):
    """sizeof / offsetof of every struct in include/avlmaps_b200.h, as gcc lays them out, against the ctypes mirrors."""
    import ctypes as C
    import subprocess
    from pathlib import Path

    from avlmaps_b200 import _lib as L

    root = Path(__file__).resolve().parents[1]
    mirrors = {"avl_frame": L.Frame, "avl_grid_spec": L.GridSpec, "avl_global_grid_spec": L.GlobalGridSpec,
               "avl_index_stats": L.IndexStats}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "avlmaps_b200.h"', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {field}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(root / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    for line in out:
        name, size, *offs = line.split()
        cls = mirrors[name]
        assert C.sizeof(cls) == int(size), name
        assert [getattr(cls, f).offset for f, _ in cls._fields_] == [int(o) for o in offs], name


def test_avlmap_builds_its_pose_converter_on_first_use(tmp_path):
    """AVLMap.load_map creates a VLMapsDataloaderHabitat in the reference (avlmap.py:54); here it appears when a
    modality first needs it, unless the caller attached one.  The round trip pose -> cell -> pose stays within a cell."""
    import synth
    from avlmaps_b200.dataloader import VLMapsDataloaderHabitat
    from avlmaps_b200.map import AVLMap
    from avlmaps_b200.utils.mapping_utils import cvt_pose_vec2tf

    map_config = synth.map_config(100, 0.05, 1.5, [40, 0, 40, 0, 40, 30, 0, 0, 1], 1)
    map_config["map_type"] = "vlmap"
    poses = synth.circle_poses(5, radius=1.0)
    np.savetxt(tmp_path / "poses.txt", poses)
    av = AVLMap({"map_config": map_config, "params": {"cs": 0.05}}, data_dir=str(tmp_path))
    occ = -np.ones((100, 100, 30), np.int32)
    occ[20:80, 30:70, 4] = 5
    av.vlmap.occupied_ids = occ
    assert av.dataloader is None
    dl = av._ensure_dataloader()
    assert isinstance(dl, VLMapsDataloaderHabitat) and av._ensure_dataloader() is dl
    assert (dl.rmin, dl.rmax, dl.cmin, dl.cmax) == (20, 79, 30, 69)
    dl.from_habitat_tf(cvt_pose_vec2tf(poses[0]))
    assert dl.to_full_map_pose()[:2] == [50, 50]                 # the first pose is the map origin = the grid centre
    for p in poses:
        tf = cvt_pose_vec2tf(p)
        dl.from_habitat_tf(tf)
        back = dl.to_habitat_tf()
        assert np.linalg.norm(back[:3, 3] - tf[:3, 3]) < 0.05 * 2 ** 0.5 + 1e-9
    sentinel = object()
    av2 = AVLMap({"map_config": map_config, "params": {"cs": 0.05}})
    av2.dataloader = sentinel
    assert av2._ensure_dataloader() is sentinel


def test_category_lookup_takes_an_optional_resolver():
    """find_similar_category_id (index_utils.py:8-32): literal match first; the reference then asks an LLM for the closest
    name -- here a caller-supplied resolver stands in for it, and without one the lookup raises."""
    import pytest

    from avlmaps_b200.map import vlmap

    cats = ["chair", "table", "sofa"]
    assert vlmap.find_similar_category_id("table", cats) == 1
    with pytest.raises(KeyError):
        vlmap.find_similar_category_id("couch", cats)
    assert vlmap.find_similar_category_id("couch", cats, resolver=lambda name, lst: "sofa") == 2
    with pytest.raises(KeyError):
        vlmap.find_similar_category_id("couch", cats, resolver=lambda name, lst: "bench")
