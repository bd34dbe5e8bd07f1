"""GPU: the drop-in classes (VLMapBuilder / VLMap) used the way application/create_map.py and
application/index_map.py use the reference's, on a synthetic scene written to disk."""
from pathlib import Path

import numpy as np
import pytest

import synth
from oracle import avl_oracle as O

pytestmark = pytest.mark.gpu


def write_scene(root: Path, poses, depths, rgbs):
    import cv2

    (root / "rgb").mkdir(parents=True)
    (root / "depth").mkdir()
    for i, (d, c) in enumerate(zip(depths, rgbs)):
        cv2.imwrite(str(root / "rgb" / f"{i:06d}.png"), cv2.cvtColor(c, cv2.COLOR_RGB2BGR))
        np.save(root / "depth" / f"{i:06d}.npy", d)
    np.savetxt(root / "poses.txt", poses)


def fake_encoder(dim):
    def enc(texts):
        out = np.zeros((len(texts), dim), np.float32)
        for i, t in enumerate(texts):
            out[i] = np.random.default_rng(abs(hash(t)) % (2 ** 32)).standard_normal(dim)
        return out

    return enc


def test_create_load_index_like_the_applications(lib, tmp_path):
    from avlmaps_b200.map import VLMap

    n_frames, h, w, fh, fw, d = 4, 60, 80, 49, 65, 32
    cfg = synth.map_config(48, 0.1, 1.6, [40, 0, 40, 0, 40, 30, 0, 0, 1], 2)
    poses = synth.circle_poses(n_frames, radius=0.3)
    depths, rgbs, feats = synth.build_inputs(n_frames, h, w, fh, fw, d, seed=3)
    write_scene(tmp_path, poses, depths, rgbs)
    it = iter(feats)
    vlmap = VLMap(cfg, feature_fn=lambda rgb: next(it))
    np.random.seed(21)
    vlmap.create_map(tmp_path)                    # create_map.py:17 -> vlmap.py:33-48
    assert vlmap.grid_feat is None                # the reference does not populate memory on create either
    assert vlmap.load_map(tmp_path) is True       # index_map.py:27
    np.random.seed(21)
    sidx = [O.sample_order(h * w, 2) for _ in range(n_frames)]
    ref = O.build_map(cfg, poses, depths, rgbs, feats, sidx, capacity=48 * 48 * 16)
    assert np.array_equal(vlmap.grid_pos, ref["grid_pos"]) and np.array_equal(vlmap.occupied_ids, ref["occupied_ids"])
    assert np.allclose(vlmap.grid_feat, ref["grid_feat"], rtol=1e-3, atol=1e-5)
    assert vlmap.mapped_iter_list == list(range(n_frames))

    # ---- index like index_map.py / AVLMap.index_object
    enc = fake_encoder(d)
    vlmap.set_text_encoder(enc, d)
    with pytest.raises(Exception, match="Categories are not preloaded"):
        vlmap.index_map("chair", with_init_cat=True)
    mask = vlmap.index_map("chair", with_init_cat=False)           # vlmap.py:113-124
    from avlmaps_b200.utils.clip_utils import landmark_text_feats

    tf, _, _ = landmark_text_feats(enc, ["chair"], d, True, 0, True)
    want = O.index_mask(O.scores(vlmap.grid_feat, tf), 0)
    assert mask.dtype == bool and np.array_equal(mask, want)
    cats = ["chair", "table", "sofa", "plant"]
    sm = vlmap.init_categories(cats)                                # vlmap.py:92-102
    tfc, _, _ = landmark_text_feats(enc, cats, d, True, 0, True)
    ref_scores = O.scores(vlmap.grid_feat, tfc)
    assert sm.shape == (vlmap.grid_feat.shape[0], 5) and np.array_equal(sm, ref_scores)
    for i, c in enumerate(cats):
        assert np.array_equal(vlmap.index_map(c, with_init_cat=True), O.index_mask(ref_scores, i))
    # get_lseg_score keeps the reference signature (numpy map in, (N, C+1) scores out)
    from avlmaps_b200.utils.clip_utils import get_lseg_score

    s2 = get_lseg_score(enc, cats, vlmap.grid_feat, d, use_multiple_templates=True, add_other=True)
    assert np.array_equal(s2, ref_scores)
    s3 = get_lseg_score(enc, cats, vlmap.grid_feat, d, use_multiple_templates=True, avg_mode=1)
    assert s3.shape == (vlmap.grid_feat.shape[0], 5)


def test_avlmap_index_object(lib):
    """AVLMap.index_object (avlmap.py:67-76): mask -> nearest-target distance decay, incl. the
    `init_categories[1:-1]` quirk, and get_max_pos_3d."""
    from avlmaps_b200.map import AVLMap
    from avlmaps_b200.utils.clip_utils import landmark_text_feats

    d = 32
    feat, _ = synth.index_inputs(3000, d, 1, seed=2)
    pos = np.random.default_rng(0).integers(0, 40, (3000, 3)).astype(np.int32)
    cfg = {"map_config": synth.map_config(48, 0.05, 1.6, [40, 0, 40, 0, 40, 30, 0, 0, 1], 1), "params": {"cs": 0.05}}
    av = AVLMap(cfg)
    av.vlmap.set_map_arrays(feat, grid_pos=pos)
    enc = fake_encoder(d)
    av.vlmap.set_text_encoder(enc, d)
    heat = av.index_object("chair", decay_rate=0.1)
    tf, _, _ = landmark_text_feats(enc, ["chair"], d, True, 0, True)
    mask = O.index_mask(O.scores(feat, tf), 0)
    assert np.array_equal(heat, O.heatmap_from_mask_3d(pos, mask, 0.05, 0.1))
    cats = ["void", "chair", "table", "sofa", "misc"]
    heat2 = av.index_object("table", init_categories=cats, decay_rate=0.1)
    assert av.vlmap.categories == ["chair", "table", "sofa"]
    tf2, _, _ = landmark_text_feats(enc, cats[1:-1], d, True, 0, True)
    mask2 = O.index_mask(O.scores(feat, tf2), 1)
    assert np.array_equal(heat2, O.heatmap_from_mask_3d(pos, mask2, 0.05, 0.1))
    assert np.array_equal(av.get_max_pos_3d(heat2), pos[int(np.argmax(heat2))])


def test_area_and_sound_heat_2d(lib):
    """index_area_2d / index_sound_2d cores (avlmap.py:78-133): one kernel instead of F full-grid EDTs, same bits."""
    from avlmaps_b200.map import AVLMap

    rng = np.random.default_rng(4)
    shape = (90, 70)
    # area: 12 frames, two outside the grid, min-max normalised scores (one is exactly 0, one exactly 1)
    raw = rng.standard_normal(12).astype(np.float32)
    scores = (raw - raw.min()) / (raw.max() - raw.min())
    cells = [None if i in (3, 8) else (int(rng.integers(0, 90)), int(rng.integers(0, 70))) for i in range(12)]
    got = AVLMap.area_heat_2d(shape, cells, scores, decay_rate=0.1)
    want = O.area_heat_2d(shape, cells, scores, decay_rate=0.1)
    assert got.dtype == want.dtype == np.float64 and np.array_equal(got, want)
    # sound: 7 segments with 1-5 locations each (one negative index that wraps like numpy)
    probs = rng.uniform(0, 1, 7).astype(np.float32)
    probs = (probs - probs.min()) / (probs.max() - probs.min())
    segs = [[(int(rng.integers(0, 90)), int(rng.integers(0, 70))) for _ in range(int(rng.integers(1, 6)))] for _ in range(7)]
    segs[2][0] = (-3, 5)
    got = AVLMap.sound_heat_2d(shape, segs, probs, decay_rate=0.01)
    want = O.sound_heat_2d(shape, segs, probs, decay_rate=0.01)
    assert got.dtype == want.dtype == np.float32 and np.array_equal(got, want)
    with pytest.raises(IndexError):
        AVLMap.sound_heat_2d(shape, [[(95, 0)]], np.ones(1, np.float32))
    # 2-D -> 3-D lift through grid_pos == the reference's loop over occupied cells
    occ = -np.ones((90, 70, 4), np.int32)
    flat = rng.choice(90 * 70 * 4, 500, replace=False)
    occ.reshape(-1)[flat] = np.arange(500)
    pos = np.stack(np.unravel_index(flat, occ.shape), 1).astype(np.int32)
    cfg = {"map_config": synth.map_config(90, 0.05, 0.2, [40, 0, 40, 0, 40, 30, 0, 0, 1], 1), "params": {"cs": 0.05}}
    av = AVLMap(cfg)
    av.vlmap.grid_pos, av.vlmap.occupied_ids = pos, occ
    assert np.array_equal(av.lift_heat_2d_to_3d(want), O.lift_heat_2d_to_3d(want, occ, 500))


def test_dynamic_obstacles_map(lib):
    """get_dynamic_obstacles_map_3d (index_utils.py:138-184) on the fused argmax == the reference's numpy steps."""
    from avlmaps_b200.utils.clip_utils import landmark_text_feats
    from avlmaps_b200.utils.index_utils import get_dynamic_obstacles_map_3d

    d, n = 32, 4000
    feat, _ = synth.index_inputs(n, d, 1, seed=8)
    rng = np.random.default_rng(1)
    pos = np.stack([rng.integers(5, 45, n), rng.integers(10, 60, n), rng.integers(0, 8, n)], 1).astype(np.int32)
    rmin, cmin = 5, 10
    obstacles_cropped = rng.uniform(size=(40, 50)) > 0.5
    potential = ["chair", "wall", "wall above the door", "table", "window", "floor", "stairs", "other"]
    obstacle = ["wall", "chair", "table", "window", "stairs", "other"]
    enc = fake_encoder(d)
    got = get_dynamic_obstacles_map_3d(enc, obstacles_cropped, potential, obstacle, feat, pos, rmin, cmin, d)
    tf, _, _ = landmark_text_feats(enc, potential, d, True, 0, True)
    predict = O.argmax(O.scores(feat, tf))
    inds = [i for o in obstacle for i, pn in enumerate(potential) if o == pn]
    pts = np.isin(predict, inds)
    want = np.zeros_like(obstacles_cropped, dtype=bool)
    want[pos[pts, 0] - rmin, pos[pts, 1] - cmin] = 1
    want = np.logical_not(np.logical_and(want, obstacles_cropped == 0))
    assert np.array_equal(got, want)


def test_multi_floor_create_and_load_like_the_reference(lib, tmp_path):
    """VLMapMultiFloor.create_map / load_map on the golden scene written the way the reference's dataset is
    laid out (rgb/*.png, depth/*.png 16-bit mm, pose/*.txt 4x4): `Map.create` dispatch on map_type
    'vlmap_openmap' (map.py:121-129), both passes on the device, the reference's own output as the check."""
    import cv2

    from avlmaps_b200.map import Map, VLMapMultiFloor

    g = np.load(Path(__file__).resolve().parent / "golden" / "mf_wrap.npz")
    cfg = synth.multi_floor_config(float(g["cfg_cs"]), g["cfg_calib"], int(g["cfg_rate"]), skip_frame=int(g["cfg_skip"]))
    for sub in ("rgb", "depth", "pose"):
        (tmp_path / sub).mkdir()
    for i in range(int(g["n_frames"])):
        cv2.imwrite(str(tmp_path / "rgb" / f"{i:06d}.png"), cv2.cvtColor(g["rgbs"][i], cv2.COLOR_RGB2BGR))
        cv2.imwrite(str(tmp_path / "depth" / f"{i:06d}.png"), g["depths"][i])
        np.savetxt(tmp_path / "pose" / f"{i:06d}.txt", g["poses"][i].reshape(-1))
    vlmap = Map.create(cfg)
    assert isinstance(vlmap, VLMapMultiFloor)
    feats = iter(g["feats"][int(i)] for i in g["used_frames"])
    vlmap.feature_fn = lambda rgb: next(feats)
    np.random.seed(int(g["seed"]))  # the reference draws both passes' sample orders from the global RNG
    vlmap.create_map(tmp_path)
    assert vlmap.grid_feat is None
    assert vlmap.load_map(tmp_path) is True
    assert np.array_equal(vlmap.pcd_min, g["pcd_min"]) and np.array_equal(vlmap.pcd_max, g["pcd_max"])
    assert float(vlmap.cs) == float(g["cfg_cs"])
    assert np.array_equal(vlmap.grid_pos, g["grid_pos"]) and np.array_equal(vlmap.occupied_ids, g["occupied_ids"])
    assert np.allclose(vlmap.grid_feat, g["grid_feat"], rtol=1e-3, atol=1e-5)
    assert np.allclose(vlmap.weight, g["weight"], rtol=1e-3)
    assert vlmap.mapped_iter_list == [int(i) for i in g["used_frames"]]
    assert vlmap.map_builder.device_builder.num_rejected_oob == 0
    # the index surface is VLMap's
    enc = fake_encoder(int(g["d"]))
    vlmap.set_text_encoder(enc, int(g["d"]))
    from avlmaps_b200.utils.clip_utils import landmark_text_feats

    tf, _, _ = landmark_text_feats(enc, ["sofa"], int(g["d"]), True, 0, True)
    assert np.array_equal(vlmap.index_map("sofa", with_init_cat=False), O.index_mask(O.scores(vlmap.grid_feat, tf), 0))
    assert vlmap.map_builder.create_mobile_base_map() is NotImplementedError


def test_create_map_twice_resumes_like_the_reference(lib, tmp_path):
    """A second create_map over an existing vlmaps file reloads it and fuses every frame on top
    (vlmap_builder.py:212-222, the loop never skips mapped frames): checked against the C oracle."""
    from avlmaps_b200.map import VLMap

    n_frames, h, w, fh, fw, d = 3, 60, 80, 49, 65, 8
    cfg = synth.map_config(48, 0.1, 1.6, [40, 0, 40, 0, 40, 30, 0, 0, 1], 2)
    poses = synth.circle_poses(n_frames, radius=0.3)
    depths, rgbs, feats = synth.build_inputs(n_frames, h, w, fh, fw, d, seed=8)
    write_scene(tmp_path, poses, depths, rgbs)
    for seed in (5, 6):
        it = iter(feats)
        vlmap = VLMap(cfg, feature_fn=lambda rgb: next(it))
        np.random.seed(seed)
        vlmap.create_map(tmp_path)
    assert vlmap.load_map(tmp_path) is True
    np.random.seed(5)
    s1 = [O.sample_order(h * w, 2) for _ in range(n_frames)]
    first = O.build_map(cfg, poses, depths, rgbs, feats, s1, capacity=48 * 48 * 16)
    np.random.seed(6)
    s2 = [O.sample_order(h * w, 2) for _ in range(n_frames)]
    ref = O.build_map(cfg, poses, depths, rgbs, feats, s2, capacity=48 * 48 * 16, resume=first)
    assert np.array_equal(vlmap.grid_pos, ref["grid_pos"]) and np.array_equal(vlmap.occupied_ids, ref["occupied_ids"])
    assert np.allclose(vlmap.weight, ref["weight"], rtol=1e-3)
    assert np.allclose(vlmap.grid_feat, ref["grid_feat"], rtol=1e-3, atol=1e-5)


def test_area_and_sound_map_similarity_call_sites(lib):
    """AreaMap.init_categories / index_map (area_map.py:99-119) and SoundMap (sound_map.py:102-153): the golden
    logits of the reference's torch expression, its argmax retrieval and its min-max."""
    from avlmaps_b200.map import AreaMap, SoundMap

    g = np.load(Path(__file__).resolve().parent / "golden" / "sound_m64_c12.npz")
    a, t = g["a"], g["t"]
    db = {i: {"audio_features": a[i], "locations": [np.array([i, 0.0, -i])]} for i in range(a.shape[0])}
    cats = [f"sound{j}" for j in range(t.shape[0])]
    sm = SoundMap(cats, text_encoder=lambda texts: t, logit_scale_at=np.log(1 / 0.07) + 3.0, audio_database=db,
                  audio_encoder=lambda path, sr: a[17])
    assert sm.scale_audio_text == float(g["scale"]) == 100.0
    cat = int(g["cat_id"])
    prob, locs = sm.get_distribution_and_locations(cats[cat])
    assert np.allclose(prob, g["prob"], atol=2e-6) and len(locs) == a.shape[0]
    assert np.array_equal(sm.get_pos(cats[cat])[0], db[int(g["retrievals"][cat])]["locations"][0])
    assert np.array_equal(sm.get_pos_with_audio(__file__, 44100)[0], db[17]["locations"][0])   # a @ a[17] peaks at 17
    assert sm.get_pos_with_audio("/nonexistent.wav", 44100) == ([], [])
    # AreaMap: true cosine of unit rows; text rows are normalised by get_text_feats like the reference
    rng = np.random.default_rng(3)
    frames = rng.standard_normal((40, 768)).astype(np.float32)
    frames /= np.linalg.norm(frames, axis=1, keepdims=True)
    enc = fake_encoder(768)
    am = AreaMap(text_encoder=enc)
    am.set_sparse_map(frames, [np.eye(4)] * 40)
    with pytest.raises(Exception, match="Categories are not preloaded"):
        am.index_map("kitchen")
    tf = enc(["kitchen", "bedroom"])
    tf /= np.linalg.norm(tf, axis=1, keepdims=True)
    sc = am.init_categories(["kitchen", "bedroom"])
    assert np.array_equal(sc, O.scores(frames, tf))
    assert np.allclose(sc, frames @ tf.T, atol=1e-6)
    assert np.array_equal(am.index_map("bedroom"), sc[:, 1])
    assert np.array_equal(am.index_map("kitchen", with_init_cat=False), O.scores(frames, tf[:1]).flatten())


def test_image_heat_planar(lib):
    """AVLMap.index_image's numeric lines (avlmap.py:156-162) as numpy executes them, bit for bit."""
    from avlmaps_b200.map import AVLMap

    rng = np.random.default_rng(9)
    pos = rng.integers(0, 1000, (5000, 3)).astype(np.int32)
    cfg = {"map_config": synth.map_config(1000, 0.05, 1.5, [40, 0, 40, 0, 40, 30, 0, 0, 1], 1), "params": {"cs": 0.05}}
    av = AVLMap(cfg)
    av.vlmap.grid_pos = pos
    row, col, height, decay = 417, 633, 1.5 / 0.05, 0.01
    p = np.array([row, col, height])
    sim_mat = np.zeros((pos.shape[0], 1))
    sim_mat[:, 0] = np.clip(1.0 - decay * np.linalg.norm((pos - p)[:, :2], axis=1), 0, 1)
    want = np.max(sim_mat, axis=1).flatten()
    got = av.image_heat(row, col, decay)
    assert got.dtype == np.float64 and np.array_equal(got, want)


def test_avlmap_area_sound_image_against_the_reference_methods(lib):
    """AVLMap.index_area(_2d) / index_sound(_2d) / index_image with the collaborators the reference uses (an area map with
    frame scores and poses, a sound map with probabilities and locations, a localiser, the map-pose converter) against the
    outputs of the reference's own methods on the same fakes (tests/golden/avlmap_heats.npz): same dtypes, same bits."""
    import types

    from avlmaps_b200.map import AVLMap

    g = np.load(Path(__file__).resolve().parent / "golden" / "avlmap_heats.npz")
    rows, cols, vh = g["occupied_ids"].shape

    class Loader:  # same convention as oracle/ref_shim._FakeDataloader: translation = (row, 0, col)
        def from_habitat_tf(self, tf):
            self.tf = np.asarray(tf)

        def to_full_map_pose(self):
            return int(self.tf[0, 3]), int(self.tf[2, 3]), 0.0

    def tf_of(cell):
        tf = np.eye(4)
        tf[0, 3], tf[2, 3] = cell[0], cell[1]
        return tf

    segs, o = [], 0
    for n in g["sound_seg_len"]:
        segs.append([np.array([r, 0.0, c]) for r, c in g["sound_cells"][o:o + n]])
        o += n
    cfg = {"map_config": synth.map_config(rows, 0.05, 1.5, [40, 0, 40, 0, 40, 30, 0, 0, 1], 1), "params": {"cs": 0.05}}
    av = AVLMap(cfg)
    av.vlmap.grid_pos, av.vlmap.occupied_ids = g["grid_pos"], g["occupied_ids"]
    av.dataloader = Loader()
    av.area_map = types.SimpleNamespace(index_map=lambda name, with_init_cat=False: g["frame_scores"].copy(),
                                        robot_pose_list=[tf_of(c) for c in g["frame_cells"]])
    av.sound_map = types.SimpleNamespace(get_distribution_and_locations=lambda name: (g["sound_probs"].copy(), segs))
    av.visual_map = types.SimpleNamespace(localize_image=lambda image, query_cam_intrinsic_mat=None: (None, tf_of(g["image_cell"])))
    for got, key in ((av.index_area_2d("kitchen", decay_rate=0.1), "area_2d"), (av.index_area("kitchen", decay_rate=0.1), "area_3d"),
                     (av.index_sound_2d("door", decay_rate=0.01), "sound_2d"), (av.index_sound("door", decay_rate=0.01), "sound_3d"),
                     (av.index_image(None, decay_rate=0.01), "image_3d")):
        assert got.dtype == g[key].dtype and np.array_equal(got, g[key]), key


def test_templates_and_dynamic_obstacles_against_the_reference_functions(lib):
    """get_lseg_score with the 63 templates (both avg modes) and get_dynamic_obstacles_map_3d against the outputs of
    the reference's own functions (tests/golden/templates_dynobs.npz, produced through oracle/ref_shim.py)."""
    from avlmaps_b200.utils.clip_utils import get_lseg_score
    from avlmaps_b200.utils.index_utils import get_dynamic_obstacles_map_3d

    g = np.load(Path(__file__).resolve().parent / "golden" / "templates_dynobs.npz")
    d = int(g["d"])
    feat, _ = synth.index_inputs(int(g["n"]), d, 1, seed=int(g["seed"]))
    enc = synth.crc_text_encoder(d)
    cats = ["chair", "table", "sofa", "potted plant"]
    for mode, key in ((0, "scores_avg0"), (1, "scores_avg1")):
        sc = get_lseg_score(enc, cats, feat, d, use_multiple_templates=True, avg_mode=mode)
        ref = g[key]
        assert sc.dtype == np.float32 and sc.shape == ref.shape
        assert np.max(np.abs(sc - ref)) <= 2e-6 * np.abs(ref).max()       # fp64-accumulated vs the reference's sgemm
    potential = ["chair", "wall", "wall above the door", "table", "window", "floor", "stairs", "other"]
    obstacle = ["wall", "chair", "table", "window", "stairs", "other"]
    got = get_dynamic_obstacles_map_3d(enc, g["obstacles_cropped"], potential, obstacle, feat, g["grid_pos"],
                                       int(g["rmin"]), int(g["cmin"]), d)
    assert got.dtype == bool and np.array_equal(got, g["dynamic_obstacles"])


def test_load_map_from_a_memory_mapped_file_and_slab_wise(lib, tmp_path):
    """grid_feat goes from the page cache to the device without a host copy (`mmap_load`), and
    ShardedMap.from_file uploads only this rank's row slab: same answers as the copying load."""
    from avlmaps_b200.map import VLMap
    from avlmaps_b200.sharded import ShardedMap
    from avlmaps_b200.utils import mapping_utils

    feat, q = synth.index_inputs(6000, 64, 9, seed=12)
    gp = np.random.default_rng(0).integers(0, 48, (6000, 3)).astype(np.int32)
    (tmp_path / "vlmap").mkdir()
    mapping_utils.save_3d_map(tmp_path / "vlmap" / "vlmaps.h5df", feat, gp, np.ones(6000, np.float32),
                              -np.ones((48, 48, 16), np.int32), [0], np.zeros((6000, 3), np.uint8))
    cfg = synth.map_config(48, 0.1, 1.6, [40, 0, 40, 0, 40, 30, 0, 0, 1], 1)
    a, b = VLMap(cfg), VLMap(cfg)
    b.mmap_load = True
    assert a.load_map(tmp_path) and b.load_map(tmp_path)
    assert isinstance(b.grid_feat, np.memmap) and not isinstance(a.grid_feat, np.memmap)
    assert np.array_equal(a.grid_feat, b.grid_feat)
    want = O.argmax(O.scores(feat, q))
    assert np.array_equal(a.device_map.argmax(q), want) and np.array_equal(b.device_map.argmax(q), want)
    sm = ShardedMap.from_file(tmp_path / "vlmap" / "vlmaps.h5df")           # world of one: the slab is the whole map
    idx, val = sm.topk(q, 8)
    ri, rv = O.topk(O.scores(feat, q), 8)
    assert (sm.row_lo, sm.row_hi, sm.n_total) == (0, 6000, 6000) and np.array_equal(sm.grid_pos, gp)
    assert np.array_equal(idx, ri) and np.array_equal(val, rv)
