"""The oracle against the UNMODIFIED reference executed live (oracle/ref_shim.py), on randomised scenes beyond the
committed golden vectors.  CPU only; skipped where the reference tree does not exist (the GPU box), so nothing that
runs there reads /root/reference.  What is pinned here and not by tests/golden: depth-mask edge cases (NaN / inf /
negative / too-far / no valid pixel), several sampling rates and seeds, the heat function on random masks, and the
scoring function on random shapes."""
from __future__ import annotations

import numpy as np
import pytest

import synth
from oracle import avl_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (it never travels to the GPU box)")


def _assert_same_map(out, ref):
    for k in ("grid_pos", "occupied_ids", "weight", "grid_feat", "grid_rgb"):
        assert np.array_equal(out[k], ref[k]), k


@pytest.mark.parametrize("seed,rate,radius", [(41, 1, 0.3), (42, 3, 0.5), (43, 7, 0.2)])
def test_build_oracle_equals_reference_on_random_scenes(seed, rate, radius):
    cfg = synth.map_config(64, 0.1, 1.6, [40, 0, 40, 0, 40, 30, 0, 0, 1], rate)   # 64^2 rows: no capacity doubling (dtype drift)
    poses = synth.circle_poses(3, radius=radius)
    depths, rgbs, feats = synth.build_inputs(3, 60, 80, 49, 65, 8, seed=seed)
    ref = ref_shim.ref_build(cfg, poses, depths, rgbs, feats, seed=seed)
    out = O.build_map(cfg, poses, depths, rgbs, feats, ref["sample_idx"])
    assert 50 < ref["grid_feat"].shape[0] < 64 * 64
    _assert_same_map(out, ref)
    # the sampling order itself: the reference's global-RNG shuffle (vlmap_builder.py:275-277)
    np.random.seed(seed)
    for s in ref["sample_idx"]:
        assert np.array_equal(O.sample_order(60 * 80, rate), s)


def test_build_oracle_equals_reference_on_depth_mask_edge_cases():
    """`min_depth < z < max_depth` (mapping_utils.py:246-248) is strict and false for NaN; float32(0.1) is above the
    float64 0.1 and passes; a frame without any valid pixel leaves the map empty."""
    cfg = synth.map_config(64, 0.1, 1.6, [40, 0, 40, 0, 40, 30, 0, 0, 1], 1)
    poses = synth.circle_poses(4, radius=0.3)
    depths, rgbs, feats = synth.build_inputs(4, 60, 80, 49, 65, 8, seed=31)
    depths = [d.copy() for d in depths]
    depths[0][:] = 0.0
    depths[1][:20] = np.nan
    depths[1][20:30] = np.inf
    depths[2][:, :30] = -1.0
    depths[2][:, 30:40] = 50.0
    depths[3][::2] = np.float32(0.1)            # > 0.1 as a double: accepted, hundreds of points in a few cells
    depths[3][1::4] = 0.0999
    depths[3][3::4, :40] = 6.0                  # exactly max_depth: rejected
    ref = ref_shim.ref_build(cfg, poses, depths, rgbs, feats, seed=5)
    out = O.build_map(cfg, poses, depths, rgbs, feats, ref["sample_idx"])
    assert 100 < ref["grid_feat"].shape[0] < 64 * 64
    _assert_same_map(out, ref)
    # two frames, neither with a valid pixel (a one-line poses.txt makes np.loadtxt return a 1-D array and the
    # reference itself fails on it, vlmap_builder.py:64-67, so the smallest scene has two frames)
    e_d, e_r, e_f = [depths[0], depths[0]], rgbs[:2], feats[:2]
    only_empty = ref_shim.ref_build(cfg, poses[:2], e_d, e_r, e_f, seed=5)
    assert only_empty["grid_feat"].shape[0] == 0 and np.all(only_empty["occupied_ids"] == -1)
    o2 = O.build_map(cfg, poses[:2], e_d, e_r, e_f, only_empty["sample_idx"])
    assert o2["grid_feat"].shape == (0, 8) and o2["num_accepted"] == 0


@pytest.mark.parametrize("n,d,nq,seed", [(777, 48, 5, 1), (2048, 512, 33, 2), (129, 1024, 2, 3)])
def test_scores_and_mask_equal_reference_on_random_shapes(n, d, nq, seed):
    feat, q = synth.index_inputs(n, d, nq, seed)
    ref = ref_shim.ref_get_lseg_score(feat, q)                       # the reference's float32 `@`
    s = O.scores(feat, q)
    floor = (np.linalg.norm(feat, axis=1)[:, None] * np.linalg.norm(q, axis=1)[None, :]) / np.sqrt(d)
    assert ref.dtype == np.float32 and np.all(np.abs(s - ref) <= 1e-3 * np.maximum(np.abs(ref), floor))
    assert np.array_equal(O.argmax(s), np.argmax(ref, axis=1))
    for c in range(min(nq, 3)):
        assert np.array_equal(O.index_mask(s, c), ref_shim.ref_index_mask(ref, c))
    assert np.array_equal(O.ref_scores_fp32(feat, q), ref)           # the literal restatement: same bits


@pytest.mark.parametrize("seed,frac,decay", [(0, 0.02, 0.1), (1, 0.3, 0.01), (2, 0.002, 0.05)])
def test_heat_equals_reference_on_random_masks(seed, frac, decay):
    rng = np.random.default_rng(seed)
    pos = np.unique(rng.integers(0, 30, (500, 3)).astype(np.int32), axis=0)
    mask = rng.random(pos.shape[0]) < frac
    mask[int(rng.integers(0, pos.shape[0]))] = True                 # the reference's argmin needs one target
    ref = ref_shim.ref_heatmap_from_mask_3d(pos, mask, cell_size=0.05, decay_rate=decay)
    out = O.heatmap_from_mask_3d(pos, mask, 0.05, decay)
    assert np.array_equal(out, ref) and out.dtype == ref.dtype


@pytest.mark.parametrize("seed,rate,skip,n_frames", [(11, 1, 1, 3), (12, 3, 2, 6)])
def test_multi_floor_oracle_equals_reference_on_random_scenes(seed, rate, skip, n_frames):
    """VLMapBuilderMultiFloor.create_global_map (vlmap_builder_multi_floor.py:60-199) run live: both passes, uint16-mm
    depth, np.round cells, numpy negative-index wrap."""
    k10 = [40, 0, 32, 0, 40, 24, 0, 0, 1]
    cfg = synth.multi_floor_config(0.05, k10, rate, skip_frame=skip)
    poses = synth.global_cam_poses(n_frames)
    depths, rgbs, feats = synth.multi_floor_inputs(n_frames, 48, 64, 39, 52, 6, seed=seed)
    try:
        ref = ref_shim.ref_build_multi_floor(cfg, poses, depths, rgbs, feats, seed=seed)
    except IndexError:
        pytest.skip("the reference itself raises IndexError on this random scene (height >= n_height)")
    assert ref["used_frames"] == list(range(0, n_frames, skip))      # the oracle applies skip_frame itself
    out = O.build_map_multi_floor(cfg, poses, depths, rgbs, feats, ref["sample_idx_pass1"], ref["sample_idx_pass2"])
    assert np.array_equal(out["pcd_min"], ref["pcd_min"]) and np.array_equal(out["pcd_max"], ref["pcd_max"])
    _assert_same_map(out, ref)


def test_pose_converter_equals_the_reference_class(tmp_path):
    """avlmaps_b200.dataloader.VLMapsDataloaderHabitat against the reference's class (habitat_dataloader.py:21-148)
    loaded unmodified (hydra / avlmaps.map stubbed: they pull the simulator stack), on random base poses: same
    (row, col, angle), same cropped pose, same habitat transform back, and the reference's own round-trip check."""
    import sys
    import types

    from avlmaps_b200.dataloader import VLMapsDataloaderHabitat
    from avlmaps_b200.map.map import Map

    ref_shim._install_stubs()
    hydra = types.ModuleType("hydra")
    hydra.main = lambda **kw: (lambda fn: fn)
    sys.modules.setdefault("hydra", hydra)
    for name in ("avlmaps.map", "avlmaps.map.map"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["avlmaps.map.map"].Map = object
    ref = ref_shim.load("ref_habitat_dataloader", "avlmaps/dataloader/habitat_dataloader.py")

    cfg = synth.map_config(200, 0.05, 1.5, [40, 0, 40, 0, 40, 30, 0, 0, 1], 1)
    cfg["map_type"] = "vlmap"
    poses = synth.circle_poses(12, radius=1.7)
    np.savetxt(tmp_path / "poses.txt", poses)
    m = Map(cfg, data_dir=str(tmp_path))
    occ = -np.ones((200, 200, 30), np.int32)
    occ[60:140, 70:150, 5] = 1 + np.arange(80 * 80, dtype=np.int32).reshape(80, 80)
    m.occupied_ids = occ
    ours = VLMapsDataloaderHabitat(tmp_path, cfg, m)
    theirs = ref.VLMapsDataloaderHabitat(tmp_path, ref_shim.AttrDict(cfg), m)
    assert (ours.rmin, ours.rmax, ours.cmin, ours.cmax) == (60, 139, 70, 149) == (theirs.rmin, theirs.rmax, theirs.cmin, theirs.cmax)
    rng = np.random.default_rng(0)
    for i in range(12):
        tf = O.cvt_pose_vec2tf(poses[i])
        tf[:3, 3] += rng.uniform(-1.5, 1.5, 3) * [1, 0, 1]
        ours.from_habitat_tf(tf)
        theirs.from_habitat_tf(tf)
        assert ours.to_full_map_pose() == theirs.to_full_map_pose()
        assert ours.to_cropped_map_pose() == theirs.to_cropped_map_pose()
        assert np.array_equal(ours.to_habitat_tf(), theirs.to_habitat_tf())
        assert np.linalg.norm(tf - ours.to_habitat_tf()) < 1            # the reference's own self-check (:170-172)
        cam = np.linalg.inv(ours.base2cam_tf) @ tf
        ours.from_camera_tf(cam)
        theirs.from_camera_tf(cam)
        assert ours.to_full_map_pose() == theirs.to_full_map_pose()
    ours.from_cropped_map_pose(3, 4, 90.0)
    theirs.from_cropped_map_pose(3, 4, 90.0)
    assert ours.to_full_map_pose() == theirs.to_full_map_pose() == [63, 74, 90.0]
    assert np.array_equal(ours.to_habitat_tf(), theirs.to_habitat_tf())


def test_committed_golden_vectors_are_reproducible(tmp_path, monkeypatch):
    """tests/golden/gen_golden.py, run again against the reference tree, reproduces every committed .npz array for
    array: the fixtures are what the committed script makes from the unmodified reference, nothing hand-edited."""
    import importlib.util
    from pathlib import Path

    gdir = Path(__file__).resolve().parent / "golden"
    spec = importlib.util.spec_from_file_location("gen_golden_live", gdir / "gen_golden.py")
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    monkeypatch.setattr(gen, "OUT", tmp_path)
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        gen.main()
    made = sorted(p.name for p in tmp_path.glob("*.npz"))
    committed = sorted(p.name for p in gdir.glob("*.npz"))
    assert made == committed and len(made) == 16
    for name in made:
        with np.load(tmp_path / name) as a, np.load(gdir / name) as b:
            assert sorted(a.files) == sorted(b.files), name
            for k in a.files:
                assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and np.array_equal(a[k], b[k], equal_nan=a[k].dtype.kind == "f"), (name, k)
