"""GPU parity of the map-build path: CUDA (through the C-ABI) vs the reference's golden vectors and
the C oracle.  Bar: grid_pos / occupied_ids / voxel count bit-exact; grid_feat / weight within 1e-3
relative (fp32 atomics sum in a different order than the reference's sequential running mean; the
observed difference is ~1e-5).  grid_rgb is outside the contract (uint8 truncation at every update
has no order-free form): it is only checked loosely."""
from pathlib import Path

import numpy as np
import pytest

import synth
from oracle import avl_oracle as O

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"
FEAT_RTOL = 1e-3


@pytest.fixture(scope="module")
def eng(lib):
    from avlmaps_b200 import engine

    return engine


def scene_mats(cfg, poses):
    base2cam, base_tf = O.setup_transforms(cfg["pose_info"])
    tfs = O.frame_transforms(poses, base2cam, base_tf)
    calib = np.array(cfg["cam_calib_mat"], np.float64).reshape(3, 3)
    return tfs, calib, np.linalg.inv(calib)


def gpu_build(eng, cfg, poses, depths, rgbs, feats, sample_idx, layout=0, capacity=None, torch_inputs=False):
    cs, gs = cfg["cell_size"], cfg["grid_size"]
    vh = int(cfg["pose_info"]["camera_height"] / cs)
    tfs, calib, kinv = scene_mats(cfg, poses)
    b = eng.DeviceBuilder(gs, vh, cs, feats[0].shape[1], capacity=capacity)
    for i, tf in enumerate(tfs):
        f = feats[i]
        kfeat = O.get_sim_cam_mat(f.shape[2], f.shape[3])
        if layout == 1:
            f = np.ascontiguousarray(f[0].transpose(1, 2, 0))
        d, r, s = depths[i], None if rgbs is None else rgbs[i], sample_idx[i]
        if torch_inputs:
            import torch

            f, d, s = torch.from_numpy(f).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(s).cuda()
            r = None if r is None else torch.from_numpy(r).cuda()
        b.add_frame(d, f, kinv, calib, kfeat, tf, rgb=r, sample_idx=s, feat_layout=layout)
    out = b.export()
    out["num_accepted"] = b.num_accepted
    return out, b


def assert_build_equal(out, ref):
    assert out["grid_feat"].shape == ref["grid_feat"].shape
    assert np.array_equal(out["grid_pos"], ref["grid_pos"])
    assert np.array_equal(out["occupied_ids"], ref["occupied_ids"])
    scale = np.maximum(np.abs(ref["grid_feat"]), 1e-3 * np.abs(ref["grid_feat"]).max())
    assert np.max(np.abs(out["grid_feat"] - ref["grid_feat"]) / scale) < FEAT_RTOL
    assert np.max(np.abs(out["weight"] - ref["weight"]) / ref["weight"]) < FEAT_RTOL
    if "grid_rgb" in ref and ref["grid_rgb"] is not None and out["grid_rgb"].size:
        assert np.abs(out["grid_rgb"].astype(int) - ref["grid_rgb"].astype(int)).max() <= 16


def load_golden(name):
    g = np.load(G / f"build_{name}.npz")
    cfg = synth.map_config(int(g["cfg_gs"]), float(g["cfg_cs"]), float(g["cfg_cam_h"]), g["cfg_calib"], int(g["cfg_rate"]))
    if "depths" in g:
        depths, rgbs, feats = list(g["depths"]), list(g["rgbs"]), list(g["feats"])
    else:
        depths, rgbs, feats = synth.build_inputs(int(g["n_frames"]), int(g["h"]), int(g["w"]), int(g["fh"]), int(g["fw"]),
                                                 int(g["d"]), seed=int(g["seed"]), depth_hi=float(g["depth_hi"]))
    return g, cfg, depths, rgbs, feats


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("name", ["small_rate1", "hazard_1080", "full_1080_rate100", "revisit"])
def test_golden_reference_builds(eng, name, layout):
    """The UNMODIFIED reference's VLMapBuilder output, including the 1080x720 integer-boundary hazards."""
    g, cfg, depths, rgbs, feats = load_golden(name)
    out, b = gpu_build(eng, cfg, g["poses"], depths, rgbs, feats, list(g["sample_idx"]), layout=layout)
    assert_build_equal(out, g)
    b.close()


def random_scene(n_frames, h, w, fh, fw, d, gs, cs, cam_h, calib, rate, seed, radius=0.5):
    cfg = synth.map_config(gs, cs, cam_h, calib, rate)
    poses = synth.circle_poses(n_frames, radius=radius)
    depths, rgbs, feats = synth.build_inputs(n_frames, h, w, fh, fw, d, seed=seed)
    np.random.seed(seed)
    sidx = [O.sample_order(h * w, rate) for _ in range(n_frames)]
    return cfg, poses, depths, rgbs, feats, sidx


def test_random_scene_vs_oracle_and_capacity_growth(eng):
    cfg, poses, depths, rgbs, feats, sidx = random_scene(5, 120, 160, 98, 130, 32, 96, 0.05, 1.6,
                                                         [80, 0, 80, 0, 80, 60, 0, 0, 1], 1, seed=5)
    ref = O.build_map(cfg, poses, depths, rgbs, feats, sidx, capacity=96 * 96 * 32)
    out, b = gpu_build(eng, cfg, poses, depths, rgbs, feats, sidx)
    assert_build_equal(out, ref)
    assert out["num_accepted"] == ref["num_accepted"]
    b.close()
    # tiny initial capacity: rows are re-allocated like _reserve_map_space doubles them (vlmap_builder.py:286-311)
    out2, b2 = gpu_build(eng, cfg, poses, depths, rgbs, feats, sidx, capacity=64)
    assert_build_equal(out2, ref)
    b2.close()
    # device-pointer inputs (torch CUDA tensors) take the same path
    out3, b3 = gpu_build(eng, cfg, poses, depths, rgbs, feats, sidx, layout=1, torch_inputs=True)
    assert_build_equal(out3, ref)
    b3.close()


def test_ids_are_deterministic_and_order_matters(eng):
    cfg, poses, depths, rgbs, feats, sidx = random_scene(4, 60, 80, 49, 65, 8, 48, 0.1, 1.6,
                                                         [40, 0, 40, 0, 40, 30, 0, 0, 1], 2, seed=9, radius=0.3)
    a, b1 = gpu_build(eng, cfg, poses, depths, None, feats, sidx)
    c, b2 = gpu_build(eng, cfg, poses, depths, None, feats, sidx)
    assert np.array_equal(a["grid_pos"], c["grid_pos"]) and np.array_equal(a["occupied_ids"], c["occupied_ids"])
    assert np.allclose(a["grid_feat"], c["grid_feat"], rtol=1e-5, atol=1e-6)
    # fusion is order dependent (first touch is weighted alpha^2): reversing the frames changes ids,
    # and the oracle agrees on the reversed order too
    rev = lambda x: list(reversed(x))  # noqa: E731
    # the map origin is the FIRST pose, so reverse only the frame payloads that share transforms
    ref = O.build_map(cfg, poses, rev(depths), None, rev(feats), rev(sidx), capacity=48 * 48 * 16)
    out, b3 = gpu_build(eng, cfg, poses, rev(depths), None, rev(feats), rev(sidx))
    assert_build_equal(out, ref)
    for x in (b1, b2, b3):
        x.close()


def test_reduced_c4_slice(eng):
    """BASELINE config 4 geometry (480x640 -> 390x520, gs=256, vh=32, D=512), reduced to 3 frames:
    two at depth_sample_rate 100, one at rate 1 (307 200 points)."""
    h, w, fh, fw, d, gs = 480, 640, 390, 520, 512, 256
    cfg = synth.map_config(gs, 0.05, 1.6, [320, 0, 320, 0, 320, 240, 0, 0, 1], 100)
    poses = synth.circle_poses(3, radius=2.0)
    depths, rgbs, feats = synth.build_inputs(3, h, w, fh, fw, d, seed=4, pool=1, depth_lo=0.5, depth_hi=6.0)
    np.random.seed(7)
    sidx = [O.sample_order(h * w, 100), O.sample_order(h * w, 100), O.sample_order(h * w, 1)]
    ref = O.build_map(cfg, poses, depths, None, feats, sidx, capacity=400_000)
    out, b = gpu_build(eng, cfg, poses, depths, None, feats, sidx)
    assert_build_equal(out, ref)
    assert out["num_accepted"] == ref["num_accepted"]
    b.close()


def test_build_then_index_without_leaving_hbm(eng):
    cfg, poses, depths, rgbs, feats, sidx = random_scene(3, 60, 80, 49, 65, 64, 48, 0.1, 1.6,
                                                         [40, 0, 40, 0, 40, 30, 0, 0, 1], 1, seed=13, radius=0.3)
    out, b = gpu_build(eng, cfg, poses, depths, None, feats, sidx)
    m = b.to_map()
    q = synth.index_inputs(1, 64, 5, seed=1)[1]
    idx, val = m.topk(q, 8)
    ri, rv = O.topk(O.scores(out["grid_feat"], q), 8)
    assert np.array_equal(idx, ri) and np.array_equal(val, rv)
    m.close()
    b.close()
