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


# ------------------------------------------------------------------------------------------ multi-floor
MF_CASES = ["rate1", "wrap", "skip2"]


def load_mf(name):
    g = np.load(G / f"mf_{name}.npz")
    cfg = synth.multi_floor_config(float(g["cfg_cs"]), g["cfg_calib"], int(g["cfg_rate"]), skip_frame=int(g["cfg_skip"]))
    return g, cfg


def gpu_build_multi_floor(eng, g, cfg, torch_inputs=False, layout=0, slab=None):
    cs = cfg["cell_size"]
    calib = np.array(cfg["cam_calib_mat"], np.float64).reshape(3, 3)
    kinv = np.linalg.inv(calib)
    used = [int(i) for i in g["used_frames"]]
    tfs = {i: g["poses"][i] @ O.HABITAT2CAM_ROT_TF for i in used}
    fb = eng.FrameBounds()
    for j, i in enumerate(used):
        d, s = g["depths"][i], g["sample_idx_pass1"][j]
        if torch_inputs:
            import torch

            d, s = torch.from_numpy(d).cuda(), torch.from_numpy(s).cuda()
        fb.add_frame(d, kinv, tfs[i], sample_idx=s, min_depth=0.1, max_depth=100)
    pcd_min, pcd_max, n_points = fb.get()
    fb.close()
    n_row, n_col, n_height = O.global_grid_size(pcd_min, pcd_max, cs)
    b = eng.DeviceBuilder.global_grid(n_row, n_col, n_height, cs, pcd_min, int(g["d"]))
    if slab is not None:
        b.set_slab(*slab(n_row))
    for j, i in enumerate(used):
        f = g["feats"][i]
        kfeat = O.get_sim_cam_mat(f.shape[2], f.shape[3])
        if layout == 1:
            f = np.ascontiguousarray(f[0].transpose(1, 2, 0))
        d, r, s = g["depths"][i], g["rgbs"][i], g["sample_idx_pass2"][j]
        if torch_inputs:
            import torch

            f, d, s, r = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (f, d, s, r))
        b.add_frame(d, f, kinv, calib, kfeat, tfs[i], rgb=r, sample_idx=s, feat_layout=layout, min_depth=0.1, max_depth=100)
    out = b.export()
    out.update(pcd_min=pcd_min, pcd_max=pcd_max, n_points=n_points, num_oob=b.num_rejected_oob)
    return out, b


@pytest.mark.parametrize("torch_inputs", [False, True])
@pytest.mark.parametrize("name", MF_CASES)
def test_multi_floor_golden_reference_builds(eng, name, torch_inputs):
    """The UNMODIFIED reference's VLMapBuilderMultiFloor.create_global_map: pcd_min / pcd_max bit-exact from the
    device bounds pass, grid_pos (negative where numpy wrapped) / occupied_ids bit-exact, features within 1e-3."""
    g, cfg = load_mf(name)
    out, b = gpu_build_multi_floor(eng, g, cfg, torch_inputs=torch_inputs, layout=int(torch_inputs))
    assert np.array_equal(out["pcd_min"], g["pcd_min"]) and np.array_equal(out["pcd_max"], g["pcd_max"])
    assert out["occupied_ids"].shape == g["occupied_ids"].shape
    assert_build_equal(out, g)
    assert out["num_oob"] == 0
    b.close()


def test_multi_floor_points_the_reference_would_crash_on_are_counted(eng):
    """Shrink the grid by hand: second-pass points above the (fake) bounds in height raise IndexError in the
    reference; the kernel rejects and counts them, exactly like the C oracle."""
    g, cfg = load_mf("rate1")
    cs = cfg["cell_size"]
    calib = np.array(cfg["cam_calib_mat"], np.float64).reshape(3, 3)
    kinv = np.linalg.inv(calib)
    n_row, n_col, n_height = O.global_grid_size(g["pcd_min"], g["pcd_max"], cs)
    n_height //= 2
    ob = O.BuildOracle.global_grid(n_row, n_col, n_height, cs, g["pcd_min"], int(g["d"]), capacity=n_row * n_col * n_height)
    b = eng.DeviceBuilder.global_grid(n_row, n_col, n_height, cs, g["pcd_min"], int(g["d"]))
    for j, i in enumerate(int(x) for x in g["used_frames"]):
        f = g["feats"][i]
        args = (kinv, calib, O.get_sim_cam_mat(f.shape[2], f.shape[3]), g["poses"][i] @ O.HABITAT2CAM_ROT_TF)
        ob.add_frame(g["depths"][i], f, g["rgbs"][i], g["sample_idx_pass2"][j], *args, min_depth=0.1, max_depth=100)
        b.add_frame(g["depths"][i], f, *args, rgb=g["rgbs"][i], sample_idx=g["sample_idx_pass2"][j], min_depth=0.1,
                    max_depth=100)
    assert ob.num_oob > 0 and b.num_rejected_oob == ob.num_oob and b.num_accepted == ob.num_accepted
    assert_build_equal(b.export(), ob.export())
    b.close()
    ob.close()


# ------------------------------------------------------------------------------------------ slab-sharded build
@pytest.mark.parametrize("world", [2, 3])
def test_slab_sharded_build_equals_single_build(eng, world):
    """Rows split into `world` slabs (here: builders side by side on one GPU; across GPUs the key exchange is
    one all-gather, tests/test_sharded_cpu.py).  Local first-touch ids + ranked keys == the single build's ids,
    and each slab's rows equal the single build's rows of those voxels bit for bit."""
    from avlmaps_b200.sharded import slab_bounds

    cfg, poses, depths, rgbs, feats, sidx = random_scene(5, 60, 80, 49, 65, 16, 48, 0.1, 1.6,
                                                         [40, 0, 40, 0, 40, 30, 0, 0, 1], 1, seed=21, radius=0.3)
    full, bf = gpu_build(eng, cfg, poses, depths, rgbs, feats, sidx)
    cs, gs = cfg["cell_size"], cfg["grid_size"]
    vh = int(cfg["pose_info"]["camera_height"] / cs)
    tfs, calib, kinv = scene_mats(cfg, poses)
    shards = []
    for r in range(world):
        b = eng.DeviceBuilder(gs, vh, cs, 16)
        b.set_slab(*slab_bounds(gs, world, r))
        for i, tf in enumerate(tfs):
            b.add_frame(depths[i], feats[i], kinv, calib, O.get_sim_cam_mat(49, 65), tf, rgb=rgbs[i], sample_idx=sidx[i])
        shards.append((b, b.export(), b.export_keys()))
    keys = [k for _, _, k in shards]
    assert all(np.all(np.diff(k.astype(np.int64)) > 0) for k in keys if k.size > 1)  # ascending = local id order
    assert sum(k.size for k in keys) == full["grid_feat"].shape[0]
    assert sum(b.num_accepted for b, _, _ in shards) == bf.num_accepted
    occ = np.full_like(full["occupied_ids"], -1)
    for r, (b, out, k) in enumerate(shards):
        gids = eng.rank_keys(keys, r)
        assert np.array_equal(gids, np.searchsorted(np.sort(np.concatenate(keys)), k))
        assert np.array_equal(out["grid_pos"], full["grid_pos"][gids])
        lo, hi = slab_bounds(gs, world, r)
        assert np.all((out["grid_pos"][:, 0] >= lo) & (out["grid_pos"][:, 0] < hi))
        # same points, same atomics per voxel row -> sums differ only by fp32 atomic ordering
        assert np.allclose(out["grid_feat"], full["grid_feat"][gids], rtol=1e-4, atol=1e-6)
        assert np.allclose(out["weight"], full["weight"][gids], rtol=1e-5)
        m = out["occupied_ids"] >= 0
        occ[m] = gids[out["occupied_ids"][m]]
        b.close()
    assert np.array_equal(occ, full["occupied_ids"])
    bf.close()


def test_set_slab_after_first_frame_is_a_state_error(eng):
    from avlmaps_b200._lib import AvlError

    cfg, poses, depths, rgbs, feats, sidx = random_scene(1, 30, 40, 24, 32, 4, 32, 0.1, 1.6,
                                                         [20, 0, 20, 0, 20, 15, 0, 0, 1], 1, seed=2, radius=0.3)
    out, b = gpu_build(eng, cfg, poses, depths, None, feats, sidx)
    with pytest.raises(AvlError):
        b.set_slab(0, 16)
    b.close()


def test_resume_from_saved_map(eng):
    """avl_builder_import + frames == the reference run on top of a saved map (vlmap_builder.py:212-222)."""
    from avlmaps_b200._lib import AvlError

    g = np.load(G / "build_resume.npz")
    cfg = synth.map_config(int(g["cfg_gs"]), float(g["cfg_cs"]), float(g["cfg_cam_h"]), g["cfg_calib"], int(g["cfg_rate"]))
    depths, rgbs, feats = synth.build_inputs(int(g["n_frames"]), int(g["h"]), int(g["w"]), int(g["fh"]), int(g["fw"]),
                                             int(g["d"]), seed=int(g["seed"]))
    cs, gs = cfg["cell_size"], cfg["grid_size"]
    vh = int(cfg["pose_info"]["camera_height"] / cs)
    tfs, calib, kinv = scene_mats(cfg, g["poses"])
    b = eng.DeviceBuilder(gs, vh, cs, int(g["d"]), capacity=64)  # smaller than the saved map: rows are re-allocated
    b.import_state(g["first_grid_feat"], g["first_grid_pos"], g["first_weight"], g["first_grid_rgb"])
    assert b.num_voxels == g["first_grid_feat"].shape[0]
    for i, tf in enumerate(tfs):
        b.add_frame(depths[i], feats[i], kinv, calib, O.get_sim_cam_mat(int(g["fh"]), int(g["fw"])), tf, rgb=rgbs[i],
                    sample_idx=g["sample_idx"][i])
    out = b.export()
    ref = {k: g[k] for k in ("grid_feat", "grid_pos", "occupied_ids")}
    ref["weight"] = g["weight"].astype(np.float32)
    ref["grid_rgb"] = np.clip(g["grid_rgb"], 0, 255).astype(np.uint8)
    assert_build_equal(out, ref)
    with pytest.raises(AvlError):  # only a fresh builder can adopt a saved map
        b.import_state(g["first_grid_feat"], g["first_grid_pos"], g["first_weight"])
    b.close()


def test_full_size_c4_properties(eng):
    """BASELINE config 4 at full size -- 2000 frames of 480x640 at depth_sample_rate 1 into a 256 x 256 x 32 grid,
    D = 512, device-resident pixel-major features -- through size-independent properties: voxel ids are a
    permutation of 0..V-1 in bijection with the occupied cells, grid_pos inverts occupied_ids, the accepted-point
    count and the ids are deterministic across two builds, and every weight is positive."""
    import torch

    from avlmaps_b200 import _lib as L
    from avlmaps_b200.map import Map, VLMapBuilder
    from avlmaps_b200.utils.mapping_utils import get_sim_cam_mat

    frames, h, w, fh, fw, d, gs, cs, cam_h = 2000, 480, 640, 390, 520, 512, 256, 0.05, 1.6
    cfg = synth.map_config(gs, cs, cam_h, [320, 0, 320, 0, 320, 240, 0, 0, 1], 1)
    poses = synth.circle_poses(frames, radius=2.0)
    host = Map(cfg)
    tfs = VLMapBuilder("", cfg, None, [], [], host.base2cam_tf, host.base_transform)._frame_transforms(poses)
    calib = np.array(cfg["cam_calib_mat"]).reshape(3, 3)
    kinv, kfeat = np.linalg.inv(calib), get_sim_cam_mat(fh, fw)
    gen = torch.Generator(device="cuda").manual_seed(0)
    np.random.seed(7)
    sidx = [torch.from_numpy(VLMapBuilder._sample_order(h * w, 1)).cuda() for _ in range(4)]
    depths = [torch.rand((h, w), device="cuda", generator=gen) * 5.5 + 0.5 for _ in range(4)]
    pool = [torch.randn((fh, fw, d), device="cuda", generator=gen) * (14.2857 / d ** 0.5) for _ in range(4)]
    vh = int(cam_h / cs)
    outs = []
    for rep in range(2):
        b = eng.DeviceBuilder(gs, vh, cs, d, capacity=gs * gs * vh)
        for i in range(frames):
            b.add_frame(depths[i % 4], pool[i % 4], kinv, calib, kfeat, tfs[i], sample_idx=sidx[i % 4], feat_layout=L.FEAT_HWC)
        out = b.export(want_rgb=False, want_feat=False)
        out["accepted"] = b.num_accepted
        outs.append(out)
        b.close()
    a, c = outs
    v = a["grid_pos"].shape[0]
    occ = a["occupied_ids"]
    assert v > 1_000_000 and a["accepted"] == c["accepted"] > 150_000_000
    assert np.array_equal(a["grid_pos"], c["grid_pos"]) and np.array_equal(occ, c["occupied_ids"])
    ids = occ[occ >= 0]
    assert ids.size == v and np.array_equal(np.sort(ids), np.arange(v))            # bijection cells <-> ids
    gp = a["grid_pos"]
    assert np.array_equal(occ[gp[:, 0], gp[:, 1], gp[:, 2]], np.arange(v))            # grid_pos inverts occupied_ids
    assert np.all(a["weight"] > 0) and np.allclose(a["weight"], c["weight"], rtol=1e-4)


@pytest.mark.parametrize("n_frames", [1, 5, 8, 11, 16, 19])
def test_batched_frames_equal_frame_by_frame(eng, n_frames):
    """avl_builder_add_frames: up to 16 frames share one geometry / id-scan / scatter launch triple.  ids, positions,
    accepted-point counts are identical to a loop of add_frame, and both match the C oracle; includes a frame
    without samples in the middle and frames with different sample counts."""
    import torch

    cfg, poses, depths, rgbs, feats, sidx = random_scene(n_frames, 60, 80, 49, 65, 16, 48, 0.1, 1.6,
                                                         [40, 0, 40, 0, 40, 30, 0, 0, 1], 2, seed=31, radius=0.3)
    if n_frames >= 5:
        sidx[2] = sidx[2][:0]            # an empty sample list still consumes a frame number
        sidx[3] = sidx[3][::3].copy()    # a shorter one
    ref = O.build_map(cfg, poses, depths, rgbs, feats, sidx, capacity=48 * 48 * 16)
    cs, gs = cfg["cell_size"], cfg["grid_size"]
    vh = int(cfg["pose_info"]["camera_height"] / cs)
    tfs, calib, kinv = scene_mats(cfg, poses)
    frames = []
    for i, tf in enumerate(tfs):
        hwc = np.ascontiguousarray(feats[i][0].transpose(1, 2, 0))
        frames.append(dict(depth=torch.from_numpy(depths[i]).cuda(), feat=torch.from_numpy(hwc).cuda(), kinv=kinv, k=calib,
                           kfeat=O.get_sim_cam_mat(49, 65), tf=tf, rgb=torch.from_numpy(rgbs[i]).cuda(),
                           sample_idx=torch.from_numpy(sidx[i]).cuda(), feat_layout=1))
    b = eng.DeviceBuilder(gs, vh, cs, 16)
    b.add_frames(frames)
    out = b.export()
    assert b.n_frames == n_frames and b.num_accepted == ref["num_accepted"]
    assert_build_equal(out, ref)
    b.close()
    b2 = eng.DeviceBuilder(gs, vh, cs, 16)
    prep = b2.prepare_frames(frames[:1] * n_frames)          # marshalled once, pose patched per frame ...
    for i, fr in enumerate(frames):
        if i % 2:
            b2.add_frame(**fr)
        else:                                                # ... only valid here for frames sharing frame 0's buffers
            one = b2.prepare_frames([fr])
            one.set_tf(0, fr["tf"])
            b2.add_prepared(one)
    assert prep.n == n_frames
    out2 = b2.export()
    assert np.array_equal(out["grid_pos"], out2["grid_pos"]) and np.array_equal(out["occupied_ids"], out2["occupied_ids"])
    assert np.allclose(out["grid_feat"], out2["grid_feat"], rtol=1e-4, atol=1e-6)
    b2.close()
    # host arrays / channel-major features take the per-frame path inside the same call
    b3 = eng.DeviceBuilder(gs, vh, cs, 16)
    b3.add_frames([dict(depth=depths[i], feat=feats[i], kinv=kinv, k=calib, kfeat=O.get_sim_cam_mat(49, 65), tf=tfs[i],
                        rgb=rgbs[i], sample_idx=sidx[i]) for i in range(n_frames)])
    assert_build_equal(b3.export(), ref)
    b3.close()


def test_frames_without_valid_depth_and_non_finite_depths(eng):
    """Edge cases of `_backproject_depth`'s mask (`min_depth < z < max_depth`, vlmap_builder.py:266-281): a first
    frame with no valid pixel at all (the map stays empty, export of zero voxels works), then frames with NaN / inf /
    negative / too-far patches, which the mask drops."""
    cfg, poses, depths, rgbs, feats, sidx = random_scene(4, 60, 80, 49, 65, 16, 48, 0.1, 1.6,
                                                         [40, 0, 40, 0, 40, 30, 0, 0, 1], 1, seed=31, radius=0.3)
    depths = [d.copy() for d in depths]
    depths[0][:] = 0.0                                  # below min_depth everywhere
    depths[1][:20] = np.nan
    depths[1][20:30] = np.inf
    depths[2][:, :30] = -1.0
    depths[2][:, 30:40] = 50.0                          # beyond max_depth
    depths[3][::2] = 0.0999                             # just below min_depth (float32(0.1) itself is > 0.1 and passes)
    tfs, calib, kinv = scene_mats(cfg, poses)
    kfeat = O.get_sim_cam_mat(49, 65)
    b = eng.DeviceBuilder(48, 16, 0.1, 16)
    b.add_frame(depths[0], feats[0], kinv, calib, kfeat, tfs[0], rgb=rgbs[0], sample_idx=sidx[0])
    empty = b.export()
    assert b.num_voxels == 0 and b.num_accepted == 0 and empty["grid_feat"].shape == (0, 16)
    assert empty["grid_pos"].shape == (0, 3) and np.all(empty["occupied_ids"] == -1)
    for i in range(1, 4):
        b.add_frame(depths[i], feats[i], kinv, calib, kfeat, tfs[i], rgb=rgbs[i], sample_idx=sidx[i])
    out = b.export()
    ref = O.build_map(cfg, poses, depths, rgbs, feats, sidx, capacity=48 * 48 * 16)
    assert ref["grid_feat"].shape[0] > 100              # the scene is not degenerate
    assert_build_equal(out, ref)
    assert b.num_accepted == ref["num_accepted"]
    b.close()


def test_scratch_reserved_ahead_or_grown_in_the_loop_builds_the_same_map(eng):
    """avl_builder_reserve only moves the scratch allocation out of the frame loop: a builder whose scratch was
    reserved for the largest call, one whose reservation is too small (grows inside add_frames) and one that never
    reserved produce identical maps; a negative size is an argument error."""
    import torch

    n_frames = 6
    cfg, poses, depths, rgbs, feats, sidx = random_scene(n_frames, 60, 80, 49, 65, 16, 48, 0.1, 1.6,
                                                         [40, 0, 40, 0, 40, 30, 0, 0, 1], 2, seed=5, radius=0.3)
    ref = O.build_map(cfg, poses, depths, rgbs, feats, sidx, capacity=48 * 48 * 16)
    cs, gs = cfg["cell_size"], cfg["grid_size"]
    vh = int(cfg["pose_info"]["camera_height"] / cs)
    tfs, calib, kinv = scene_mats(cfg, poses)
    frames = [dict(depth=torch.from_numpy(depths[i]).cuda(),
                   feat=torch.from_numpy(np.ascontiguousarray(feats[i][0].transpose(1, 2, 0))).cuda(), kinv=kinv, k=calib,
                   kfeat=O.get_sim_cam_mat(49, 65), tf=tf, rgb=torch.from_numpy(rgbs[i]).cuda(),
                   sample_idx=torch.from_numpy(sidx[i]).cuda(), feat_layout=1) for i, tf in enumerate(tfs)]
    total = sum(int(s.size) for s in sidx)
    outs = []
    for reserve in (None, 7, total, 4 * total):
        b = eng.DeviceBuilder(gs, vh, cs, 16)
        if reserve is not None:
            b.reserve(reserve)
        b.add_frames(frames[:2])
        b.add_frames(frames[2:])
        assert b.num_accepted == ref["num_accepted"]
        outs.append(b.export())
        if reserve == total:
            with pytest.raises(Exception):
                b.reserve(-1)
        b.close()
    assert_build_equal(outs[0], ref)
    for o in outs[1:]:
        assert np.array_equal(o["grid_pos"], outs[0]["grid_pos"]) and np.array_equal(o["occupied_ids"], outs[0]["occupied_ids"])
        assert np.allclose(o["grid_feat"], outs[0]["grid_feat"], rtol=1e-4, atol=1e-6)
