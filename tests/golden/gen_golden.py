"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) through
oracle/ref_shim.py.  Run in the build container only (the reference tree does not travel):

    python tests/golden/gen_golden.py

Every file stores the inputs (or the seeds that regenerate them through tests/synth.py) and the
outputs of the reference's own functions:
  index_*.npz  get_lseg_score (avlmaps/utils/clip_utils.py:196-242) + vlmap.py:123-124 argmax/mask
  sound_*.npz  torch `scale * A @ T.T`, min-max, argmax (avlmaps/map/sound_map.py:108-113,151-152)
  build_*.npz  VLMapBuilder.create_mobile_base_map (avlmaps/map/vlmap_builder.py:54-185)
  heat_*.npz   get_heatmap_from_mask_3d (avlmaps/utils/visualize_utils.py:29-49)
  mf_*.npz     VLMapBuilderMultiFloor.create_global_map (avlmaps/map/vlmap_builder_multi_floor.py:60-199)
  avlmap_heats.npz  AVLMap.index_area(_2d) / index_sound(_2d) / index_image (avlmaps/map/avlmap.py:78-163)
  templates_dynobs.npz  get_lseg_score with the 63 prompt templates, avg_mode 0 / 1 (clip_utils.py:216-234), and
                    get_dynamic_obstacles_map_3d (avlmaps/utils/index_utils.py:138-184)
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import ref_shim  # noqa: E402
import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def gen_index(name, n, d, nq, seed):
    feat, q = synth.index_inputs(n, d, nq, seed)
    scores = ref_shim.ref_get_lseg_score(feat, q)  # reference code path, float32 BLAS
    assert scores.dtype == np.float32 and scores.shape == (n, nq)
    arg = np.argmax(scores, axis=1).astype(np.int32)
    mask0 = ref_shim.ref_index_mask(scores, 0)
    # margin of the reference's own decision, so tests can tell a genuine mismatch from an fp32 tie
    part = np.partition(scores, nq - 2, axis=1) if nq > 1 else scores
    gap = (part[:, -1] - part[:, -2]) if nq > 1 else np.full(n, np.inf, np.float32)
    np.savez_compressed(OUT / f"index_{name}.npz", n=n, d=d, nq=nq, seed=seed, scores=scores, argmax=arg, mask0=mask0,
                        gap=gap.astype(np.float32))
    print(f"index_{name}: n={n} d={d} nq={nq} min gap {gap.min():.3e}")


def gen_sound(name, m, c, seed):
    import torch

    rng = np.random.default_rng(seed)
    a = rng.standard_normal((m, 1024)).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    t = rng.standard_normal((c, 1024)).astype(np.float32)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    logit_scale_at = torch.tensor(np.log(1 / 0.07) + 3.0)  # clamps to 100 like a trained AudioCLIP head
    with torch.no_grad():
        audio_features = torch.from_numpy(a)
        text_features = torch.from_numpy(t)
        scale_audio_text = torch.clamp(logit_scale_at.exp(), min=1.0, max=100.0)  # sound_map.py:108
        logits = scale_audio_text * audio_features @ text_features.T  # sound_map.py:109
    logits = logits.cpu().numpy()
    retrievals = np.argmax(logits, axis=0)  # sound_map.py:113
    cat_id = 1
    prob = logits[:, cat_id]
    prob = (prob - np.min(prob)) / (np.max(prob) - np.min(prob))  # sound_map.py:151-152
    np.savez_compressed(OUT / f"sound_{name}.npz", a=a, t=t, scale=np.float32(scale_audio_text.item()), logits=logits,
                        retrievals=retrievals, cat_id=cat_id, prob=prob.astype(np.float32))
    print(f"sound_{name}: m={m} c={c} scale={scale_audio_text.item()}")


def gen_build(name, n_frames, h, w, fh, fw, d, gs, cs, cam_h, calib, rate, seed, radius=0.6, store_inputs=True,
              depth_hi=6.5):
    cfg = synth.map_config(gs, cs, cam_h, calib, rate)
    poses = synth.circle_poses(n_frames, radius=radius)
    depths, rgbs, feats = synth.build_inputs(n_frames, h, w, fh, fw, d, seed=seed, depth_hi=depth_hi)
    out = ref_shim.ref_build(cfg, poses, depths, rgbs, feats, seed=7 + seed)
    v = out["grid_feat"].shape[0]
    kw = dict(cfg_gs=gs, cfg_cs=cs, cfg_cam_h=cam_h, cfg_calib=np.asarray(calib, np.float64), cfg_rate=rate,
              seed=seed, n_frames=n_frames, h=h, w=w, fh=fh, fw=fw, d=d, radius=radius, depth_hi=depth_hi,
              poses=poses, grid_feat=out["grid_feat"], grid_pos=out["grid_pos"], weight=out["weight"],
              occupied_ids=out["occupied_ids"], grid_rgb=out["grid_rgb"],
              sample_idx=np.stack(out["sample_idx"]).astype(np.int32))
    if store_inputs:
        kw.update(depths=np.stack(depths), rgbs=np.stack(rgbs), feats=np.stack(feats))
    np.savez_compressed(OUT / f"build_{name}.npz", **kw)
    print(f"build_{name}: frames={n_frames} {h}x{w} -> {fh}x{fw} D={d} rate={rate}: {v} voxels, "
          f"weight dtype {out['weight'].dtype}, rgb dtype {out['grid_rgb'].dtype}")


def gen_build_resume(name, n_first, n_frames, h, w, fh, fw, d, gs, cs, cam_h, calib, rate, seed, radius=0.3):
    """Run the reference twice: frames [0, n_first) from scratch, then ALL frames again with the first result
    on disk, so that _init_map takes its reload branch (vlmap_builder.py:212-222)."""
    cfg = synth.map_config(gs, cs, cam_h, calib, rate)
    poses = synth.circle_poses(n_frames, radius=radius)
    depths, rgbs, feats = synth.build_inputs(n_frames, h, w, fh, fw, d, seed=seed)
    first = ref_shim.ref_build(cfg, poses[:n_first], depths[:n_first], rgbs[:n_first], feats[:n_first], seed=seed)
    out = ref_shim.ref_build(cfg, poses, depths, rgbs, feats, seed=seed + 1, resume=first)
    kw = dict(cfg_gs=gs, cfg_cs=cs, cfg_cam_h=cam_h, cfg_calib=np.asarray(calib, np.float64), cfg_rate=rate, seed=seed,
              n_first=n_first, n_frames=n_frames, h=h, w=w, fh=fh, fw=fw, d=d, radius=radius, poses=poses,
              sample_idx=np.stack(out["sample_idx"]).astype(np.int32), mapped_iter_list=np.array(out["mapped_iter_list"]))
    for k in ("grid_feat", "grid_pos", "weight", "occupied_ids", "grid_rgb"):
        kw["first_" + k] = first[k]
        kw[k] = out[k]
    np.savez_compressed(OUT / f"build_{name}.npz", **kw)
    print(f"build_{name}: {first['grid_feat'].shape[0]} voxels reloaded -> {out['grid_feat'].shape[0]}; dtypes after "
          f"the reference's capacity doubling: weight {out['weight'].dtype}, grid_rgb {out['grid_rgb'].dtype}")


def gen_multi_floor(name, n_frames, h, w, fh, fw, d, cs, calib, rate, skip, seed):
    cfg = synth.multi_floor_config(cs, calib, rate, skip_frame=skip)
    poses = synth.global_cam_poses(n_frames)
    depths, rgbs, feats = synth.multi_floor_inputs(n_frames, h, w, fh, fw, d, seed=seed)
    out = ref_shim.ref_build_multi_floor(cfg, poses, depths, rgbs, feats, seed=seed)  # raises where the reference does
    np.savez_compressed(OUT / f"mf_{name}.npz", cfg_cs=cs, cfg_calib=np.asarray(calib, np.float64), cfg_rate=rate,
                        cfg_skip=skip, seed=seed, n_frames=n_frames, h=h, w=w, fh=fh, fw=fw, d=d,
                        poses=np.stack(poses), depths=np.stack(depths), rgbs=np.stack(rgbs), feats=np.stack(feats),
                        grid_feat=out["grid_feat"], grid_pos=out["grid_pos"], weight=out["weight"],
                        occupied_ids=out["occupied_ids"], grid_rgb=out["grid_rgb"], pcd_min=out["pcd_min"],
                        pcd_max=out["pcd_max"], sample_idx_pass1=np.stack(out["sample_idx_pass1"]),
                        sample_idx_pass2=np.stack(out["sample_idx_pass2"]), used_frames=np.array(out["used_frames"]))
    print(f"mf_{name}: frames={n_frames} skip={skip} rate={rate}: {out['grid_feat'].shape[0]} voxels, grid "
          f"{out['occupied_ids'].shape}, min grid_pos {out['grid_pos'].min(0)}")


def gen_heat(name, n, seed):
    rng = np.random.default_rng(seed)
    pos = rng.integers(0, 40, (n, 3)).astype(np.int32)
    mask = rng.uniform(size=n) < 0.05
    heat = ref_shim.ref_heatmap_from_mask_3d(pos, mask, cell_size=0.05, decay_rate=0.1)
    np.savez_compressed(OUT / f"heat_{name}.npz", pos=pos, mask=mask, heat=heat, cell_size=0.05, decay_rate=0.1)
    print(f"heat_{name}: n={n} targets={int(mask.sum())}")


def gen_avlmap_heats(seed=60):
    rng = np.random.default_rng(seed)
    rows, cols, vh, n = 90, 70, 4, 500
    occ = -np.ones((rows, cols, vh), np.int32)
    flat = rng.choice(rows * cols * vh, n, replace=False)
    occ.reshape(-1)[flat] = np.arange(n)
    pos = np.stack(np.unravel_index(flat, occ.shape), 1).astype(np.int32)
    # area: 12 frames, two outside the grid; raw CLIP-like scores (index_area_2d min-max normalises them itself)
    frame_cells = [(int(rng.integers(0, rows)), int(rng.integers(0, cols))) for _ in range(12)]
    frame_cells[3], frame_cells[8] = (-4, 10), (rows + 2, 5)
    frame_scores = rng.standard_normal(12).astype(np.float32)
    # sound: 7 segments with 1-5 locations each (one negative row that wraps like numpy), min-max probabilities
    sound_cells = [[(int(rng.integers(0, rows)), int(rng.integers(0, cols))) for _ in range(int(rng.integers(1, 6)))] for _ in range(7)]
    sound_cells[2][0] = (-3, 5)
    probs = rng.uniform(0, 1, 7).astype(np.float32)
    probs = (probs - probs.min()) / (probs.max() - probs.min())
    image_cell = (41, 33)
    out = ref_shim.ref_avlmap_heats(occ, pos, frame_cells, frame_scores, sound_cells, probs, image_cell)
    flat_cells = np.array([c for seg in sound_cells for c in seg], np.int32)
    seg_len = np.array([len(seg) for seg in sound_cells], np.int32)
    np.savez_compressed(OUT / "avlmap_heats.npz", occupied_ids=occ, grid_pos=pos, frame_cells=np.array(frame_cells, np.int32),
                        frame_scores=frame_scores, sound_cells=flat_cells, sound_seg_len=seg_len, sound_probs=probs,
                        image_cell=np.array(image_cell, np.int32), **out)
    print("avlmap_heats:", {k: (v.shape, str(v.dtype)) for k, v in out.items()})


POTENTIAL_OBSTACLES = ["chair", "wall", "wall above the door", "table", "window", "floor", "stairs", "other"]
OBSTACLES = ["wall", "chair", "table", "window", "stairs", "other"]


def gen_templates_dynobs(seed=70):
    d, n = 64, 4000
    feat, _ = synth.index_inputs(n, d, 1, seed=seed)
    enc = synth.crc_text_encoder(d)
    cats = ["chair", "table", "sofa", "potted plant"]
    s0 = ref_shim.ref_get_lseg_score_templates(feat, cats, enc, avg_mode=0)
    s1 = ref_shim.ref_get_lseg_score_templates(feat, cats, enc, avg_mode=1)
    rng = np.random.default_rng(seed + 1)
    pos = np.stack([rng.integers(5, 45, n), rng.integers(10, 60, n), rng.integers(0, 8, n)], 1).astype(np.int32)
    obstacles_cropped = rng.uniform(size=(40, 50)) > 0.5
    dyn = ref_shim.ref_dynamic_obstacles(enc, obstacles_cropped, POTENTIAL_OBSTACLES, OBSTACLES, feat, pos, 5, 10)
    np.savez_compressed(OUT / "templates_dynobs.npz", n=n, d=d, seed=seed, scores_avg0=s0, scores_avg1=s1, grid_pos=pos,
                        obstacles_cropped=obstacles_cropped, dynamic_obstacles=dyn, rmin=5, cmin=10)
    print("templates_dynobs:", s0.shape, s1.shape, dyn.shape, int(dyn.sum()))


def main():
    assert ref_shim.available(), "reference tree not found"
    # index path: BASELINE config 1 exactly, then batched / other dims
    gen_index("c1_10k_q2", 10_000, 512, 2, seed=0)
    gen_index("4k_q64", 4096, 512, 64, seed=10)
    gen_index("3k_d768_q9", 3000, 768, 9, seed=20)
    gen_index("odd_1001_d100_q3", 1001, 100, 3, seed=30)
    gen_sound("m64_c12", 64, 12, seed=40)
    # build path
    k10 = [32, 0, 32, 0, 32, 24, 0, 0, 1]  # 640x480 sim camera / 10
    gen_build("small_rate1", 3, 48, 64, 39, 52, 16, gs=64, cs=0.05, cam_h=1.6, calib=k10, rate=1, seed=0)
    # the dataset's native 1080x720 -> 520x347 feature map, /10: structural integer-boundary hazards
    k1080 = [54, 0, 54, 0, 54, 36, 0, 0, 1]
    gen_build("hazard_1080", 2, 72, 108, 35, 52, 8, gs=80, cs=0.05, cam_h=1.5, calib=k1080, rate=3, seed=1)
    # full-resolution 1080x720 frame, the default subsampling, tiny D
    kfull = [540, 0, 540, 0, 540, 360, 0, 0, 1]
    gen_build("full_1080_rate100", 2, 720, 1080, 347, 520, 4, gs=200, cs=0.05, cam_h=1.5, calib=kfull, rate=100,
              seed=2, store_inputs=False)
    # larger spread: points leave the grid, several frames revisit the same cells
    gen_build("revisit", 6, 60, 80, 49, 65, 12, gs=48, cs=0.1, cam_h=1.6, calib=[40, 0, 40, 0, 40, 30, 0, 0, 1],
              rate=2, seed=3, radius=0.3)
    gen_build_resume("resume", 2, 4, 60, 80, 49, 65, 6, gs=48, cs=0.1, cam_h=1.6, calib=[40, 0, 40, 0, 40, 30, 0, 0, 1],
                     rate=2, seed=6)
    gen_heat("n600", 600, seed=50)
    gen_avlmap_heats()
    gen_templates_dynobs()
    # multi-floor builder: uint16 mm depth, global-frame grid from a first pass, np.round cells
    gen_multi_floor("rate1", 4, 48, 64, 39, 52, 8, 0.05, k10, rate=1, skip=1, seed=0)
    # second-pass samples fall below pcd_min in all three axes: numpy negative-index wrap-around
    gen_multi_floor("wrap", 5, 48, 64, 39, 52, 6, 0.05, k10, rate=2, skip=1, seed=2)
    gen_multi_floor("skip2", 5, 48, 64, 39, 52, 6, 0.05, k10, rate=4, skip=2, seed=3)


if __name__ == "__main__":
    main()
