"""GPU bring-up of the map-build path against the golden vectors / the C oracle."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import synth  # noqa: E402
from avlmaps_b200 import engine  # noqa: E402
from avlmaps_b200 import _lib as L  # noqa: E402
from oracle import avl_oracle as O  # noqa: E402

G = ROOT / "tests" / "golden"


def run_case(name, layout):
    g = np.load(G / f"build_{name}.npz")
    cfg = synth.map_config(int(g["cfg_gs"]), float(g["cfg_cs"]), float(g["cfg_cam_h"]), g["cfg_calib"], int(g["cfg_rate"]))
    if "depths" in g:
        depths, rgbs, feats = list(g["depths"]), list(g["rgbs"]), list(g["feats"])
    else:
        depths, rgbs, feats = synth.build_inputs(int(g["n_frames"]), int(g["h"]), int(g["w"]), int(g["fh"]), int(g["fw"]),
                                                 int(g["d"]), seed=int(g["seed"]), depth_hi=float(g["depth_hi"]))
    cs, gs = cfg["cell_size"], cfg["grid_size"]
    vh = int(cfg["pose_info"]["camera_height"] / cs)
    base2cam, base_tf = O.setup_transforms(cfg["pose_info"])
    tfs = O.frame_transforms(g["poses"], base2cam, base_tf)
    calib = np.array(cfg["cam_calib_mat"]).reshape(3, 3)
    kinv = np.linalg.inv(calib)
    b = engine.DeviceBuilder(gs, vh, cs, int(g["d"]))
    for i, tf in enumerate(tfs):
        f = feats[i]
        kfeat = O.get_sim_cam_mat(f.shape[2], f.shape[3])
        if layout == L.FEAT_HWC:
            f = np.ascontiguousarray(f[0].transpose(1, 2, 0))
        b.add_frame(depths[i], f, kinv, calib, kfeat, tf, rgb=rgbs[i], sample_idx=g["sample_idx"][i], feat_layout=layout)
    out = b.export()
    nacc = b.num_accepted
    r = {"case": name, "layout": layout, "V": int(out["grid_feat"].shape[0]), "V_ref": int(g["grid_feat"].shape[0]),
         "accepted": nacc}
    if r["V"] == r["V_ref"]:
        r["pos_eq"] = bool(np.array_equal(out["grid_pos"], g["grid_pos"]))
        r["occ_eq"] = bool(np.array_equal(out["occupied_ids"], g["occupied_ids"]))
        den = np.maximum(np.abs(g["grid_feat"]), 1e-3 * np.abs(g["grid_feat"]).max())
        r["feat_max_rel"] = float((np.abs(out["grid_feat"] - g["grid_feat"]) / den).max())
        r["weight_max_rel"] = float((np.abs(out["weight"] - g["weight"]) / g["weight"]).max())
        r["rgb_max_diff"] = int(np.abs(out["grid_rgb"].astype(int) - g["grid_rgb"].astype(int)).max())
        r["ok"] = r["pos_eq"] and r["occ_eq"] and r["feat_max_rel"] < 1e-3 and r["weight_max_rel"] < 1e-3
    else:
        r["ok"] = False
    b.close()
    return r


def perf_case(n_frames=24, d=512, rate=1, layout=L.FEAT_HWC, gs=256, cam_h=1.6):
    """BASELINE config 4 geometry (480x640 -> 390x520, 2M-cell grid), inputs resident in HBM."""
    import torch

    h, w, fh, fw = 480, 640, 390, 520
    cs = 0.05
    vh = int(cam_h / cs)
    cfg = synth.map_config(gs, cs, cam_h, [320, 0, 320, 0, 320, 240, 0, 0, 1], rate)
    poses = synth.circle_poses(n_frames, radius=2.0)
    base2cam, base_tf = O.setup_transforms(cfg["pose_info"])
    tfs = O.frame_transforms(poses, base2cam, base_tf)
    calib = np.array(cfg["cam_calib_mat"]).reshape(3, 3)
    kinv = np.linalg.inv(calib)
    kfeat = O.get_sim_cam_mat(fh, fw)
    gen = torch.Generator(device="cuda").manual_seed(0)
    pool = []
    for i in range(4):
        shape = (fh, fw, d) if layout == L.FEAT_HWC else (1, d, fh, fw)
        pool.append(torch.randn(shape, device="cuda", generator=gen) * (14.2857 / d ** 0.5))
    depths = [torch.rand((h, w), device="cuda", generator=gen) * 5.5 + 0.5 for _ in range(4)]
    np.random.seed(7)
    sidx = [torch.from_numpy(O.sample_order(h * w, rate)).cuda() for _ in range(4)]
    b = engine.DeviceBuilder(gs, vh, cs, d, capacity=gs * gs * vh)
    torch.cuda.synchronize()
    times = []
    for rep in range(3):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(n_frames):
            b.add_frame(depths[i % 4], pool[i % 4], kinv, calib, kfeat, tfs[i], sample_idx=sidx[i % 4], feat_layout=layout,
                        stream=torch.cuda.current_stream())
        ev1.record()
        torch.cuda.synchronize()
        times.append(ev0.elapsed_time(ev1) / n_frames)
    nacc = b.num_accepted
    nvox = b.num_voxels
    pacc = nacc / (3 * n_frames)
    bytes_per_frame = h * w * 4 + pacc * (3 * d * 4 + 24)
    b.close()
    return {"case": f"perf_build_layout{layout}_rate{rate}", "ms_per_frame": times, "frames_per_s": 1e3 / min(times),
            "accepted_per_frame": pacc, "voxels": nvox, "alg_GBs": bytes_per_frame / min(times) / 1e6, "ok": True}


def main():
    out = []
    for name in ["small_rate1", "hazard_1080", "full_1080_rate100", "revisit"]:
        for layout in (L.FEAT_CHW, L.FEAT_HWC):
            try:
                r = run_case(name, layout)
            except Exception as e:  # noqa: BLE001
                r = {"case": name, "layout": layout, "ok": False, "error": repr(e)}
            out.append(r)
            print(("PASS " if r.get("ok") else "FAIL ") + json.dumps(r), flush=True)
    for kw in (dict(layout=L.FEAT_HWC, rate=1), dict(layout=L.FEAT_CHW, rate=1), dict(layout=L.FEAT_HWC, rate=100, n_frames=100)):
        try:
            r = perf_case(**kw)
        except Exception as e:  # noqa: BLE001
            r = {"case": f"perf {kw}", "ok": False, "error": repr(e)}
        out.append(r)
        print(("PASS " if r.get("ok") else "FAIL ") + json.dumps(r), flush=True)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "bringup_build.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
