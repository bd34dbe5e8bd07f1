#!/usr/bin/env python
"""The two reference applications end to end on a synthetic scene (needs a B200):

    application/create_map.py  ->  AVLMap(config).create_map(scene_dir)
    application/index_map.py   ->  AVLMap(config, data_dir=scene_dir); load_map; index_object(name, decay_rate=0.01)

with `avlmaps_b200.map.AVLMap` in place of `avlmaps.map.avlmap.AVLMap`.  The encoders (LSeg for pixels, CLIP for
text) are not part of this engine: random-feature stand-ins are injected where the reference loads checkpoints.
"""
import argparse
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import cv2

    import synth
    from avlmaps_b200.map import AVLMap

    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--rate", type=int, default=1, help="depth_sample_rate (the reference's default is 100)")
    args = ap.parse_args()
    h, w, fh, fw = 480, 640, 390, 520
    map_config = synth.map_config(1000, 0.05, 1.5, [320, 0, 320, 0, 320, 240, 0, 0, 1], args.rate)
    config = {"map_config": map_config, "params": {"cs": 0.05, "gs": 1000}}
    rng = np.random.default_rng(0)
    pool = [rng.standard_normal((1, args.dim, fh, fw), dtype=np.float32) * np.float32(14.2857 / np.sqrt(args.dim)) for _ in range(2)]
    with tempfile.TemporaryDirectory() as td:
        scene = Path(td)
        (scene / "rgb").mkdir()
        (scene / "depth").mkdir()
        for i in range(args.frames):
            cv2.imwrite(str(scene / "rgb" / f"{i:06d}.png"), rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
            base = 2.5 + 1.5 * np.sin(np.linspace(0, 3, w))[None, :] + 0.5 * np.cos(np.linspace(0, 2, h))[:, None]
            np.save(scene / "depth" / f"{i:06d}.npy", (base + 0.02 * rng.standard_normal((h, w))).astype(np.float32))
        np.savetxt(scene / "poses.txt", synth.circle_poses(args.frames, radius=2.0))

        calls = {"n": 0}

        def feature_fn(rgb):          # stands in for get_lseg_feat (lseg_utils.py:20-119)
            calls["n"] += 1
            return pool[calls["n"] % 2]

        def text_encoder(texts):      # stands in for clip_model.encode_text
            return np.stack([np.random.default_rng(abs(hash(t)) % 2 ** 32).standard_normal(args.dim) for t in texts]).astype(np.float32)

        avlmap = AVLMap(config, feature_fn=feature_fn)
        np.random.seed(0)
        t0 = time.perf_counter()
        avlmap.create_map(scene)                                   # create_map.py:17
        t_build = time.perf_counter() - t0
        avlmap = AVLMap(config, data_dir=str(scene))
        t0 = time.perf_counter()
        avlmap.load_map(scene)                                     # index_map.py:27
        t_load = time.perf_counter() - t0
        avlmap.vlmap.set_text_encoder(text_encoder, args.dim)      # index_map.py:28 (_init_clip) stand-in
        n = avlmap.vlmap.grid_feat.shape[0]
        avlmap.index_object("chair", decay_rate=0.01)              # warm-up
        t0 = time.perf_counter()
        heat = avlmap.index_object("sofa", decay_rate=0.01)        # index_map.py:38
        t_index = time.perf_counter() - t0
        goal = avlmap.get_max_pos_3d(heat)
        print(f"create_map: {args.frames} frames of {h}x{w} at rate {args.rate}, D = {args.dim}: {t_build:.2f} s "
              f"({args.frames / t_build:.1f} frames/s incl. PNG/npy reading, host feature hand-off and the map file)")
        print(f"load_map:   {n} voxels x {args.dim} in {t_load:.2f} s (file -> host -> HBM, bf16/fp16 copy + norms)")
        print(f"index_object('sofa'): {t_index * 1e3:.2f} ms for the text -> mask -> heat chain over {n} voxels "
              f"({int((heat >= 1).sum())} target voxels, {int((heat > 0).sum())} with heat > 0); goal voxel {goal.tolist()}")


if __name__ == "__main__":
    main()
