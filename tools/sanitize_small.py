"""Small invocation of every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
    compute-sanitizer --tool synccheck python tools/sanitize_small.py

Shapes are tiny (the sanitizer slows kernels by 10-100x); results are still checked against the oracle."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import synth  # noqa: E402
from avlmaps_b200 import _lib, engine  # noqa: E402
from oracle import avl_oracle as O  # noqa: E402

t0 = time.time()
_lib.load()
_lib.require_device()
feat, q = synth.index_inputs(3000, 64, 9, seed=0)
ref = O.scores(feat, q)
m = engine.DeviceMap(feat)
assert np.array_equal(m.argmax(q), O.argmax(ref))
idx, val = m.topk(q, 4)
ri, rv = O.topk(ref, 4)
assert np.array_equal(idx, ri) and np.array_equal(val, rv)
assert np.array_equal(m.scores(q), ref)
print(f"index ok {time.time() - t0:.1f}s", flush=True)
gp = np.random.default_rng(1).integers(0, 20, (600, 3)).astype(np.int32)
mask = np.zeros(600, bool)
mask[::37] = True
h = engine.heat_from_mask_3d(gp, mask, 0.05, 0.1)
assert h.shape == (600,) and h[mask].min() == 1.0
print(f"heat ok {time.time() - t0:.1f}s", flush=True)
cfg = synth.map_config(32, 0.1, 1.6, [20, 0, 20, 0, 20, 15, 0, 0, 1], 1)
poses = synth.circle_poses(2, radius=0.3)
depths, rgbs, feats = synth.build_inputs(2, 30, 40, 25, 33, 16, seed=1)
np.random.seed(3)
sidx = [O.sample_order(30 * 40, 1) for _ in range(2)]
want = O.build_map(cfg, poses, depths, rgbs, feats, sidx, capacity=32 * 32 * 16)
b2c, bt = O.setup_transforms(cfg["pose_info"])
tfs = O.frame_transforms(poses, b2c, bt)
calib = np.array(cfg["cam_calib_mat"]).reshape(3, 3)
b = engine.DeviceBuilder(32, 16, 0.1, 16)
for i in range(2):
    b.add_frame(depths[i], feats[i], np.linalg.inv(calib), calib, O.get_sim_cam_mat(25, 33), tfs[i], rgb=rgbs[i], sample_idx=sidx[i])
got = b.export()
assert np.array_equal(got["grid_pos"], want["grid_pos"]) and np.array_equal(got["occupied_ids"], want["occupied_ids"])
b.close()
m.close()
print(f"build ok {time.time() - t0:.1f}s", flush=True)
