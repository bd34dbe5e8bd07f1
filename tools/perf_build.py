#!/usr/bin/env python
"""A/B of the per-frame build pipeline on one GPU (BASELINE config 4 geometry, HWC, device-resident).
AVL_BUILD_3PASS=1 selects the original count / scan / assign kernels instead of the look-back kernel."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from avlmaps_b200 import _lib as L  # noqa: E402
from avlmaps_b200 import engine  # noqa: E402

L.load()
frames = 48
sc = bench.build_scene(torch, frames)
d = sc["d"]
pool = [torch.randn((sc["fh"], sc["fw"], d), device="cuda", generator=sc["gen"]) * (14.2857 / d ** 0.5) for _ in range(4)]
stream = torch.cuda.current_stream()
best = None
for rep in range(4):
    b = engine.DeviceBuilder(sc["gs"], sc["vh"], sc["cs"], d, capacity=sc["gs"] * sc["gs"] * sc["vh"])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for i in range(frames):
        b.add_frame(sc["depths"][i % 4], pool[i % 4], sc["kinv"], sc["calib"], sc["kfeat"], sc["tfs"][i],
                    sample_idx=sc["sidx"][i % 4], feat_layout=L.FEAT_HWC, stream=stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / frames
    best = ms if best is None else min(best, ms)
    nv = b.num_voxels
    b.close()
print(json.dumps({"ms_per_frame": best, "frames_per_s": 1e3 / best, "voxels": nv}))
