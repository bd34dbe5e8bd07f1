#!/usr/bin/env python
"""A/B of the per-frame build pipeline on one GPU (BASELINE config 4 geometry, HWC, device-resident).
AVL_BUILD_3PASS=1 selects the original count / scan / assign kernels instead of the look-back kernel."""
import json
import os
import time
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
frames = int(os.environ.get("AVL_FRAMES", "48"))
sc = bench.build_scene(torch, frames)
d = sc["d"]
pool = [torch.randn((sc["fh"], sc["fw"], d), device="cuda", generator=sc["gen"]) * (14.2857 / d ** 0.5) for _ in range(4)]
stream = torch.cuda.current_stream()
batch = int(os.environ.get("AVL_BATCH", "1"))
fr = [dict(depth=sc["depths"][i % 4], feat=pool[i % 4], kinv=sc["kinv"], k=sc["calib"], kfeat=sc["kfeat"], tf=sc["tfs"][i],
           sample_idx=sc["sidx"][i % 4], feat_layout=L.FEAT_HWC) for i in range(frames)]
best = None
for rep in range(4):
    b = engine.DeviceBuilder(sc["gs"], sc["vh"], sc["cs"], d, capacity=sc["gs"] * sc["gs"] * sc["vh"])
    if os.environ.get("AVL_SLAB"):     # one rank's view of a slab-sharded build: rows "lo,hi"
        lo, hi = (int(x) for x in os.environ["AVL_SLAB"].split(","))
        b.set_slab(lo, hi)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    t_host = time.perf_counter()
    if os.environ.get("AVL_PREPARED") == "1":     # descriptors marshalled once (what bench.py and the sharded check do)
        prep = b.prepare_frames(fr)
        b.add_prepared(prep, 0, batch, stream=stream)
        torch.cuda.synchronize()
        e0.record(stream)
        t_host = time.perf_counter()
        for i in range(batch, frames, batch):
            b.add_prepared(prep, i, min(batch, frames - i), stream=stream)
    elif batch > 1:
        for i in range(0, frames, batch):
            b.add_frames(fr[i:i + batch], stream=stream)
    else:
        for f in fr:
            b.add_frame(stream=stream, **f)
    e1.record(stream)
    host_us = (time.perf_counter() - t_host) / frames * 1e6   # time to ENQUEUE a frame (no sync yet)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / frames
    best = ms if best is None else min(best, ms)
    nv = b.num_voxels
    b.close()
print(json.dumps({"batch": batch, "host_enqueue_us_per_frame": host_us, "ms_per_frame": best, "frames_per_s": 1e3 / best, "voxels": nv}))
