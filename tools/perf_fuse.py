#!/usr/bin/env python
"""BASELINE config 3 (1M x 512 + 1M x 1024, 32 + 32 queries, scale 100, product, top-16): one timed call,
for `ncu --metrics gpu__time_duration.sum` launch lists."""
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
dev = torch.device("cuda", 0)
n3 = 1_000_000
g = torch.Generator(device=dev).manual_seed(7)
mv = engine.DeviceMap(bench.make_shard(torch, n3, 512, 11, dev))
fa = torch.randn((n3, 1024), device=dev, generator=g)
fa = fa / fa.norm(dim=1, keepdim=True)
ma = engine.DeviceMap(fa)
del fa
qv = torch.randn((32, 512), device=dev, generator=g)
qv = (qv / qv.norm(dim=1, keepdim=True)).contiguous()
qa = torch.randn((32, 1024), device=dev, generator=g)
qa = (qa / qa.norm(dim=1, keepdim=True)).contiguous()
sa = torch.full((32,), 100.0, device=dev)
stream = torch.cuda.current_stream()
tt = []
for i in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    engine.fuse_topk(mv, qv, ma, qa, 16, scale_b=sa, combine=L.FUSE_PRODUCT)
    b.record(stream)
    torch.cuda.synchronize()
    tt.append(a.elapsed_time(b))
print("fuse ms per call:", tt)
