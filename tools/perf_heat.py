#!/usr/bin/env python
"""AVLMap.index_object's heat (get_heatmap_from_mask_3d) at 1M voxels, 1 % targets: timing of both search strategies."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from avlmaps_b200 import _lib as L  # noqa: E402
from avlmaps_b200 import engine  # noqa: E402

L.load()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
n3 = 1_000_000
pos = torch.randint(0, 1000, (n3, 3), device=dev, dtype=torch.int32, generator=g)
pos[:, 2] = pos[:, 2] % 30
mask = (torch.rand(n3, device=dev, generator=g) < 0.01)
stream = torch.cuda.current_stream()
for decay in (0.1, 0.01, 0.004):
    for brute in (False, "bitmap", True):
        os.environ.pop("AVL_HEAT_BRUTE", None)
        os.environ.pop("AVL_HEAT_BITMAP", None)
        if brute is True:
            os.environ["AVL_HEAT_BRUTE"] = "1"
        elif brute == "bitmap":
            os.environ["AVL_HEAT_BITMAP"] = "1"
        tt = []
        for i in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            h = engine.heat_from_mask_3d(pos, mask, 0.05, decay)
            b.record(stream)
            torch.cuda.synchronize()
            tt.append(a.elapsed_time(b))
        print(f"decay {decay} {'brute' if brute is True else ('bitmap walk' if brute else 'scatter')}: {min(tt[1:]):.3f} ms, hot voxels {(h > 0).sum().item()}")
