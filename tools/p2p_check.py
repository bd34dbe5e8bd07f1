"""Check and time the fused peer-memory exchange (csrc/p2p_exchange.cu) against the NCCL all-gather + merge path.

    python tools/p2p_check.py                                                      # one GPU: exchange with itself
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/p2p_check.py

Every rank holds a slab of a small map; per-slab top-k results are exchanged both ways and must agree bit for bit,
also with ties across slabs and empty (-1) slots.  Prints one JSON line (rank 0) -> gpurun_out/p2p_check_nN.json."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import synth  # noqa: E402
from avlmaps_b200 import engine  # noqa: E402
from avlmaps_b200.sharded import ShardedMap, slab_bounds  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n, d, nq, k = 40_000, 128, 256, 16
    feat, q = synth.index_inputs(n, d, nq, seed=5)
    feat[n // 2 + 7] = feat[11]                       # an exact tie across slabs: the lower global row must win
    lo, hi = slab_bounds(n, world, rank)
    dmap = engine.DeviceMap(feat[lo:hi])
    sm = ShardedMap(dmap, lo)
    qd = torch.from_numpy(q).to(dev)
    ok = True
    timing = {}
    for name, flag in (("nccl", "0"), ("p2p", "1")):
        os.environ["AVL_P2P_EXCHANGE"] = flag
        if world == 1 and flag == "1":
            # ShardedMap only exchanges for world > 1: drive the object directly, one rank talking to itself
            ex = engine.P2PExchange()
            ti, tv = dmap.topk(qd, k)
            res = ex.exchange_merge(ti, tv)
            torch.cuda.synchronize()
            assert ex.timed_out_source() == -1
        else:
            res = sm.topk(qd, k)
        for _ in range(5):
            sm.topk(qd, k)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            sm.topk(qd, k)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        timing[name] = float(t.item())
        if name == "nccl":
            ref = res
        else:
            ok = ok and torch.equal(res[0], ref[0]) and torch.equal(res[1], ref[1])
            if sm._p2p is not None:
                ok = ok and sm._p2p.timed_out_source() == -1
    # ground truth on rank 0: the single-map result
    if rank == 0:
        full = engine.DeviceMap(feat)
        gi, gv = full.topk(q, k)
        ok = ok and np.array_equal(ref[0].cpu().numpy(), gi) and np.array_equal(ref[1].cpu().numpy(), gv)
        line = {"world": world, "ok": bool(ok), "ms_per_step_nccl": timing["nccl"], "ms_per_step_p2p": timing["p2p"],
                "shape": f"{n} x {d} map in {world} slab(s), {nq} queries, top-{k}"}
        print(json.dumps(line))
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        (out / f"p2p_check_n{world}.json").write_text(json.dumps(line) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
