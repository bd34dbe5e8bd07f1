"""BASELINE config 5 on real GPUs: a 16 777 216 x 512 map in 8 shards of 2 097 152 rows (shard s seeded 1000 + s), 256
queries, top-16, slab-sharded over the ranks of one box.  Parity of the merged result, for BOTH exchanges (the fused
peer-memory kernel and NCCL all-gather + merge), against

  (a) the exact columns: every rank scores its slab EXACTLY (avl_sim_dense, fp64-accumulated canonical scores) for 3
      queries, takes its exact top-16 with the (score desc, row asc) rule, the ranks' lists are gathered and merged on the
      host -- ids and score bits must equal the engine's merged result for those queries;
  (b) the CPU oracle on sampled rows: the feature rows the engine returned for 8 queries (and the rows next to them) are
      re-scored by oracle.avl_oracle on the host -- the returned scores must equal the oracle's bit for bit, and be
      ordered (score desc, row asc).

The map does not depend on the number of ranks (rank r holds shards [r * 8 / world, (r + 1) * 8 / world)), so the
digest printed at N = 1, 2, 4, 8 must be the same line.  Also times the step (device-resident queries, asynchronous
calls, max over ranks): the strong-scaling line of the fixed 16 M map.

    python tools/sharded_index_check.py                                                        # one GPU, whole map
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_index_check.py
"""
import hashlib
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from avlmaps_b200 import engine  # noqa: E402
from avlmaps_b200.sharded import ShardedMap, merge_topk  # noqa: E402

N_SHARDS, SHARD_ROWS, DIM, NQ, K = 8, 2_097_152, 512, 256, 16


def make_shard(seed, device, rows=SHARD_ROWS):
    g = torch.Generator(device=device).manual_seed(seed)
    feat = torch.empty((rows, DIM), dtype=torch.float32, device=device)
    step = 1 << 19
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        feat[r0:r1] = torch.randn((r1 - r0, DIM), device=device, generator=g)
        feat[r0:r1] *= 14.2857 * (0.05 + 0.95 * torch.rand((r1 - r0, 1), device=device, generator=g)) / (DIM ** 0.5)
    return feat


def exact_local_topk(scores_col, row_offset, k):
    """scores_col (n,) float32 CUDA, exact -> (ids int64 (k,), vals float32 (k,)) by (score desc, row asc)."""
    v, i = torch.topk(scores_col, min(4 * k, scores_col.numel()))
    v, i = v.cpu().numpy(), i.cpu().numpy().astype(np.int64)
    order = np.lexsort((i, -v.astype(np.float64)))[:k]
    return i[order] + row_offset, v[order]


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    shard_rows = int(os.environ.get("AVL_CHECK_SHARD_ROWS", SHARD_ROWS))   # smaller map for a quick run
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert N_SHARDS % world == 0, "world must divide 8"
    per = N_SHARDS // world
    feat = torch.cat([make_shard(1000 + s, dev, shard_rows) for s in range(rank * per, (rank + 1) * per)])
    lo = rank * per * shard_rows
    dmap = engine.DeviceMap(feat)
    sm = ShardedMap(dmap, lo)
    sm.check_exchange = True
    g = torch.Generator(device=dev).manual_seed(7)
    q = torch.randn((NQ, DIM), device=dev, generator=g)
    q = (q / q.norm(dim=1, keepdim=True)).contiguous()
    ok, notes, results, timing = True, [], {}, {}
    for name, flag in (("p2p", "1"), ("nccl", "0")):
        if world == 1 and name == "nccl":
            continue
        os.environ["AVL_P2P_EXCHANGE"] = flag
        mi, mv = sm.topk(q, K)
        torch.cuda.synchronize()
        results[name] = (mi.cpu().numpy().copy(), mv.cpu().numpy().copy())
        # ---- timing: 30 asynchronous steps, max over ranks
        for _ in range(5):
            sm.topk(q, K)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 30
        e0.record()
        pend = None
        for _ in range(steps):
            pend = sm.topk_async(q, K) if (name == "p2p" and world > 1) else sm.topk(q, K)
        if hasattr(pend, "result"):
            pend.result()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        timing[name] = float(t.item())
    ref_i, ref_v = results["p2p"]
    if "nccl" in results:
        same = np.array_equal(results["nccl"][0], ref_i) and np.array_equal(results["nccl"][1], ref_v)
        ok &= same
        notes.append(f"p2p == nccl: {same}")
    # ---- (a) exact columns for 3 queries over all shards
    qa = [0, 101, 255]
    sc = dmap.scores(q[qa].contiguous())          # (n_local, 3) exact canonical scores, on the device
    loc_i = np.stack([exact_local_topk(sc[:, j], lo, K)[0] for j in range(3)])
    loc_v = np.stack([exact_local_topk(sc[:, j], lo, K)[1] for j in range(3)])
    if world > 1:
        gi = [None] * world
        gv = [None] * world
        dist.all_gather_object(gi, loc_i)
        dist.all_gather_object(gv, loc_v)
        gi, gv = np.stack(gi), np.stack(gv)
    else:
        gi, gv = loc_i[None], loc_v[None]
    ei, ev = merge_topk(gi, gv, K)
    a_ok = np.array_equal(ei, ref_i[qa]) and np.array_equal(ev, ref_v[qa])
    ok &= a_ok
    notes.append(f"(a) exact columns of queries {qa}: ids and score bits equal: {a_ok}")
    del sc
    # ---- (b) CPU oracle on the returned rows that live in this rank's slab (+ their neighbours), 8 queries
    from oracle import avl_oracle as O

    qb = list(range(0, NQ, NQ // 8))
    qh = q[qb].cpu().numpy()
    b_ok, n_checked = True, 0
    for j, qq in enumerate(qb):
        ids = ref_i[qq]
        mine = ids[(ids >= lo) & (ids < lo + feat.shape[0])] - lo
        if mine.size == 0:
            continue
        rows = np.unique(np.clip(np.concatenate([mine, mine + 1, mine - 1]), 0, feat.shape[0] - 1))
        fr = feat[torch.from_numpy(rows).to(dev)].cpu().numpy()
        s = O.scores(fr, qh[j:j + 1])[:, 0]
        lookup = dict(zip(rows.tolist(), s.tolist()))
        for rid, val in zip(ids, ref_v[qq]):
            if lo <= rid < lo + feat.shape[0]:
                b_ok &= np.float32(lookup[int(rid - lo)]) == np.float32(val)
                n_checked += 1
        # order: (score desc, row asc)
        b_ok &= all((ref_v[qq][t] > ref_v[qq][t + 1]) or (ref_v[qq][t] == ref_v[qq][t + 1] and ids[t] < ids[t + 1]) for t in range(K - 1))
    flag = torch.tensor([1 if b_ok else 0, n_checked], device=dev)
    if world > 1:
        dist.all_reduce(flag[:1], op=dist.ReduceOp.MIN)
        dist.all_reduce(flag[1:], op=dist.ReduceOp.SUM)
    ok &= bool(flag[0].item())
    notes.append(f"(b) oracle on {int(flag[1].item())} returned (row, query) pairs: score bits equal, order right: {bool(flag[0].item())}")
    digest = hashlib.sha256(ref_i.tobytes() + ref_v.tobytes()).hexdigest()[:16]
    if rank == 0:
        n_total = N_SHARDS * shard_rows
        line = {"config": f"{n_total} x {DIM} map in {N_SHARDS} shards (seeds 1000..1007), {NQ} queries, top-{K}", "world": world,
                "ok": bool(ok), "digest_ids_scores": digest, "checks": notes,
                "ms_per_step": timing, "queries_per_s": {k_: NQ / (v / 1e3) for k_, v in timing.items()},
                "scaling": "strong (the 16M map is fixed, rows split over the ranks)"}
        print(json.dumps(line))
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        (out / f"sharded_index_check_n{world}.json").write_text(json.dumps(line) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
