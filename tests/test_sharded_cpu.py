"""CPU, world_size 2 over gloo: the slab-sharded top-k exchange + merge equals the single-map result.
The per-rank scoring is stubbed with the oracle (there is no GPU here); what is tested is the host
logic of ShardedMap (offsets, the all-gather, the merge and its tie rule)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth
from avlmaps_b200.sharded import ShardedMap, merge_topk, merge_topk_torch, slab_bounds
from oracle import avl_oracle as O


class OracleSlab:
    def __init__(self, feat):
        self.feat = feat

    def topk(self, queries, k, scale=None, normalize_map=False):
        return O.topk(O.scores(self.feat, queries, scale=scale, normalize=normalize_map), k)

    def argmax(self, queries, scale=None, normalize_map=False):
        return O.argmax(O.scores(self.feat, queries, scale=scale, normalize=normalize_map))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    feat, q = synth.index_inputs(5001, 64, 7, seed=3)
    feat[4000] = feat[10]  # a cross-shard exact tie: the lower global row must win
    lo, hi = slab_bounds(feat.shape[0], world, rank)
    sm = ShardedMap(OracleSlab(feat[lo:hi]), lo)
    idx, val = sm.topk(q, 16)
    if rank == 0:
        ri, rv = O.topk(O.scores(feat, q), 16)
        out.put((np.array_equal(idx, ri), np.array_equal(val, rv)))
    dist.destroy_process_group()


def test_sharded_topk_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() == (True, True)


def test_slab_bounds_cover_all_rows():
    for n, w in ((10, 3), (16_777_216, 8), (5, 8), (0, 2)):
        spans = [slab_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))


def test_merge_numpy_and_torch_agree_with_ties_and_padding():
    rng = np.random.default_rng(0)
    idx = rng.permutation(4 * 5 * 8).reshape(4, 5, 8).astype(np.int64)
    val = rng.integers(0, 6, (4, 5, 8)).astype(np.float32)  # many ties
    idx[3, :, 5:] = -1
    val[3, :, 5:] = -np.inf
    mi, mv = merge_topk(idx, val, 8)
    ti, tv = merge_topk_torch(torch.from_numpy(idx), torch.from_numpy(val), 8)
    assert np.array_equal(mi, ti.numpy()) and np.array_equal(mv, tv.numpy())
    for j in range(5):
        flat_i, flat_v = idx[:, j].reshape(-1), val[:, j].reshape(-1)
        keep = flat_i >= 0
        order = np.lexsort((flat_i[keep], -flat_v[keep]))[:8]
        assert np.array_equal(mi[j], flat_i[keep][order])
