"""CPU, world_size 2 over gloo: the slab-sharded top-k exchange + merge equals the single-map result.
The per-rank scoring is stubbed with the oracle (there is no GPU here); what is tested is the host
logic of ShardedMap (offsets, the all-gather, the merge and its tie rule)."""
import os
import socket

import numpy as np
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


def _file_worker(rank, world, port, path, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seen = {}

    def factory(slab):
        seen["slab"] = slab
        return OracleSlab(np.asarray(slab))

    sm = ShardedMap.from_file(path, local_factory=factory)
    feat, q = synth.index_inputs(5001, 64, 7, seed=3)
    lo, hi = slab_bounds(5001, world, rank)
    idx, val = sm.topk(q, 16)
    ri, rv = O.topk(O.scores(feat, q), 16)
    ok = (isinstance(seen["slab"], np.memmap) and seen["slab"].shape == (hi - lo, 64) and not seen["slab"].flags.writeable
          and np.array_equal(seen["slab"], feat[lo:hi]) and (sm.row_lo, sm.row_hi, sm.n_total) == (lo, hi, 5001)
          and sm.grid_pos.shape == (5001, 3) and np.array_equal(idx, ri) and np.array_equal(val, rv))
    out.put(bool(ok))
    dist.destroy_process_group()


def test_sharded_map_from_file_world2_gloo(tmp_path):
    """Every rank memory-maps the saved map and hands ONLY its row slab to its local map; results carry global ids."""
    from avlmaps_b200.utils import h5lite

    feat, _ = synth.index_inputs(5001, 64, 7, seed=3)
    path = tmp_path / "vlmaps.h5df"
    h5lite.write_file(path, {"grid_feat": feat, "grid_pos": np.zeros((5001, 3), np.int32), "weight": np.ones(5001, np.float32)})
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_file_worker, args=(r, 2, port, str(path), out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() is True and out.get() is True


def test_sharded_map_from_file_single_process_and_chunked_fallback(tmp_path):
    from avlmaps_b200.utils import h5lite

    feat, q = synth.index_inputs(700, 32, 3, seed=5)
    path = tmp_path / "m.h5df"
    h5lite.write_file(path, {"grid_feat": feat.astype(np.float64), "grid_pos": np.zeros((700, 3), np.int32)})
    sm = ShardedMap.from_file(path, local_factory=lambda a: OracleSlab(np.asarray(a)))   # float64 storage: converted copy
    assert sm.local.feat.dtype == np.float32 and not isinstance(sm.local.feat, np.memmap)
    idx, val = sm.topk(q, 4)
    ri, rv = O.topk(O.scores(feat, q), 4)
    assert np.array_equal(idx, ri) and np.array_equal(val, rv)


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


# ------------------------------------------------------------------------------------------ sharded build
class OracleSlabBuilder:
    """Stand-in for engine.DeviceBuilder restricted to a row slab: runs the C oracle on every frame, then keeps
    the voxels whose row lies in the slab.  Keys: the single-build voxel id (any strictly increasing function
    of first-touch order ranks the same as the device's frame_seq << 32 | sample position)."""

    def __init__(self, cfg, d):
        cs, gs = cfg["cell_size"], cfg["grid_size"]
        self.vh = int(cfg["pose_info"]["camera_height"] / cs)
        self.grid_shape = (gs, gs, self.vh)
        self.o = O.BuildOracle(gs, self.vh, cs, d, capacity=gs * gs * self.vh)
        self.lo, self.hi = 0, gs

    def set_slab(self, lo, hi):
        self.lo, self.hi = lo, hi

    def add_frame(self, *a, **k):
        self.o.add_frame(*a, **k)

    def _sel(self):
        full = self.o.export()
        rows = full["grid_pos"][:, 0]
        return full, np.nonzero((rows >= self.lo) & (rows < self.hi))[0]

    def export(self):
        full, sel = self._sel()
        local = -np.ones(full["grid_feat"].shape[0], np.int32)
        local[sel] = np.arange(sel.size)
        occ = np.where(full["occupied_ids"] >= 0, local[np.clip(full["occupied_ids"], 0, None)], -1).astype(np.int32)
        return dict(grid_feat=full["grid_feat"][sel], grid_pos=full["grid_pos"][sel], weight=full["weight"][sel],
                    grid_rgb=full["grid_rgb"][sel], occupied_ids=occ)

    def export_keys(self):
        return self._sel()[1].astype(np.uint64) * np.uint64(7) + np.uint64(3)


def _numpy_rank(keys_per_shard, shard):
    return np.searchsorted(np.sort(np.concatenate(keys_per_shard)), keys_per_shard[shard]).astype(np.int64)


def _build_worker(rank, world, port, out):
    from avlmaps_b200.sharded import ShardedBuilder

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = synth.map_config(48, 0.1, 1.6, [40, 0, 40, 0, 40, 30, 0, 0, 1], 2)
    poses = synth.circle_poses(3, radius=0.3)
    depths, rgbs, feats = synth.build_inputs(3, 60, 80, 49, 65, 8, seed=4)
    np.random.seed(11)
    sidx = [O.sample_order(60 * 80, 2) for _ in range(3)]
    b2c, bt = O.setup_transforms(cfg["pose_info"])
    tfs = O.frame_transforms(poses, b2c, bt)
    calib = np.array(cfg["cam_calib_mat"]).reshape(3, 3)
    sb = ShardedBuilder(OracleSlabBuilder(cfg, 8), rank_fn=_numpy_rank)
    for i in range(3):
        sb.add_frame(depths[i], feats[i], rgbs[i], sidx[i], np.linalg.inv(calib), calib, O.get_sim_cam_mat(49, 65), tfs[i])
    res = sb.finalize()
    ref = O.build_map(cfg, poses, depths, rgbs, feats, sidx, capacity=48 * 48 * 16)
    ok = (res["n_voxels_total"] == ref["grid_feat"].shape[0]
          and np.array_equal(res["occupied_ids"], ref["occupied_ids"])          # global ids on every rank
          and np.array_equal(res["grid_pos"], ref["grid_pos"][res["global_ids"]])
          and np.array_equal(res["grid_feat"], ref["grid_feat"][res["global_ids"]])
          and np.all(np.diff(res["global_ids"]) > 0)
          and (sb.row_lo, sb.row_hi) == slab_bounds(48, world, rank))
    # the slab-built shard answers queries with GLOBAL ids through ShardedMap(global_ids=...)
    q = synth.index_inputs(1, 8, 3, seed=6)[1]
    sm = ShardedMap(OracleSlab(res["grid_feat"]), global_ids=res["global_ids"])
    idx, val = sm.topk(q, 5)
    ri, rv = O.topk(O.scores(ref["grid_feat"], q), 5)
    out.put(bool(ok) and np.array_equal(idx, ri) and np.array_equal(val, rv))
    dist.destroy_process_group()


def test_sharded_build_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_build_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get() is True and out.get() is True


def test_frame_row_range_contains_every_row_the_reference_geometry_reaches():
    """The host-side frustum test that lets a slab's rank skip a frame must never exclude a row a point of that frame
    falls into: rows from the reference's own arithmetic (mapping_utils.py:239-246, 305-315, 345-349) on random pixels,
    depths and poses stay inside frame_row_range."""
    from avlmaps_b200.sharded import frame_row_range

    rng = np.random.default_rng(5)
    cfg = synth.map_config(256, 0.05, 1.6, [320, 0, 320, 0, 320, 240, 0, 0, 1], 1)
    b2c, bt = O.setup_transforms(cfg["pose_info"])
    poses = synth.circle_poses(40, radius=2.0)
    tfs = O.frame_transforms(poses, b2c, bt)
    kinv = np.linalg.inv(np.array(cfg["cam_calib_mat"], np.float64).reshape(3, 3))
    h, w, gs, cs = 480, 640, 256, 0.05
    narrow = 0
    for tf in tfs:
        lo, hi = frame_row_range((h, w), kinv, tf, gs, cs, 0.1, 6.0)
        u = rng.integers(0, w, 4000) + 0.5
        v = rng.integers(0, h, 4000) + 0.5
        z = rng.uniform(0.1001, 5.9999, 4000)
        p = (kinv @ np.stack([u, v, np.ones_like(u)])) * z
        gx = tf[0, :3] @ p + tf[0, 3]
        rows = (gs / 2 - np.trunc(gx / cs)).astype(np.int64)
        assert rows.min() >= lo and rows.max() <= hi
        narrow += (hi - lo) < gs
    assert narrow > 0      # the test is not vacuous: some frames cannot reach every row


def test_balanced_row_bounds_tile_the_grid_and_even_out_the_points():
    from avlmaps_b200.sharded import balanced_row_bounds

    cfg = synth.map_config(256, 0.05, 1.6, [320, 0, 320, 0, 320, 240, 0, 0, 1], 1)
    b2c, bt = O.setup_transforms(cfg["pose_info"])
    poses = synth.circle_poses(24, radius=2.0)
    tfs = O.frame_transforms(poses, b2c, bt)
    kinv = np.linalg.inv(np.array(cfg["cam_calib_mat"], np.float64).reshape(3, 3))
    rng = np.random.default_rng(2)
    frames = [dict(depth=rng.uniform(0.5, 6.0, (480, 640)).astype(np.float32), kinv=kinv, tf=tf) for tf in tfs]
    for world in (1, 2, 3, 8):
        b = balanced_row_bounds(frames, 256, 0.05, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == 256
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1)) and all(hi > lo for lo, hi in b)
    b8 = balanced_row_bounds(frames, 256, 0.05, 8)
    widths = [hi - lo for lo, hi in b8]
    assert max(widths) > 2 * min(widths)       # the centre slabs are much narrower than the edge slabs


def test_prepared_frames_are_fed_as_runs_and_skipped_frames_keep_their_numbers():
    """ShardedBuilder.add_prepared: frames whose frustum cannot reach the slab are counted (skip_frames), the others go
    to the local builder in runs, in order, so that every frame keeps its number in the (frame, sample) first-touch
    order; whatever the batch boundaries, the calls cover each frame exactly once.  A skipped frame really has no
    point in the slab (rows by the reference's arithmetic)."""
    from types import SimpleNamespace

    from avlmaps_b200.sharded import ShardedBuilder

    class Recorder:
        grid_shape, cs, mode = (256, 256, 32), 0.05, 0

        def __init__(self):
            self.calls, self.seq = [], 0

        def set_slab(self, lo, hi):
            self.slab = (lo, hi)

        def prepare_frames(self, frames):
            return SimpleNamespace(n=len(frames), frames=frames)

        def add_prepared(self, prep, start, count, **kw):
            assert count > 0
            self.calls.append(("add", self.seq, start, count))
            self.seq += count

        def skip_frames(self, n):
            self.calls.append(("skip", self.seq, None, n))
            self.seq += n

    cfg = synth.map_config(256, 0.05, 1.6, [320, 0, 320, 0, 320, 240, 0, 0, 1], 1)
    b2c, bt = O.setup_transforms(cfg["pose_info"])
    tfs = O.frame_transforms(synth.circle_poses(60, radius=2.0), b2c, bt)
    kinv = np.linalg.inv(np.array(cfg["cam_calib_mat"], np.float64).reshape(3, 3))
    depth = np.zeros((480, 640), np.float32)
    frames = [dict(depth=depth, kinv=kinv, tf=tf) for tf in tfs]
    rng = np.random.default_rng(1)
    for bounds in ((0, 54), (120, 141), (202, 256)):
        for batch in (1, 7, 16):
            rec = Recorder()
            sb = ShardedBuilder(rec, row_bounds=bounds)
            assert rec.slab == bounds
            prep = sb.prepare_frames(frames)
            for i in range(0, 60, batch):
                sb.add_prepared(prep, i, min(batch, 60 - i), stream=None)
            assert rec.seq == 60
            covered = []
            for kind, seq, start, n in rec.calls:
                covered += list(range(seq, seq + n))
                if kind == "add":
                    assert start == seq                       # the frame number on the device == its index in the list
                    assert all(prep.touches[start:start + n])
                else:
                    assert not any(prep.touches[seq:seq + n])
            assert covered == list(range(60))
            assert getattr(sb, "n_skipped", 0) == sum(1 for t in prep.touches if not t)
        assert 0 < sum(prep.touches) < 60 or bounds == (120, 141)
        for i in np.nonzero(~np.array(prep.touches))[0]:
            u, v = rng.integers(0, 640, 2000) + 0.5, rng.integers(0, 480, 2000) + 0.5
            z = rng.uniform(0.1001, 5.9999, 2000)
            p = (kinv @ np.stack([u, v, np.ones_like(u)])) * z
            rows = (128 - np.trunc((tfs[i][0, :3] @ p + tfs[i][0, 3]) / 0.05)).astype(np.int64)
            assert not np.any((rows >= bounds[0]) & (rows < bounds[1]))
