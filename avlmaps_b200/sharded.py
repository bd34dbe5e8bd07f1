"""Slab-sharded map over the GPUs of one box: one process per GPU (torchrun), each rank owns a
contiguous slab of voxel rows; a query batch is scored against every slab independently and the
per-slab top-k (world * Q * k * 12 bytes) are exchanged and merged by ONE kernel over NVLink peer memory
(csrc/p2p_exchange.cu; AVL_P2P_EXCHANGE=0: one NCCL all-gather + a merge kernel).
The per-voxel argmax needs no exchange at all (it stays sharded).

torch.distributed is plumbing only; scoring runs in the C-ABI library."""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np


def merge_topk(idx: np.ndarray, val: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Merge per-shard results (S, Q, k) with GLOBAL row ids into the global top-k (Q, k), ordered by
    (score desc, row asc) -- the order each shard already uses, so ties stay on the lowest row."""
    s, q, kk = idx.shape
    fi = np.transpose(idx, (1, 0, 2)).reshape(q, s * kk)
    fv = np.transpose(val, (1, 0, 2)).reshape(q, s * kk).astype(np.float32)
    out_i = np.full((q, k), -1, np.int64)
    out_v = np.full((q, k), -np.inf, np.float32)
    for j in range(q):
        valid = fi[j] >= 0
        ci, cv = fi[j][valid], fv[j][valid]
        order = np.lexsort((ci, -cv.astype(np.float64)))[:k]
        out_i[j, :order.size] = ci[order]
        out_v[j, :order.size] = cv[order]
    return out_i, out_v


def merge_topk_torch(idx, val, k: int):
    """Same merge on device tensors (S, Q, k): sort by (score desc, row asc) with two stable sorts."""
    import torch

    s, q, kk = idx.shape
    fi = idx.permute(1, 0, 2).reshape(q, s * kk)
    fv = val.permute(1, 0, 2).reshape(q, s * kk)
    big = torch.iinfo(torch.int64).max
    key_i = torch.where(fi >= 0, fi, torch.full_like(fi, big))
    o1 = torch.argsort(key_i, dim=1, stable=True)
    fi, fv = torch.gather(fi, 1, o1), torch.gather(fv, 1, o1)
    o2 = torch.argsort(fv, dim=1, descending=True, stable=True)
    fi, fv = torch.gather(fi, 1, o2), torch.gather(fv, 1, o2)
    return fi[:, :k].contiguous(), fv[:, :k].contiguous()


def slab_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row slab of `rank` (the last slabs are one row shorter when world does not divide n)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class PendingTopK:
    """Result of ShardedMap.topk_async: (idx, val) CUDA tensors that become valid on a side stream."""

    def __init__(self, tensors, event):
        self._tensors, self._event = tensors, event

    def result(self):
        """Make torch's current stream wait for the exchange, return (idx int64 (Q, k), val float32 (Q, k)).  The
        tensors belong to a ring of 3 batches: copy them if they must outlive two more topk_async calls."""
        if self._event is not None:
            import torch

            torch.cuda.current_stream().wait_event(self._event)
        return self._tensors


class ShardedMap:
    """`local` is anything with .topk(queries, k, scale=, normalize_map=) and .argmax(...) over this
    rank's slab (an engine.DeviceMap in production); `row_offset` is the slab's first global row."""

    def __init__(self, local, row_offset: int = 0, group=None, global_ids=None):
        """`global_ids` (int64, one per local row, ascending) replaces `row_offset` when the slab's rows are
        not a contiguous id range -- the case of a slab-sharded BUILD, where global voxel ids are first-touch
        order across all slabs (ShardedBuilder.finalize)."""
        import torch.distributed as dist

        self.local = local
        self.row_offset = int(row_offset)
        self.global_ids = global_ids
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._p2p = None

    @classmethod
    def from_file(cls, map_path, group=None, local_factory=None, operand: str = "bf16") -> "ShardedMap":
        """Open a saved map (`vlmaps.h5df`, reference mapping_utils.py:469-505) slab-wise: every rank memory-maps
        `grid_feat` in place and uploads ONLY its own row slab, so a 16 M x 512 map (34 GB) costs each of 8 ranks
        4.3 GB of page-cache reads instead of the whole file in host RAM.  `grid_pos` (N x 3 int32, small) is kept
        whole on every rank as `.grid_pos` for the goal lookup `grid_pos[idx]` (habitat_lang_robot.py:427-430).
        `local_factory(slab_array) -> local map` defaults to engine.DeviceMap."""
        import torch.distributed as dist

        from .utils import h5lite

        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        with h5lite.File(map_path) as f:
            ds = f["grid_feat"]
            if len(ds.shape) != 2:
                raise ValueError(f"{map_path}: grid_feat must be (N, D), found {ds.shape}")
            lo, hi = slab_bounds(ds.shape[0], world, rank)
            if ds.offset is not None and ds.size and ds.dtype == np.dtype("<f4"):
                slab = ds.memmap()[lo:hi]          # a view: pages are read as the upload touches them
            else:
                slab = np.ascontiguousarray(ds.read()[lo:hi], dtype=np.float32)   # chunked / converted storage
            grid_pos = f["grid_pos"].read() if "grid_pos" in f else None
        if local_factory is None:
            from .engine import DeviceMap

            def local_factory(a):
                return DeviceMap(a, operand=operand)
        sm = cls(local_factory(slab), row_offset=lo, group=group)
        sm.n_total, sm.row_lo, sm.row_hi, sm.grid_pos = int(ds.shape[0]), lo, hi, grid_pos
        return sm

    # ---- device path -----------------------------------------------------------------------------------------
    _RING = 3   # per-slab result buffers in flight: the exchange of step i reads what the screen of step i wrote

    def _use_p2p(self, k: int) -> bool:
        # default: exchange + merge fused in ONE kernel over NVLink peer memory (csrc/p2p_exchange.cu);
        # AVL_P2P_EXCHANGE=0 selects the NCCL all-gather + merge kernel
        return os.environ.get("AVL_P2P_EXCHANGE", "1") != "0" and k * self.world <= 1024

    def _device_state(self, device, nq: int, k: int):
        import torch

        st = getattr(self, "_dev", None)
        if st is None or st["nq"] != nq or st["k"] != k:
            if st is not None:
                st["side"].synchronize()   # exchanges in flight still read the ring that is about to be released
            st = {"nq": nq, "k": k, "slot": 0,
                  "mine": [(torch.empty((nq, k), dtype=torch.int64, device=device),
                            torch.empty((nq, k), dtype=torch.float32, device=device)) for _ in range(self._RING)],
                  "out": [(torch.empty((nq, k), dtype=torch.int64, device=device),
                           torch.empty((nq, k), dtype=torch.float32, device=device)) for _ in range(self._RING)],
                  "scored": [torch.cuda.Event() for _ in range(self._RING)],
                  "merged": [None] * self._RING,
                  "side": torch.cuda.Stream(device=device)}
            if self.global_ids is not None:
                gid = self.global_ids if torch.is_tensor(self.global_ids) else torch.from_numpy(np.asarray(self.global_ids))
                self.global_ids = gid.to(device=device, dtype=torch.int64).contiguous()
            self._dev = st
        return st

    def topk_async(self, queries, k: int, scale=None, normalize_map: bool = False):
        """Device path (CUDA queries, at most 256): enqueue the slab's screen on the CURRENT stream and the peer exchange
        + merge on a side stream, return a `PendingTopK`; `.result()` makes the current stream wait for the merged
        (idx, val).  Nothing here synchronises the host, and the next batch's screen does not wait for this batch's
        exchange -- a slower peer delays the side stream only, until the ring of result buffers is used up (3 batches).
        No eager tensor ops: the slab's result lands in a ring buffer, the exchange kernel makes the ids global."""
        import torch

        if self.world > 1 and not self._use_p2p(k):
            return PendingTopK(self.topk(queries, k, scale=scale, normalize_map=normalize_map), None)   # NCCL form: in stream order
        nq = queries.shape[0]
        st = self._device_state(queries.device, nq, k)
        i = st["slot"]
        st["slot"] = (i + 1) % self._RING
        cur = torch.cuda.current_stream(queries.device)
        if st["merged"][i] is not None:
            cur.wait_event(st["merged"][i])        # the exchange that last read this ring slot (3 batches ago)
        ti, tv = st["mine"][i]
        if self.world == 1:
            self.local.topk(queries, k, scale=scale, normalize_map=normalize_map, out=(ti, tv), stats=False)
            return PendingTopK(self._globalize_single(ti, tv), None)
        self.local.topk(queries, k, scale=scale, normalize_map=normalize_map, out=(ti, tv), stats=False)
        st["scored"][i].record(cur)
        side = st["side"]
        side.wait_event(st["scored"][i])
        if self._p2p is None:
            from .engine import P2PExchange

            self._p2p = P2PExchange(self.group)
        gid = self.global_ids if self.global_ids is not None else None
        self._p2p.exchange_merge(ti, tv, row_offset=0 if gid is not None else self.row_offset, global_ids=gid,
                                 out=st["out"][i], stream=side.cuda_stream)
        ev = torch.cuda.Event()
        ev.record(side)
        st["merged"][i] = ev
        return PendingTopK(st["out"][i], ev)

    def _globalize_single(self, ti, tv):
        import torch

        if self.global_ids is not None:
            return torch.where(ti >= 0, self.global_ids[ti.clamp(min=0)], ti), tv
        if self.row_offset:
            return torch.where(ti >= 0, ti + self.row_offset, ti), tv
        return ti, tv

    def topk(self, queries, k: int, scale=None, normalize_map: bool = False):
        import torch
        import torch.distributed as dist

        on_device = type(queries).__module__.startswith("torch") and queries.is_cuda
        if on_device and self.world > 1 and queries.shape[0] <= 256 and self._use_p2p(k):
            res = self.topk_async(queries, k, scale=scale, normalize_map=normalize_map).result()
            if self._p2p is not None and getattr(self, "check_exchange", False):
                src = self._p2p.timed_out_source()   # synchronises: opt-in (tools / tests)
                if src >= 0:
                    raise RuntimeError(f"peer exchange timed out waiting for rank {src}")
            return res
        if on_device and self.world > 1 and queries.shape[0] <= 256:
            # NCCL form (AVL_P2P_EXCHANGE=0): the slab's result goes into a packed (ids | scores) buffer, ONE all-gather
            # moves world * Q * k * 12 bytes, one kernel of the library merges
            nq = queries.shape[0]
            nb_i, nb_v = nq * k * 8, nq * k * 4
            mine = torch.empty(nb_i + nb_v, dtype=torch.uint8, device=queries.device)
            ti = mine[:nb_i].view(torch.int64).view(nq, k)
            tv = mine[nb_i:].view(torch.float32).view(nq, k)
            self.local.topk(queries, k, scale=scale, normalize_map=normalize_map, out=(ti, tv), stats=False)
            if self.global_ids is not None:
                gid = self.global_ids if torch.is_tensor(self.global_ids) else torch.from_numpy(np.asarray(self.global_ids))
                self.global_ids = gid = gid.to(queries.device)
                ti.copy_(torch.where(ti >= 0, gid[ti.clamp(min=0)], ti))
            elif self.row_offset:
                ti += (ti >= 0) * self.row_offset
            gathered = torch.empty((self.world, nb_i + nb_v), dtype=torch.uint8, device=queries.device)
            dist.all_gather_into_tensor(gathered.view(-1), mine, group=self.group)   # the one collective
            gi = gathered[:, :nb_i].contiguous().view(torch.int64).view(self.world, nq, k)
            gv = gathered[:, nb_i:].contiguous().view(torch.float32).view(self.world, nq, k)
            from .engine import merge_topk_device

            return merge_topk_device(gi, gv, k)
        idx, val = self.local.topk(queries, k, scale=scale, normalize_map=normalize_map)
        if self.world == 1 and self.row_offset == 0 and self.global_ids is None:
            return idx, val  # one slab: already global ids in final order
        as_numpy = isinstance(idx, np.ndarray)
        ti = torch.from_numpy(idx) if as_numpy else idx
        tv = torch.from_numpy(val) if as_numpy else val
        if self.global_ids is not None:
            gid = self.global_ids if torch.is_tensor(self.global_ids) else torch.from_numpy(np.asarray(self.global_ids))
            gid = gid.to(ti.device)
            ti = torch.where(ti >= 0, gid[ti.clamp(min=0)], ti)
        else:
            ti = torch.where(ti >= 0, ti + self.row_offset, ti)
        if self.world > 1:
            gi = torch.empty((self.world,) + tuple(ti.shape), dtype=ti.dtype, device=ti.device)
            gv = torch.empty((self.world,) + tuple(tv.shape), dtype=tv.dtype, device=tv.device)
            dist.all_gather_into_tensor(gi.view(-1, ti.shape[-1]), ti.contiguous(), group=self.group)
            dist.all_gather_into_tensor(gv.view(-1, tv.shape[-1]), tv.contiguous(), group=self.group)
        else:
            gi, gv = ti[None], tv[None]
        mi, mv = merge_topk_torch(gi, gv, k)
        return (mi.numpy(), mv.numpy()) if as_numpy else (mi, mv)

    def argmax(self, queries, scale=None, normalize_map: bool = False):
        """Per-voxel argmax of this rank's slab; rows are independent, nothing to exchange."""
        return self.local.argmax(queries, scale=scale, normalize_map=normalize_map)


def frame_row_range(shape_hw, kinv, tf, gs: int, cs: float, min_depth: float = 0.1, max_depth: float = 6.0,
                    mode: int = 0, origin_x: float = 0.0):
    """Grid rows a frame's points can reach, (lo, hi) inclusive with a 2-cell margin: a conservative host-side test, no
    depth image needed.  A back-projected point is p = (Kinv @ [u + .5, v + .5, 1]) * z (mapping_utils.py:239-246) and
    its row depends on g.x = (tf @ [p; 1])[0] only (mapping_utils.py:345-349; vlmap_builder_multi_floor.py:146).  g.x is
    multilinear in (u, v, z) over the box [0, w] x [0, h] x [min_depth, max_depth], so its extremes sit at the 8 corners."""
    h, w = shape_hw
    kinv = np.asarray(kinv, np.float64).reshape(3, 3)
    tf = np.asarray(tf, np.float64).reshape(4, 4)
    xs = []
    for u in (0.0, float(w)):
        for v in (0.0, float(h)):
            ray = kinv @ np.array([u, v, 1.0])
            for z in (float(min_depth), float(max_depth)):
                xs.append(float(tf[0, :3] @ (ray * z) + tf[0, 3]))
    x_lo, x_hi = min(xs), max(xs)
    if mode == 0:    # row = int(gs / 2 - int(x / cs))
        return int(np.floor(gs / 2 - x_hi / cs)) - 2, int(np.ceil(gs / 2 - x_lo / cs)) + 2
    return int(np.floor((x_lo - origin_x) / cs)) - 2, int(np.ceil((x_hi - origin_x) / cs)) + 2


def balanced_row_bounds(frames, gs: int, cs: float, world: int, max_frames: int = 64, pixel_stride: int = 8,
                        min_depth: float = 0.1, max_depth: float = 6.0):
    """Row slabs [(lo, hi)] * world that hold about the same number of POINTS each, from a pilot histogram: a subsample
    of the frames (dicts with depth, kinv, tf as for add_frames) and of their pixels is back-projected on the host
    (float64 numpy, the reference's formulas; only the row is needed) and the grid rows are cut at the quantiles.
    Equal row counts leave the centre ranks with 2-3x the points of the edge ranks (the timing is the max over ranks).
    Every rank must call it with the same frames -- the result is a pure function of them."""
    hist = np.zeros(gs, np.float64)
    pick = np.unique(np.linspace(0, len(frames) - 1, min(max_frames, len(frames))).astype(int)) if len(frames) else []
    for i in pick:
        fr = frames[i]
        depth = fr["depth"]
        depth = depth.detach().cpu().numpy() if hasattr(depth, "detach") else np.asarray(depth)
        h, w = depth.shape
        z = depth[::pixel_stride, ::pixel_stride].astype(np.float64)
        if z.dtype.kind == "u" or depth.dtype == np.uint16:
            z = z / 1000.0
        v, u = np.meshgrid(np.arange(0, h, pixel_stride) + 0.5, np.arange(0, w, pixel_stride) + 0.5, indexing="ij")
        kinv = np.asarray(fr["kinv"], np.float64).reshape(3, 3)
        tf = np.asarray(fr["tf"], np.float64).reshape(4, 4)
        ray = kinv @ np.stack([u.ravel(), v.ravel(), np.ones(u.size)])
        p = ray * z.ravel()
        ok = (p[2] > fr.get("min_depth", min_depth)) & (p[2] < fr.get("max_depth", max_depth))
        gx = tf[0, :3] @ p + tf[0, 3]
        row = (gs / 2 - np.trunc(gx / cs)).astype(np.int64)
        ok &= (row >= 0) & (row < gs)
        hist += np.bincount(row[ok], minlength=gs)[:gs]
    if hist.sum() == 0:
        return [slab_bounds(gs, world, r) for r in range(world)]
    cum = np.cumsum(hist) / hist.sum()
    cuts = [0]
    for r in range(1, world):
        c = int(np.searchsorted(cum, r / world)) + 1
        cuts.append(min(max(c, cuts[-1] + 1), gs - (world - r)))   # every slab keeps at least one row
    cuts.append(gs)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class ShardedBuilder:
    """Slab-sharded map BUILD: one process per GPU, rank r owns the grid rows
    `slab_bounds(n_rows, world, r)`.  Every rank is fed every frame (depth, pose and the sample list are
    tiny; the features come from a replicated / sharded encoder) and runs the geometry for every sample, but
    fuses only the points that fall into its own rows -- so the frame loop has NO collective and the
    feature scatter, the dominant HBM traffic, divides by the number of GPUs.

    Voxel ids of the reference are first-touch order over the whole map.  Within a slab the local ids are
    already in that order; `finalize()` makes them global with ONE all-gather of the per-voxel first-touch
    keys (8 bytes per voxel) and a rank-by-binary-search kernel (avl_rank_keys), then one all-reduce(MAX)
    of the relabelled occupied_ids grid.  Results are identical to a single-GPU build.

    `local` is an engine.DeviceBuilder (anything with set_slab / add_frame / export / export_keys /
    grid_shape); `rank_fn(keys_per_shard, shard) -> int64 ids` defaults to engine.rank_keys."""

    def __init__(self, local, group=None, rank_fn=None, row_bounds=None):
        """`row_bounds`: this rank's (lo, hi) rows when the slabs are not the equal split -- e.g.
        balanced_row_bounds(frames, gs, cs, world)[rank]; the slabs of all ranks must tile [0, gs)."""
        import torch.distributed as dist

        self.local = local
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.row_lo, self.row_hi = (int(row_bounds[0]), int(row_bounds[1])) if row_bounds is not None else \
            slab_bounds(int(local.grid_shape[0]), self.world, self.rank)
        local.set_slab(self.row_lo, self.row_hi)
        self._rank_fn = rank_fn

    # Frames whose camera frustum cannot reach this rank's rows are not handed to the GPU at all (they only keep their
    # frame number): the geometry and the ordered id scan -- the part of a frame every rank used to repeat -- then
    # divide over the ranks like the scatter does.  A frame that reaches the slab's rows is processed as before.
    def _touches(self, depth, kinv, tf, min_depth=0.1, max_depth=6.0) -> bool:
        if getattr(self.local, "mode", 0) != 0:
            return True          # global-frame grids: rows wrap like numpy's negative indices; no frame is skipped
        shape = tuple(depth.shape[-2:])
        lo, hi = frame_row_range(shape, kinv, tf, int(self.local.grid_shape[0]), float(self.local.cs), min_depth, max_depth)
        return not (hi < self.row_lo or lo >= self.row_hi)

    def add_frame(self, *args, **kwargs):
        """engine.DeviceBuilder.add_frame's arguments (depth, feat, kinv, k, kfeat, tf, ...)."""
        if hasattr(self.local, "skip_frames"):
            names = ("depth", "feat", "kinv", "k", "kfeat", "tf")
            a = dict(zip(names, args))
            a.update({n: v for n, v in kwargs.items() if n in names})
            if all(n in a for n in ("depth", "kinv", "tf")) and not self._touches(
                    a["depth"], a["kinv"], a["tf"], kwargs.get("min_depth", 0.1), kwargs.get("max_depth", 6.0)):
                self.n_skipped = getattr(self, "n_skipped", 0) + 1
                return self.local.skip_frames(1)
        return self.local.add_frame(*args, **kwargs)

    def add_frames(self, frames, **kwargs):
        return self.add_prepared(self.prepare_frames(frames), **kwargs)

    def prepare_frames(self, frames):
        prep = self.local.prepare_frames(frames)
        prep.touches = [self._touches(fr["depth"], fr["kinv"], fr["tf"], fr.get("min_depth", 0.1), fr.get("max_depth", 6.0))
                        for fr in frames]
        return prep

    def add_prepared(self, prepared, start: int = 0, count=None, **kwargs):
        n = prepared.n - start if count is None else count
        touches = getattr(prepared, "touches", None)
        if touches is None:
            return self.local.add_prepared(prepared, start, n, **kwargs)
        i, end = start, start + n
        while i < end:     # runs of frames that reach the slab go to the GPU in one call; the gaps are only counted
            j = i
            while j < end and touches[j] == touches[i]:
                j += 1
            if touches[i]:
                self.local.add_prepared(prepared, i, j - i, **kwargs)
            else:
                self.local.skip_frames(j - i)
                self.n_skipped = getattr(self, "n_skipped", 0) + (j - i)
            i = j

    def _device(self):
        import torch
        import torch.distributed as dist

        if dist.is_initialized() and dist.get_backend(self.group) == "nccl":
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def finalize(self):
        """-> dict(global_ids (V_local,) int64, n_voxels_total, occupied_ids (global ids, full grid),
        grid_feat / grid_pos / weight / grid_rgb of THIS slab's voxels, rows in local (= ascending global) order)."""
        import torch
        import torch.distributed as dist

        out = self.local.export()
        keys = np.ascontiguousarray(self.local.export_keys(), np.uint64)
        if self.world > 1:
            dev = self._device()
            counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
            mine = torch.tensor([keys.size], dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(counts, mine, group=self.group)
            counts = counts.cpu().numpy()
            vmax = int(counts.max())
            pad = torch.zeros(max(vmax, 1), dtype=torch.int64, device=dev)
            pad[:keys.size] = torch.from_numpy(keys.view(np.int64)).to(dev)
            gathered = torch.empty((self.world, max(vmax, 1)), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(gathered.view(-1), pad, group=self.group)  # the one exchange of the build
            g = gathered.cpu().numpy().view(np.uint64)
            keys_per_shard = [g[r, :int(counts[r])] for r in range(self.world)]
        else:
            keys_per_shard = [keys]
        rank_fn = self._rank_fn
        if rank_fn is None:
            from .engine import rank_keys as rank_fn
        gids = np.asarray(rank_fn(keys_per_shard, self.rank), np.int64)
        total = int(sum(len(k) for k in keys_per_shard))
        occ = out["occupied_ids"]
        occ_g = np.where(occ >= 0, gids[np.clip(occ, 0, max(gids.size - 1, 0))] if gids.size else -1, -1).astype(np.int32)
        if self.world > 1:
            t = torch.from_numpy(occ_g).to(self._device())
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)  # other slabs hold -1 in this rank's rows
            occ_g = t.cpu().numpy()
        out.update(global_ids=gids, n_voxels_total=total, occupied_ids=occ_g)
        return out
