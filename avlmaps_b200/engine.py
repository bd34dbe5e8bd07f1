"""Pythonic handles over the C-ABI (include/avlmaps_b200.h): `DeviceMap` for the landmark-index
path, `DeviceBuilder` for the map-build path.  Inputs may be numpy arrays (host pointers, the
library does the H2D / D2H copies) or torch CUDA tensors (device pointers, nothing leaves HBM).

PyTorch is only plumbing here (device memory, streams, torch.distributed); every result comes from
the hand-written kernels in csrc/.  There is no CPU path: without a CUDA device every method raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib as L


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _stream_ptr(stream) -> Optional[int]:
    if stream is None:
        return None
    return int(getattr(stream, "cuda_stream", stream))


def _current_torch_stream():
    import torch

    return torch.cuda.current_stream().cuda_stream


class _Arg:
    """Uniform view of a numpy / torch argument: pointer + whether it is a device pointer."""

    def __init__(self, x, dtype, name):
        self.keep = None
        self.device = False
        self.ptr = None
        self.shape = None
        if x is None:
            return
        if _is_torch(x):
            import torch

            want = {np.float32: torch.float32, np.int32: torch.int32, np.int64: torch.int64, np.uint8: torch.uint8,
                    np.uint16: torch.uint16, np.uint64: torch.uint64, np.float16: torch.float16}[dtype]
            if x.dtype != want or not x.is_contiguous():
                x = x.to(want).contiguous()
            if not x.is_cuda:
                x = x.numpy()
            else:
                if x.data_ptr() & 15:
                    # a view that starts off a 16-byte boundary (t[1:], an odd offset into a packed buffer): the
                    # kernels read their inputs with 8 / 16-byte vector loads, so hand them an aligned copy
                    x = x.clone()
                self.keep, self.device, self.ptr, self.shape = x, True, C.c_void_p(x.data_ptr()), tuple(x.shape)
                return
        a = np.ascontiguousarray(x, dtype=dtype)
        # __array_interface__ is several times cheaper than a.ctypes (this runs per frame and per query batch)
        self.keep, self.ptr, self.shape = a, C.c_void_p(a.__array_interface__["data"][0]), a.shape


def _flags(*args: _Arg) -> int:
    dev = [a.device for a in args if a.ptr is not None]
    if any(dev) and not all(dev):
        raise ValueError("all array arguments of one call must be on the same side (all numpy or all torch.cuda)")
    return L.AVL_ON_DEVICE if dev and dev[0] else 0


class DeviceMap:
    """grid_feat (N, D) resident in HBM: fp32 copy (exact re-scoring), bf16 copy (tcgen05 screen),
    per-row norms and bf16 rounding residuals.  Created once per loaded map (VLMap.load_map)."""

    def __init__(self, grid_feat, stream=None, _handle=None, operand: str = "bf16"):
        """operand: element type of the tensor-core copy, "bf16" or "f16" (8x tighter error band, same speed and
        bytes; falls back to bf16 by itself if a value exceeds the fp16 range).  Results do not depend on it."""
        self._lib = L.load()
        L.require_device()
        self._h = C.c_void_p()
        if operand not in ("bf16", "f16"):
            raise ValueError("operand must be 'bf16' or 'f16'")
        if _handle is not None:
            self._h = _handle
        else:
            a = _Arg(grid_feat, np.float32, "grid_feat")
            if len(a.shape) != 2:
                raise ValueError("grid_feat must be (N, D)")
            flags = _flags(a) | (L.AVL_MAP_F16 if operand == "f16" else 0)
            L.check(self._lib.avl_map_create(a.ptr, a.shape[0], a.shape[1], flags, _stream_ptr(stream), C.byref(self._h)))
        n, d = C.c_int64(), C.c_int32()
        L.check(self._lib.avl_map_shape(self._h, C.byref(n), C.byref(d)))
        self.n, self.dim = n.value, d.value
        self.operand = "f16" if self._lib.avl_map_operand_f16(self._h) else "bf16"
        self.last_stats = None

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.avl_map_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    @property
    def device_bytes(self) -> int:
        return int(self._lib.avl_map_device_bytes(self._h))

    def _queries(self, queries, scale):
        q = _Arg(queries, np.float32, "queries")
        if len(q.shape) != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be (Q, {self.dim}), got {q.shape}")
        s = _Arg(scale, np.float32, "scale")
        if s.ptr is not None and tuple(s.shape) != (q.shape[0],):
            raise ValueError("scale must be (Q,)")
        return q, s

    def _out(self, like_device: bool, shape, dtype):
        if like_device:
            import torch

            t = torch.empty(shape, dtype={np.float32: torch.float32, np.int32: torch.int32, np.int64: torch.int64}[dtype],
                            device="cuda")
            return t, C.c_void_p(t.data_ptr())
        a = np.empty(shape, dtype)
        return a, a.ctypes.data_as(C.c_void_p)

    # -- scores (N, Q): the reference's `map_feats @ text_feats.T`
    def scores(self, queries, scale=None, normalize_map: bool = False, stream=None):
        q, s = self._queries(queries, scale)
        out = []
        for c0 in range(0, q.shape[0], L.AVL_MAX_QUERIES):
            qc = q.keep[c0:c0 + L.AVL_MAX_QUERIES]
            sc = None if s.ptr is None else s.keep[c0:c0 + L.AVL_MAX_QUERIES]
            qa, sa = _Arg(qc, np.float32, "q"), _Arg(sc, np.float32, "s")
            o, optr = self._out(q.device, (self.n, qa.shape[0]), np.float32)
            L.check(self._lib.avl_sim_dense(self._h, qa.ptr, qa.shape[0], sa.ptr, int(normalize_map), optr,
                                            _flags(qa, sa), _stream_ptr(stream)))
            out.append(o)
        if len(out) == 1:
            return out[0]
        if q.device:
            import torch

            return torch.cat(out, dim=1)
        return np.concatenate(out, axis=1)

    def screen_scores(self, queries, cta_group: int = 0, stream=None):
        """Diagnostic: raw bf16 tensor-core scores of the screen kernel."""
        q, _ = self._queries(queries, None)
        o, optr = self._out(q.device, (self.n, q.shape[0]), np.float32)
        L.check(self._lib.avl_sim_screen_dense(self._h, q.ptr, q.shape[0], cta_group, optr, _flags(q), _stream_ptr(stream)))
        return o

    # -- per-voxel argmax over the query batch (index_map)
    def argmax(self, queries, scale=None, normalize_map: bool = False, stream=None, want_stats: bool = False):
        q, s = self._queries(queries, scale)
        if q.shape[0] > L.AVL_MAX_QUERIES:
            # the fused kernel holds at most 256 queries; the reference has no limit (vlmap.py:123: np.argmax over any
            # number of categories), so wider batches take the exact scores and the first maximum per row
            sc = self.scores(queries, scale=scale, normalize_map=normalize_map, stream=stream)
            self.last_stats = None
            if q.device:
                import torch

                return sc.argmax(dim=1).to(torch.int32)
            return np.argmax(sc, axis=1).astype(np.int32)
        o, optr = self._out(q.device, (self.n,), np.int32)
        st = L.IndexStats()
        L.check(self._lib.avl_sim_argmax(self._h, q.ptr, q.shape[0], s.ptr, int(normalize_map), optr, _flags(q, s),
                                         _stream_ptr(stream), C.byref(st) if (want_stats or not q.device) else None))
        self.last_stats = st.as_dict()
        return o

    # -- per-query top-k rows
    def flush(self, stream=None) -> None:
        """After pipelined top-k calls: make `stream` (default: torch's current stream) wait for their tails."""
        sp = _stream_ptr(stream) if stream is not None else _current_torch_stream()
        L.check(self._lib.avl_map_flush(self._h, sp))

    def tail_stream(self):
        """The stream the tails of pipelined top-k calls run on, as a torch.cuda.ExternalStream."""
        import torch

        p = C.c_void_p()
        L.check(self._lib.avl_map_tail_stream(self._h, C.byref(p)))
        return torch.cuda.ExternalStream(p.value)

    def topk(self, queries, k: int, scale=None, normalize_map: bool = False, stream=None, out=None, stats: bool = True,
             pipelined: bool = False):
        """out=(idx int64 (Q, k), score float32 (Q, k)) torch CUDA tensors: write there (device path, Q <= 256).
        With out= and stats=False the call is asynchronous: it only enqueues work on `stream` (no host round trip --
        the exact fallback for overflowed queries is decided on the device) and `last_stats` is not updated.
        pipelined=True (with out=, stats=False): AVL_PIPELINED -- the call's tail runs on `tail_stream()` next to the
        following call's screen; its results are ordered on `stream` after the next pipelined call or `flush()`, and
        queries / out must stay untouched until then."""
        q, s = self._queries(queries, scale)
        if out is not None:
            if not q.device or q.shape[0] > L.AVL_MAX_QUERIES:
                raise ValueError("out= needs CUDA queries and at most 256 of them")
            st = L.IndexStats() if stats else None
            fl = _flags(q, s) | (L.AVL_PIPELINED if (pipelined and not stats) else 0)
            sp = _stream_ptr(stream) if stream is not None else (_current_torch_stream() if pipelined else None)
            L.check(self._lib.avl_sim_topk(self._h, q.ptr, q.shape[0], s.ptr, int(normalize_map), k,
                                           C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()),
                                           fl, sp, C.byref(st) if stats else None))
            if stats:
                self.last_stats = st.as_dict()
            return out
        outs_i, outs_s = [], []
        stats = None
        for c0 in range(0, q.shape[0], L.AVL_MAX_QUERIES):
            qa = _Arg(q.keep[c0:c0 + L.AVL_MAX_QUERIES], np.float32, "q")
            sa = _Arg(None if s.ptr is None else s.keep[c0:c0 + L.AVL_MAX_QUERIES], np.float32, "s")
            oi, oiptr = self._out(q.device, (qa.shape[0], k), np.int64)
            os_, osptr = self._out(q.device, (qa.shape[0], k), np.float32)
            st = L.IndexStats()
            L.check(self._lib.avl_sim_topk(self._h, qa.ptr, qa.shape[0], sa.ptr, int(normalize_map), k, oiptr, osptr,
                                           _flags(qa, sa), _stream_ptr(stream), C.byref(st)))
            stats = st.as_dict()
            outs_i.append(oi)
            outs_s.append(os_)
        self.last_stats = stats
        if len(outs_i) == 1:
            return outs_i[0], outs_s[0]
        if q.device:
            import torch

            return torch.cat(outs_i), torch.cat(outs_s)
        return np.concatenate(outs_i), np.concatenate(outs_s)


def merge_topk_device(gathered_idx, gathered_val, k: int, stream=None):
    """(S, Q, k) torch CUDA tensors with global row ids -> merged (Q, k) by (score desc, row asc)."""
    import torch

    lib = L.load()
    s_, q_, kk = gathered_idx.shape
    oi = torch.empty((q_, k), dtype=torch.int64, device=gathered_idx.device)
    ov = torch.empty((q_, k), dtype=torch.float32, device=gathered_idx.device)
    L.check(lib.avl_merge_topk(C.c_void_p(gathered_idx.data_ptr()), C.c_void_p(gathered_val.data_ptr()), s_, q_, kk,
                               C.c_void_p(oi.data_ptr()), C.c_void_p(ov.data_ptr()), L.AVL_ON_DEVICE, _stream_ptr(stream)))
    return oi, ov


class P2PExchange:
    """Fused exchange + merge of per-slab top-k results over NVLink peer memory (csrc/p2p_exchange.cu): one kernel
    per query batch instead of an NCCL all-gather plus a merge kernel.  One object per rank; the CUDA-IPC handles of
    the receive buffers travel once through `torch.distributed.all_gather_object`.  ShardedMap's default exchange."""

    def __init__(self, group=None, nq_max: int = L.AVL_MAX_QUERIES, k_max: Optional[int] = None):
        import torch.distributed as dist

        self._lib = L.load()
        L.require_device()
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nq_max = int(nq_max)
        self.k_max = int(k_max) if k_max is not None else min(L.AVL_MAX_TOPK, 1024 // self.world)
        self._h = C.c_void_p()
        L.check(self._lib.avl_p2p_create(self.rank, self.world, self.nq_max, self.k_max, C.byref(self._h)))
        nb = int(self._lib.avl_p2p_handle_bytes())
        mine = (C.c_uint8 * nb)()
        L.check(self._lib.avl_p2p_local_handle(self._h, mine))
        if self.world > 1:
            gathered = [None] * self.world
            dist.all_gather_object(gathered, bytes(mine), group=group)
            blob = b"".join(gathered)
            buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
            L.check(self._lib.avl_p2p_connect(self._h, buf))
            dist.barrier(group=group)   # nobody stores into a peer before every peer has mapped every buffer

    def exchange_merge(self, idx, val, row_offset: int = 0, stream=None, global_ids=None, out=None):
        """idx (nq, k) int64 row ids (-1 = empty; the others are made global inside the kernel: + `row_offset`, or a
        lookup in `global_ids`, an int64 CUDA tensor with one entry per slab row), val (nq, k) float32, both torch.cuda
        tensors of this rank's slab -> (idx, val) of the global top-k, identical on every rank.  Enqueues one kernel on
        `stream` (default: torch's current stream), no synchronisation.  Every rank issues the same sequence of calls;
        calls on one object go to one stream at a time."""
        import torch

        nq, k = idx.shape
        if not (idx.is_contiguous() and val.is_contiguous()):
            idx, val = idx.contiguous(), val.contiguous()
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.int64, device=idx.device),
                   torch.empty((nq, k), dtype=torch.float32, device=idx.device))
        sp = _stream_ptr(stream) if stream is not None else _current_torch_stream()
        gid = C.c_void_p(global_ids.data_ptr()) if global_ids is not None else None
        L.check(self._lib.avl_p2p_exchange_merge(self._h, C.c_void_p(idx.data_ptr()), C.c_void_p(val.data_ptr()), nq, k,
                                                 int(row_offset), gid, C.c_void_p(out[0].data_ptr()),
                                                 C.c_void_p(out[1].data_ptr()), L.AVL_ON_DEVICE, sp))
        return out

    def timed_out_source(self) -> int:
        """-1, or the rank whose data never arrived within the kernel's ~10 s watchdog (synchronises)."""
        s = C.c_int32(-1)
        L.check(self._lib.avl_p2p_status(self._h, C.byref(s), None))
        return int(s.value)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.avl_p2p_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def topk_vector(values, k: int, stream=None):
    """Exact top-k of a heat vector, (value desc, index asc): k = 1 is get_max_pos_3d's np.argmax."""
    lib = L.load()
    L.require_device()
    v = _Arg(values, np.float32, "values")
    n = int(np.prod(v.shape))
    if v.device:
        import torch

        oi = torch.empty(k, dtype=torch.int64, device="cuda")
        ov = torch.empty(k, dtype=torch.float32, device="cuda")
        L.check(lib.avl_topk_f32(v.ptr, n, k, C.c_void_p(oi.data_ptr()), C.c_void_p(ov.data_ptr()), L.AVL_ON_DEVICE,
                                 _stream_ptr(stream)))
        return oi, ov
    oi, ov = np.empty(k, np.int64), np.empty(k, np.float32)
    L.check(lib.avl_topk_f32(v.ptr, n, k, L.np_ptr(oi), L.np_ptr(ov), 0, _stream_ptr(stream)))
    return oi, ov


def fuse_topk(map_a: DeviceMap, queries_a, map_b: DeviceMap, queries_b, k: int, scale_a=None, scale_b=None,
              normalize_a: bool = False, normalize_b: bool = False, combine: int = L.FUSE_PRODUCT, stream=None):
    """Cross-modal goal: per pair j, heat = combine(minmax(score_a[:, j]), minmax(score_b[:, j])) -> top-k."""
    lib = L.load()
    qa, sa = map_a._queries(queries_a, scale_a)
    qb, sb = map_b._queries(queries_b, scale_b)
    if qa.shape[0] != qb.shape[0]:
        raise ValueError("fuse_topk needs the same number of queries per modality (pairs)")
    n_pairs = qa.shape[0]
    oi, oiptr = map_a._out(qa.device, (n_pairs, k), np.int64)
    oh, ohptr = map_a._out(qa.device, (n_pairs, k), np.float32)
    L.check(lib.avl_fuse_topk(map_a._h, qa.ptr, sa.ptr, int(normalize_a), map_b._h, qb.ptr, sb.ptr, int(normalize_b),
                              n_pairs, combine, k, oiptr, ohptr, _flags(qa, sa, qb, sb), _stream_ptr(stream)))
    return oi, oh


def heat_from_mask_3d(grid_pos, mask, cell_size: float = 0.05, decay_rate: float = 0.01, stream=None):
    """get_heatmap_from_mask_3d (visualize_utils.py:29-49) on the device: (N,) float32 heat."""
    lib = L.load()
    L.require_device()
    p = _Arg(grid_pos, np.int32, "grid_pos")
    if _is_torch(mask):
        import torch

        mask = mask.to(torch.uint8)
    else:
        mask = np.ascontiguousarray(mask).astype(np.uint8)
    m = _Arg(mask, np.uint8, "mask")
    n = p.shape[0]
    if len(p.shape) != 2 or p.shape[1] != 3 or tuple(m.shape) != (n,):
        raise ValueError("grid_pos must be (N, 3) and mask (N,)")
    if p.device:
        import torch

        out = torch.empty(n, dtype=torch.float32, device="cuda")
        optr = C.c_void_p(out.data_ptr())
    else:
        out = np.empty(n, np.float32)
        optr = out.ctypes.data_as(C.c_void_p)
    L.check(lib.avl_heat_from_mask_3d(p.ptr, m.ptr, n, float(cell_size), float(decay_rate), optr, _flags(p, m),
                                      _stream_ptr(stream)))
    return out


def heat_planar(grid_pos, row: float, col: float, con: float = 1.0, decay_rate: float = 0.01) -> np.ndarray:
    """Planar distance decay of AVLMap.index_image (avlmap.py:156-162): (N,) float64."""
    lib = L.load()
    L.require_device()
    p = np.ascontiguousarray(grid_pos, np.int32)
    if p.ndim != 2 or p.shape[1] != 3:
        raise ValueError("grid_pos must be (N, 3)")
    out = np.empty(p.shape[0], np.float64)
    L.check(lib.avl_heat_planar(L.np_ptr(p), p.shape[0], float(row), float(col), float(con), float(decay_rate),
                                L.np_ptr(out), 0, None))
    return out


def heat2d_sources(shape, cells_per_group, conf, decay_rate: float, mode: str):
    """2-D source heat before the final min-max.  mode "area": max-combine, float64 (avlmap.py:78-97);
    mode "sound": float32 running sum in group order (avlmap.py:111-131).  cells_per_group: list of (n_i, 2)
    int arrays of (row, col); an empty array is a frame that fell outside the grid."""
    lib = L.load()
    L.require_device()
    rows, cols = int(shape[0]), int(shape[1])
    starts = np.zeros(len(cells_per_group) + 1, np.int32)
    flat = []
    for i, c in enumerate(cells_per_group):
        c = np.asarray(c, np.int32).reshape(-1, 2)
        flat.append(c)
        starts[i + 1] = starts[i] + c.shape[0]
    cells = np.ascontiguousarray(np.concatenate(flat) if flat else np.zeros((0, 2), np.int32), np.int32)
    conf = np.ascontiguousarray(conf, np.float32)
    if conf.shape != (len(cells_per_group),):
        raise ValueError("one confidence per group")
    m = {"area": 0, "sound": 1}[mode]
    out = np.empty((rows, cols), np.float64 if m == 0 else np.float32)
    L.check(lib.avl_heat2d_sources(L.np_ptr(cells), L.np_ptr(starts), L.np_ptr(conf), len(cells_per_group), rows, cols,
                                   float(decay_rate), m, L.np_ptr(out), 0, None))
    return out


def heat2d_normalize_lift(heat2d, grid_pos=None, normalize: bool = True):
    """Min-max of a 2-D heat map in its own dtype (avlmap.py:97,131) and / or its lift to the voxels through
    grid_pos (avlmap.py:100-109,135-144), on the device.  Returns (heat2d normalised (a copy), heat3d float32 or None)."""
    lib = L.load()
    L.require_device()
    h = np.array(heat2d, copy=True, order="C")
    if h.ndim != 2 or h.dtype not in (np.float32, np.float64):
        raise ValueError("heat2d must be a (rows, cols) float32 or float64 array")
    pos, out3 = None, None
    if grid_pos is not None:
        pos = np.ascontiguousarray(grid_pos, np.int32)
        out3 = np.empty(pos.shape[0], np.float32)
    L.check(lib.avl_heat2d_normalize_lift(L.np_ptr(h), int(h.dtype == np.float64), h.shape[0], h.shape[1], int(normalize),
                                          L.np_ptr(pos), 0 if pos is None else pos.shape[0], L.np_ptr(out3), 0, None))
    return h, out3


_FRAME_OFFSETS = {name: getattr(L.Frame, name).offset for name in ("kinv", "k", "kfeat", "tf")}


def _is_u16(x) -> bool:
    """uint16 depth = millimetres (the multi-floor builder's PNG depth, vlmap_builder_multi_floor.py:103)."""
    if _is_torch(x):
        import torch

        return x.dtype == torch.uint16
    return np.asarray(x).dtype == np.uint16


def _is_f16(x) -> bool:
    """float16 features are handed over as they are (AVL_FEAT_F16): LSeg's output is fp16-exact (lseg_net.py:318-321)."""
    if x is None:
        return False
    if _is_torch(x):
        import torch

        return x.dtype == torch.float16
    return getattr(x, "dtype", None) == np.float16


def _fill_frame(depth, feat, kinv, k, kfeat, tf, rgb, sample_idx, feat_layout, min_depth, max_depth, dim=None):
    """Build the avl_frame struct; returns (frame, flags, keep-alive args)."""
    u16 = _is_u16(depth)
    f16 = _is_f16(feat)
    d_ = _Arg(depth, np.uint16 if u16 else np.float32, "depth")
    f_ = _Arg(feat, np.float16 if f16 else np.float32, "feat")
    r_ = _Arg(rgb, np.uint8, "rgb")
    s_ = _Arg(sample_idx, np.int32, "sample_idx")
    if len(d_.shape) != 2:
        raise ValueError("depth must be (H, W)")
    fh = fw = 0
    if f_.ptr is not None:
        if feat_layout == L.FEAT_CHW:
            if len(f_.shape) == 4 and f_.shape[0] == 1:
                _, dd, fh, fw = f_.shape
            elif len(f_.shape) == 3:
                dd, fh, fw = f_.shape
            else:
                raise ValueError("CHW features must be (1, D, FH, FW)")
        else:
            if len(f_.shape) != 3:
                raise ValueError("HWC features must be (FH, FW, D)")
            fh, fw, dd = f_.shape
        if dim is not None and dd != dim:
            raise ValueError(f"feature dim {dd} != builder dim {dim}")
    fr = L.Frame()
    fr.depth, fr.h, fr.w = d_.ptr, d_.shape[0], d_.shape[1]
    fr.feat, fr.fh, fr.fw, fr.feat_layout = f_.ptr, fh, fw, feat_layout
    fr.rgb = r_.ptr
    fr.sample_idx = s_.ptr
    fr.n_samples = 0 if s_.ptr is None else math.prod(s_.shape)
    if sample_idx is not None and fr.n_samples == 0:
        # an EMPTY sample list (not "every pixel", which is sample_idx=None): the C side tells the two apart by a
        # non-NULL pointer, and an empty tensor has none -- borrow the depth pointer, it is never dereferenced
        fr.sample_idx = d_.ptr
    base = C.addressof(fr)
    for name, m, n in (("kinv", kinv, 9), ("k", k, 9), ("kfeat", kfeat, 9), ("tf", tf, 16)):
        if m is None:
            continue
        arr = np.asarray(m, np.float64)
        if arr.size != n:
            raise ValueError(f"{name} must have {n} elements")
        C.memmove(base + _FRAME_OFFSETS[name], arr.tobytes(), n * 8)  # per-frame call: keep the Python cost down
    fr.min_depth, fr.max_depth = float(min_depth), float(max_depth)
    flags = _flags(d_, f_, r_, s_) | (L.AVL_DEPTH_U16_MM if u16 else 0) | (L.AVL_FEAT_F16 if f16 else 0)
    return fr, flags, (d_, f_, r_, s_)


class FrameBounds:
    """Pass 1 of VLMapBuilderMultiFloor.create_global_map (vlmap_builder_multi_floor.py:97-118): running
    min / max of the back-projected, transformed, sampled points -> pcd_min / pcd_max."""

    def __init__(self):
        self._lib = L.load()
        L.require_device()
        self._h = C.c_void_p()
        L.check(self._lib.avl_bounds_create(C.byref(self._h)))

    def add_frame(self, depth, kinv, tf, sample_idx=None, min_depth: float = 0.1, max_depth: float = 100.0, stream=None):
        fr, flags, keep = _fill_frame(depth, None, kinv, None, None, tf, None, sample_idx, L.FEAT_CHW, min_depth, max_depth)
        L.check(self._lib.avl_bounds_add_frame(self._h, C.byref(fr), flags, _stream_ptr(stream)))

    def get(self) -> Tuple[np.ndarray, np.ndarray, int]:
        mn, mx, n = (C.c_double * 3)(), (C.c_double * 3)(), C.c_int64()
        L.check(self._lib.avl_bounds_get(self._h, mn, mx, C.byref(n), None))
        return np.array(mn[:]), np.array(mx[:]), n.value

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.avl_bounds_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class DeviceBuilder:
    """Voxel map under construction in HBM (the arrays of VLMapBuilder._init_map, vlmap_builder.py:195-224)."""

    def __init__(self, gs: int, vh: int, cs: float, dim: int, capacity: Optional[int] = None):
        self._lib = L.load()
        L.require_device()
        self.gs, self.vh, self.cs, self.dim = int(gs), int(vh), float(cs), int(dim)
        self.grid_shape = (self.gs, self.gs, self.vh)
        spec = L.GridSpec(self.gs, self.vh, self.cs, self.dim, int(capacity or 0))
        self._h = C.c_void_p()
        L.check(self._lib.avl_builder_create(C.byref(spec), C.byref(self._h)))
        self.n_frames = 0
        self.mode = 0

    @classmethod
    def global_grid(cls, n_row: int, n_col: int, n_height: int, cs: float, pcd_min, dim: int,
                    capacity: Optional[int] = None) -> "DeviceBuilder":
        """The global-frame grid of VLMapBuilderMultiFloor._init_map (vlmap_builder_multi_floor.py:217-241):
        cells are np.round((p - pcd_min) / cs) as (row, height, col), occupied_ids is (n_row, n_col, n_height)."""
        self = cls.__new__(cls)
        self._lib = L.load()
        L.require_device()
        self.gs, self.vh, self.cs, self.dim = int(n_row), int(n_height), float(cs), int(dim)
        self.grid_shape = (int(n_row), int(n_col), int(n_height))
        spec = L.GlobalGridSpec()
        spec.n_row, spec.n_col, spec.n_height, spec.cs, spec.dim = int(n_row), int(n_col), int(n_height), float(cs), int(dim)
        spec.pcd_min[:] = [float(v) for v in np.asarray(pcd_min, np.float64).reshape(3)]
        spec.capacity = int(capacity or 0)
        self._h = C.c_void_p()
        L.check(self._lib.avl_builder_create_global(C.byref(spec), C.byref(self._h)))
        self.n_frames = 0
        return self

    def import_state(self, grid_feat, grid_pos, weight, grid_rgb=None) -> None:
        """Resume from a saved map like _init_map's reload (vlmap_builder.py:212-222): frames added afterwards
        are fused on top of it.  Fresh builder only."""
        f = np.ascontiguousarray(grid_feat, np.float32)
        p = np.ascontiguousarray(grid_pos, np.int32)
        w = np.ascontiguousarray(weight, np.float32)
        c = None if grid_rgb is None else np.ascontiguousarray(grid_rgb, np.uint8)
        if f.ndim != 2 or f.shape[1] != self.dim or p.shape != (f.shape[0], 3) or w.shape != (f.shape[0],):
            raise ValueError("saved map arrays do not match the builder (grid_feat (V, D), grid_pos (V, 3), weight (V,))")
        L.check(self._lib.avl_builder_import(self._h, L.np_ptr(f), L.np_ptr(p), L.np_ptr(w), L.np_ptr(c), f.shape[0], 0, None))

    def set_slab(self, row_lo: int, row_hi: int) -> None:
        """Own only the grid rows [row_lo, row_hi) (slab-sharded build, one process per GPU)."""
        L.check(self._lib.avl_builder_set_slab(self._h, int(row_lo), int(row_hi)))

    def skip_frames(self, n: int = 1) -> None:
        """Slab-sharded build: `n` frames whose frustum cannot reach this builder's rows keep their frame numbers."""
        L.check(self._lib.avl_builder_skip_frames(self._h, int(n)))
        self.n_frames += int(n)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.avl_builder_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def add_frame(self, depth, feat, kinv, k, kfeat, tf, rgb=None, sample_idx=None, feat_layout: int = L.FEAT_CHW,
                  min_depth: float = 0.1, max_depth: float = 6.0, stream=None):
        """depth (H, W) f32 metres (or uint16 millimetres); feat (1, D, FH, FW) [CHW] or (FH, FW, D) [HWC] f32;
        rgb (H, W, 3) u8 or None; sample_idx int32 pixel ids in the reference's sample order or None for every pixel."""
        if feat is None:
            raise ValueError("feat is required")
        fr, flags, keep = _fill_frame(depth, feat, kinv, k, kfeat, tf, rgb, sample_idx, feat_layout, min_depth, max_depth,
                                      dim=self.dim)
        L.check(self._lib.avl_builder_add_frame(self._h, C.byref(fr), flags, _stream_ptr(stream)))
        self.n_frames += 1

    def add_frames(self, frames, stream=None):
        """Several consecutive frames in one call: `frames` is a list of dicts with add_frame's arguments
        (depth, feat, kinv, k, kfeat, tf and optionally rgb, sample_idx, feat_layout, min_depth, max_depth).  With torch
        CUDA tensors and pixel-major features up to 16 frames share one launch triple; same result as a loop."""
        if len(frames) == 0:
            return
        self.add_prepared(PreparedFrames(self, frames), stream=stream)

    def prepare_frames(self, frames) -> "PreparedFrames":
        """Marshal a list of frame dicts (see add_frames) once.  For producers that write into a fixed ring of device
        buffers (an encoder's output slots): only the pose changes per frame -- `PreparedFrames.set_tf(i, tf)` -- and
        `add_prepared` then costs one ctypes call instead of ~25 us of Python per frame."""
        prep = PreparedFrames(self, frames)
        if prep.n and (prep.flags & L.AVL_ON_DEVICE):
            # scratch for the largest call add_prepared can make, now -- not in the middle of the frame loop
            per_frame = max((f.n_samples if f.sample_idx else f.h * f.w) for f in prep.arr[:prep.n])
            self.reserve(min(prep.n, L.AVL_MAX_BATCH) * per_frame)
        return prep

    def reserve(self, samples_per_call: int, stream=None) -> None:
        """Allocate the per-sample scratch for calls of up to `samples_per_call` samples ahead of time."""
        L.check(self._lib.avl_builder_reserve(self._h, int(samples_per_call), _stream_ptr(stream)))

    def add_prepared(self, prepared: "PreparedFrames", start: int = 0, count: Optional[int] = None, stream=None):
        n = prepared.n - start if count is None else count
        if n <= 0:
            return
        if start < 0 or start + n > prepared.n:
            raise IndexError("frame range outside the prepared list")
        ptr = C.cast(C.addressof(prepared.arr) + start * C.sizeof(L.Frame), C.POINTER(L.Frame))
        L.check(self._lib.avl_builder_add_frames(self._h, ptr, n, prepared.flags, _stream_ptr(stream)))
        self.n_frames += n

    @property
    def h2d_bytes(self) -> int:
        """Bytes uploaded from host arrays so far (see avl_builder_h2d_bytes)."""
        n = C.c_int64()
        L.check(self._lib.avl_builder_h2d_bytes(self._h, C.byref(n)))
        return int(n.value)

    def _count(self, fn) -> int:
        n = C.c_int64()
        L.check(fn(self._h, C.byref(n), None))
        return n.value

    @property
    def num_voxels(self) -> int:
        return self._count(self._lib.avl_builder_num_voxels)

    @property
    def num_accepted(self) -> int:
        return self._count(self._lib.avl_builder_num_accepted)

    @property
    def num_rejected_oob(self) -> int:
        """global-frame grid: points on which the reference would raise IndexError (rejected here)."""
        return self._count(self._lib.avl_builder_num_rejected_oob)

    def export(self, want_rgb: bool = True, want_feat: bool = True):
        """numpy arrays[:max_id] + occupied_ids, like _save_3d_map (vlmap_builder.py:313-327)."""
        v = self.num_voxels
        out = dict(
            grid_feat=np.zeros((v if want_feat else 0, self.dim), np.float32),
            grid_pos=np.zeros((v, 3), np.int32),
            weight=np.zeros((v,), np.float32),
            occupied_ids=np.empty(self.grid_shape, np.int32),
            grid_rgb=np.zeros((v, 3), np.uint8),
        )
        L.check(self._lib.avl_builder_export(self._h, L.np_ptr(out["grid_feat"]) if want_feat else None, L.np_ptr(out["grid_pos"]),
                                             L.np_ptr(out["weight"]), L.np_ptr(out["occupied_ids"]),
                                             L.np_ptr(out["grid_rgb"]) if want_rgb else None, 0, None))
        return out

    def export_keys(self) -> np.ndarray:
        """First-touch keys (frame_seq << 32 | sample position) of the voxels, ascending: uint64 (V,)."""
        keys = np.zeros((self.num_voxels,), np.uint64)
        if keys.size:
            L.check(self._lib.avl_builder_export_keys(self._h, L.np_ptr(keys), 0, None))
        return keys

    def to_map(self) -> DeviceMap:
        h = C.c_void_p()
        L.check(self._lib.avl_builder_to_map(self._h, None, C.byref(h)))
        return DeviceMap(None, _handle=h)


class PreparedFrames:
    """avl_frame array built once from frame dicts; keeps the argument buffers alive."""

    def __init__(self, builder: "DeviceBuilder", frames):
        self.n = len(frames)
        self.arr = (L.Frame * max(self.n, 1))()
        self.flags, self._keep = 0, []
        self.frames = frames     # the dicts (ShardedBuilder reads pose / intrinsics / shape for its frustum test)
        for i, fr in enumerate(frames):
            if fr.get("feat") is None:
                raise ValueError("feat is required")
            f, fl, k = _fill_frame(fr["depth"], fr["feat"], fr["kinv"], fr["k"], fr["kfeat"], fr["tf"], fr.get("rgb"),
                                   fr.get("sample_idx"), fr.get("feat_layout", L.FEAT_CHW), fr.get("min_depth", 0.1),
                                   fr.get("max_depth", 6.0), dim=builder.dim)
            if i and fl != self.flags:
                raise ValueError("all frames of one call must be on the same side and use the same depth type")
            self.flags = fl
            C.memmove(C.addressof(self.arr) + i * C.sizeof(L.Frame), C.addressof(f), C.sizeof(L.Frame))
            self._keep.append(k)

    def set_tf(self, i: int, tf) -> None:
        a = np.asarray(tf, np.float64)
        if a.size != 16:
            raise ValueError("tf must have 16 elements")
        C.memmove(C.addressof(self.arr) + i * C.sizeof(L.Frame) + _FRAME_OFFSETS["tf"], a.tobytes(), 128)


def rank_keys(keys_per_shard: Sequence[np.ndarray], shard: int) -> np.ndarray:
    """Global first-touch voxel ids of shard `shard`'s voxels: the rank of each of its keys among the keys of
    all shards (each list ascending, keys unique).  int64 (V_shard,)."""
    lib = L.load()
    L.require_device()
    keys_all = np.ascontiguousarray(np.concatenate([np.asarray(k, np.uint64) for k in keys_per_shard]))
    offsets = np.zeros(len(keys_per_shard) + 1, np.int64)
    offsets[1:] = np.cumsum([len(k) for k in keys_per_shard])
    out = np.zeros(int(offsets[shard + 1] - offsets[shard]), np.int64)
    L.check(lib.avl_rank_keys(L.np_ptr(keys_all), L.np_ptr(offsets), len(keys_per_shard), int(shard), L.np_ptr(out), 0, None))
    return out
