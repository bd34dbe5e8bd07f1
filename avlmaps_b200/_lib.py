"""ctypes binding of libavlmaps_b200.so (the C-ABI declared in include/avlmaps_b200.h).

The library is the only compute path: if it cannot be loaded, importing the engine fails loudly.
There is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libavlmaps_b200.so"

AVL_OK = 0
AVL_ON_DEVICE = 1
AVL_DEPTH_U16_MM = 2
AVL_MAP_F16 = 4
AVL_ASYNC = 16
AVL_PIPELINED = 32
AVL_FEAT_F16 = 8
AVL_MAX_QUERIES = 256
AVL_MAX_TOPK = 128
AVL_MAX_BATCH = 16     # frames per launch triple (kMaxBatch in csrc/build_path.cu)
FUSE_PRODUCT, FUSE_MAX, FUSE_SUM = 0, 1, 2
FEAT_CHW, FEAT_HWC = 0, 1

# every symbol include/avlmaps_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "avl_version", "avl_last_error", "avl_device_count", "avl_set_device", "avl_set_profiling",
    "avl_map_create", "avl_map_destroy", "avl_map_shape", "avl_map_device_bytes", "avl_map_operand_f16",
    "avl_map_screen_times", "avl_map_flush", "avl_map_tail_stream",
    "avl_sim_dense", "avl_sim_argmax", "avl_sim_topk", "avl_sim_screen_dense", "avl_topk_f32", "avl_fuse_topk",
    "avl_heat_from_mask_3d", "avl_merge_topk", "avl_heat2d_sources",
    "avl_builder_create", "avl_builder_destroy", "avl_builder_add_frame", "avl_builder_num_voxels",
    "avl_builder_num_accepted", "avl_builder_h2d_bytes", "avl_builder_export", "avl_builder_to_map",
    "avl_builder_create_global", "avl_builder_num_rejected_oob",
    "avl_bounds_create", "avl_bounds_destroy", "avl_bounds_add_frame", "avl_bounds_get",
    "avl_builder_set_slab", "avl_builder_skip_frames", "avl_builder_reserve", "avl_builder_export_keys", "avl_rank_keys", "avl_builder_import", "avl_builder_add_frames",
    "avl_heat_planar", "avl_heat2d_normalize_lift",
    "avl_p2p_create", "avl_p2p_handle_bytes", "avl_p2p_local_handle", "avl_p2p_connect", "avl_p2p_exchange_merge",
    "avl_p2p_status", "avl_p2p_destroy",
]


class IndexStats(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64), ("dim", C.c_int32), ("n_queries", C.c_int32),
        ("cta_group", C.c_int32), ("n_launches", C.c_int32),
        ("n_flagged", C.c_int64), ("n_candidates", C.c_int64),
        ("n_fallback_queries", C.c_int32), ("sample_rows", C.c_int32),
        ("ms_screen", C.c_float), ("ms_total", C.c_float),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class GridSpec(C.Structure):
    _fields_ = [("gs", C.c_int32), ("vh", C.c_int32), ("cs", C.c_double), ("dim", C.c_int32),
                ("capacity", C.c_int64)]


class GlobalGridSpec(C.Structure):
    _fields_ = [("n_row", C.c_int32), ("n_col", C.c_int32), ("n_height", C.c_int32), ("cs", C.c_double),
                ("pcd_min", C.c_double * 3), ("dim", C.c_int32), ("capacity", C.c_int64)]


class Frame(C.Structure):
    _fields_ = [
        ("depth", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32),
        ("feat", C.c_void_p), ("fh", C.c_int32), ("fw", C.c_int32), ("feat_layout", C.c_int32),
        ("rgb", C.c_void_p), ("sample_idx", C.c_void_p), ("n_samples", C.c_int32),
        ("kinv", C.c_double * 9), ("k", C.c_double * 9), ("kfeat", C.c_double * 9), ("tf", C.c_double * 16),
        ("min_depth", C.c_double), ("max_depth", C.c_double),
    ]


class AvlError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m avlmaps_b200._build` "
            "(avlmaps_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(str(LIB_PATH), mode=os.RTLD_LOCAL | os.RTLD_NOW)
    lib.avl_last_error.restype = C.c_char_p
    lib.avl_map_device_bytes.restype = C.c_int64
    lib.avl_map_device_bytes.argtypes = [C.c_void_p]
    vp, i32, i64, f32p = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    lib.avl_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.avl_set_device.argtypes = [C.c_int]
    lib.avl_set_profiling.argtypes = [C.c_int]
    lib.avl_map_create.argtypes = [f32p, i64, i32, C.c_int, vp, C.POINTER(vp)]
    lib.avl_map_destroy.argtypes = [vp]
    lib.avl_map_operand_f16.argtypes = [vp]
    lib.avl_map_shape.argtypes = [vp, C.POINTER(i64), C.POINTER(i32)]
    lib.avl_map_screen_times.argtypes = [vp, C.POINTER(C.c_float), i32, C.POINTER(i32)]
    lib.avl_map_flush.argtypes = [vp, vp]
    lib.avl_map_tail_stream.argtypes = [vp, C.POINTER(vp)]
    lib.avl_sim_dense.argtypes = [vp, f32p, i32, f32p, C.c_int, f32p, C.c_int, vp]
    lib.avl_sim_argmax.argtypes = [vp, f32p, i32, f32p, C.c_int, vp, C.c_int, vp, C.POINTER(IndexStats)]
    lib.avl_sim_topk.argtypes = [vp, f32p, i32, f32p, C.c_int, i32, vp, vp, C.c_int, vp, C.POINTER(IndexStats)]
    lib.avl_sim_screen_dense.argtypes = [vp, f32p, i32, i32, f32p, C.c_int, vp]
    lib.avl_topk_f32.argtypes = [f32p, i64, i32, vp, vp, C.c_int, vp]
    lib.avl_fuse_topk.argtypes = [vp, f32p, f32p, C.c_int, vp, f32p, f32p, C.c_int, i32, i32, i32, vp, vp,
                                  C.c_int, vp]
    lib.avl_heat2d_sources.argtypes = [vp, vp, vp, i32, i32, i32, C.c_double, i32, vp, C.c_int, vp]
    lib.avl_merge_topk.argtypes = [vp, vp, i32, i32, i32, vp, vp, C.c_int, vp]
    lib.avl_heat_from_mask_3d.argtypes = [vp, vp, i64, C.c_double, C.c_double, vp, C.c_int, vp]
    lib.avl_heat2d_normalize_lift.argtypes = [vp, i32, i32, i32, i32, vp, i64, vp, C.c_int, vp]
    lib.avl_heat_planar.argtypes = [vp, i64, C.c_double, C.c_double, C.c_double, C.c_double, vp, C.c_int, vp]
    lib.avl_builder_create.argtypes = [C.POINTER(GridSpec), C.POINTER(vp)]
    lib.avl_builder_destroy.argtypes = [vp]
    lib.avl_builder_add_frame.argtypes = [vp, C.POINTER(Frame), C.c_int, vp]
    lib.avl_builder_add_frames.argtypes = [vp, C.POINTER(Frame), i32, C.c_int, vp]
    lib.avl_builder_num_voxels.argtypes = [vp, C.POINTER(i64), vp]
    lib.avl_builder_num_accepted.argtypes = [vp, C.POINTER(i64), vp]
    lib.avl_builder_skip_frames.argtypes = [vp, i32]
    lib.avl_builder_reserve.argtypes = [vp, i64, vp]
    lib.avl_builder_h2d_bytes.argtypes = [vp, C.POINTER(i64)]
    lib.avl_builder_export.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, vp]
    lib.avl_builder_to_map.argtypes = [vp, vp, C.POINTER(vp)]
    lib.avl_builder_create_global.argtypes = [C.POINTER(GlobalGridSpec), C.POINTER(vp)]
    lib.avl_builder_num_rejected_oob.argtypes = [vp, C.POINTER(i64), vp]
    lib.avl_bounds_create.argtypes = [C.POINTER(vp)]
    lib.avl_bounds_destroy.argtypes = [vp]
    lib.avl_bounds_add_frame.argtypes = [vp, C.POINTER(Frame), C.c_int, vp]
    lib.avl_bounds_get.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64), vp]
    lib.avl_builder_set_slab.argtypes = [vp, i32, i32]
    lib.avl_builder_export_keys.argtypes = [vp, vp, C.c_int, vp]
    lib.avl_rank_keys.argtypes = [vp, vp, i32, i32, vp, C.c_int, vp]
    lib.avl_builder_import.argtypes = [vp, vp, vp, vp, vp, i64, C.c_int, vp]
    lib.avl_p2p_create.argtypes = [i32, i32, i32, i32, C.POINTER(vp)]
    lib.avl_p2p_handle_bytes.argtypes = []
    lib.avl_p2p_local_handle.argtypes = [vp, vp]
    lib.avl_p2p_connect.argtypes = [vp, vp]
    lib.avl_p2p_exchange_merge.argtypes = [vp, vp, vp, i32, i32, i64, vp, vp, vp, C.c_int, vp]
    lib.avl_p2p_status.argtypes = [vp, C.POINTER(i32), vp]
    lib.avl_p2p_destroy.argtypes = [vp]
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != AVL_OK:
        msg = load().avl_last_error().decode(errors="replace")
        raise AvlError(f"avlmaps_b200 error {rc}: {msg}")


def device_count() -> int:
    n = C.c_int(0)
    check(load().avl_device_count(C.byref(n)))
    return n.value


def require_device() -> None:
    if device_count() < 1:
        raise AvlError("no CUDA device: avlmaps_b200 has no CPU fallback")


def np_ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f32c(a, shape_last=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape_last is not None and a.shape[-1] != shape_last:
        raise ValueError(f"expected last dim {shape_last}, got {a.shape}")
    return a
