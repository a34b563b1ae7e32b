"""ctypes binding of include/das_decode.h (the C-ABI drop-in boundary).

There is deliberately no fallback: if the shared library is missing or the device is not
available every entry point raises -- the product path never routes through the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_LEVELS = 5
MAX_LAYERS = 4
MAX_JOINTS = 32
MAX_NMS_PRE = 2048
CAM_DOUBLES = 18

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_f64p = C.POINTER(C.c_double)


class LevelDesc(C.Structure):
    _fields_ = [("cls", C.c_void_p), ("ctr", C.c_void_p), ("pose", C.c_void_p),
                ("feats", C.c_void_p * MAX_LAYERS),
                ("H", C.c_int32), ("W", C.c_int32), ("stride", C.c_int32),
                ("scale_offset", C.c_float), ("scale_depth", C.c_float),
                ("scale_uv", C.c_float), ("scale_d", C.c_float)]


class Levels(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("batch", C.c_int32), ("in_dtype", C.c_int32), ("reserved_", C.c_int32),
                ("lv", LevelDesc * MAX_LEVELS)]


DTYPE_F32, DTYPE_F16, DTYPE_BF16 = 0, 1, 2      # das_levels.in_dtype (include/das_decode.h: DAS_DTYPE_*)


class DecodeCfg(C.Structure):
    _fields_ = [("num_joints", C.c_int32), ("root_idx", C.c_int32), ("num_heads", C.c_int32),
                ("feat_channels", C.c_int32), ("num_layers", C.c_int32),
                ("depth_factor", C.c_float), ("z_norm", C.c_float),
                ("nms_pre", C.c_int32), ("nms_post", C.c_int32),
                ("nms_thr", C.c_float), ("score_thr", C.c_float),
                ("peak_kernel", C.c_int32), ("refine", C.c_int32),
                ("dataset_depth_factor", C.c_double),
                ("nms_soft", C.c_int32), ("reserved_", C.c_int32)]


class Buffers(C.Structure):
    _fields_ = [("cand_score", C.c_void_p), ("cand_index", C.c_void_p),
                ("cand_pose", C.c_void_p), ("cand_center", C.c_void_p),
                ("out_count", C.c_void_p), ("out_score", C.c_void_p), ("out_slot", C.c_void_p),
                ("out_pose", C.c_void_p), ("out_center", C.c_void_p),
                ("out_cam", C.c_void_p), ("out_world", C.c_void_p)]


MAX_PEERS = 15


class PeerBlocks(C.Structure):
    _fields_ = [("n", C.c_int32), ("reserved_", C.c_int32), ("delta", C.c_int64 * MAX_PEERS),
                ("seq", C.c_void_p), ("ticket", C.c_void_p)]


class RefineScratch(C.Structure):
    _fields_ = [("unique_rows", C.c_void_p), ("unique_out", C.c_void_p), ("row_records", C.c_void_p),
                ("item_records", C.c_void_p), ("valid_list", C.c_void_p), ("counters", C.c_void_p),
                ("row_cap", C.c_int32), ("reserved_", C.c_int32), ("prev_planes", C.c_void_p)]


class RowCache(C.Structure):
    _fields_ = [("table", C.c_void_p), ("rows", C.c_void_p), ("cand_rows", C.c_void_p),
                ("table_bits", C.c_int32), ("max_rows", C.c_int32)]


# every symbol include/das_decode.h declares: (restype, argtypes)
_VP = C.c_void_p
SIGNATURES = {
    "das_version": (C.c_char_p, []),
    "das_abi_struct_sizes": (None, [_i32p]),
    "das_last_error": (C.c_char_p, []),
    "das_level_slots": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "das_candidate_slots": (C.c_int32, [C.POINTER(Levels), C.c_int32]),
    "das_output_slots": (C.c_int32, [C.c_int32, C.c_int32]),
    "das_score_topk": (C.c_int, [_VP, C.POINTER(Levels), C.c_int32, C.c_int32, _VP, _VP, C.c_int32, _VP, _VP]),
    "das_gather_refine_assemble": (C.c_int, [_VP, C.POINTER(Levels), C.POINTER(DecodeCfg), _VP, _VP, _VP, _VP, _VP,
                                             C.c_int32, _VP, _VP, _VP, _VP]),
    "das_refine_heads": (C.c_int, [_VP, C.POINTER(Levels), C.POINTER(DecodeCfg), _VP, _VP, _VP, _VP, _VP, C.c_int32,
                                   C.POINTER(RefineScratch), _VP, C.POINTER(RowCache), _VP]),
    "das_refine_tc": (C.c_int, [_VP, C.POINTER(Levels), C.POINTER(DecodeCfg), _VP, _VP, C.c_int32, C.POINTER(RefineScratch),
                                C.c_int32, _VP]),
    "das_refine_finish": (C.c_int, [C.POINTER(Levels), C.POINTER(DecodeCfg), C.c_int32, C.POINTER(RefineScratch), _VP, _VP]),
    "das_pack_tc_panels": (C.c_int, [C.POINTER(DecodeCfg), _VP, _VP, _VP]),
    "das_refine_row_cache": (C.c_int, [C.POINTER(DecodeCfg), C.POINTER(RefineScratch), C.POINTER(RowCache), _VP]),
    "das_row_cache_table_bytes": (C.c_int64, [C.c_int32]),
    "das_row_cache_clear": (C.c_int, [C.POINTER(RowCache), _VP]),
    "das_refine_cand_rows": (C.c_int, [_VP, C.POINTER(Levels), C.POINTER(DecodeCfg), _VP, _VP, C.c_int32,
                                       C.POINTER(RowCache), _VP]),
    "das_tc_set_debug_buffer": (C.c_int, [_VP]),
    "das_tc_panel_bytes": (C.c_int64, [C.POINTER(DecodeCfg)]),
    "das_plan_set_refine_mode": (C.c_int, [_VP, C.c_int32]),
    "das_plan_set_on_demand_sampling": (C.c_int, [_VP, C.c_int32]),
    "das_refine_dense_layer": (C.c_int, [_VP, C.POINTER(Levels), C.c_int32, C.c_int32, C.POINTER(DecodeCfg), _VP, _VP, _VP,
                                         _VP, _VP, _VP]),
    "das_dense_project_tc": (C.c_int, [_VP, C.POINTER(Levels), C.c_int32, C.c_int32, C.POINTER(DecodeCfg), _VP, _VP, _VP, _VP, _VP]),
    "das_pack_dense_panels": (C.c_int, [C.POINTER(DecodeCfg), _VP, _VP, _VP]),
    "das_dense_panel_bytes": (C.c_int64, [C.POINTER(DecodeCfg)]),
    "das_nms_backproject": (C.c_int, [C.POINTER(DecodeCfg), C.c_int32, C.c_int32, _VP, _VP, _VP, _VP, Buffers, _VP]),
    "das_nms_backproject_peers": (C.c_int, [C.POINTER(DecodeCfg), C.c_int32, C.c_int32, _VP, _VP, _VP, _VP, Buffers,
                                            C.POINTER(PeerBlocks), _VP]),
    "das_plan_set_output_block": (C.c_int, [_VP, _VP, C.c_int64]),
    "das_plan_set_peer_blocks": (C.c_int, [_VP, C.c_int32, C.POINTER(_VP)]),
    "das_ipc_alloc": (C.c_int, [C.c_int64, C.POINTER(_VP), C.c_char_p]),
    "das_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(_VP)]),
    "das_ipc_close": (C.c_int, [_VP]),
    "das_ipc_free": (C.c_int, [_VP]),
    "das_pack_weights": (C.c_int, [C.POINTER(DecodeCfg)] + [_VP] * 10),
    "das_packed_weight_floats": (C.c_int64, [C.POINTER(DecodeCfg)]),
    "das_plan_create": (C.c_int, [C.POINTER(DecodeCfg), C.POINTER(Levels), C.POINTER(_VP)]),
    "das_plan_destroy": (None, [_VP]),
    "das_plan_set_weights": (C.c_int, [_VP, C.c_int32] + [_VP] * 9),
    "das_plan_bind": (C.c_int, [_VP, C.POINTER(Levels), _VP]),
    "das_plan_set_metas": (C.c_int, [_VP, _VP, _VP, _VP]),
    "das_plan_run": (C.c_int, [_VP, _VP, C.c_int32]),
    "das_plan_stage_ms": (C.c_int, [_VP, _f32p]),
    "das_plan_output_block": (C.c_int, [_VP, C.POINTER(_VP), C.POINTER(C.c_int64)]),
    "das_plan_buffers": (C.c_int, [_VP, C.POINTER(Buffers), _i32p, _i32p]),
    "das_plan_kernel_launches": (C.c_int64, [_VP]),
    "das_plan_run_host": (C.c_int, [_VP, C.POINTER(Levels), _VP, _VP, Buffers, _VP]),
    "das_plan_run_host_async": (C.c_int, [_VP, C.POINTER(Levels), _VP, _VP, Buffers, _VP]),
    "das_tc_selftest": (C.c_int, [_VP, _VP, _VP, C.c_int32, C.c_int32, C.c_int32, _VP]),
    "das_tc_mma_bench": (C.c_int, [C.c_int32, C.c_int32, _VP, _VP]),
    "das_plan_h2d_bytes": (C.c_int64, [_VP]),
    "das_plan_set_host_mode": (C.c_int, [_VP, C.c_int32]),
    "das_plan_h2d_explicit_bytes": (C.c_int64, [_VP]),
    "das_plan_d2h_bytes": (C.c_int64, [_VP]),
    "das_plan_row_cache_stats": (C.c_int, [_VP, _i32p]),
    "das_plan_refine_stats": (C.c_int, [_VP, C.POINTER(C.c_int64)]),
    "das_plan_set_pdl": (C.c_int, [_VP, C.c_int32]),
    "das_peer_publish": (C.c_int, [_VP, _VP, C.c_int64, _VP]),
    "das_plan_publish_wait": (C.c_int, [_VP, _VP]),
    "das_debug_force_heads_kernel": (C.c_int, [C.c_int32]),
    "das_dense_set_debug_buffer": (C.c_int, [_VP]),
}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libdas_decode.so")
_lib = None


class DasError(RuntimeError):
    pass


def lib_path() -> str:
    return LIB_PATH


def load():
    """Load libdas_decode.so, (re)building it first when it is absent or older than its sources / the header and nvcc
    exists (a stale library would be bound with new ctypes struct layouts: silent corruption).  The library also
    reports the sizes of the by-value structs it was compiled with; a mismatch with the ctypes mirrors raises."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if _build._stale():
        if os.path.isfile(_build.NVCC):
            _build.build()
        elif not os.path.isfile(LIB_PATH):
            raise DasError(f"{LIB_PATH} is missing and nvcc ({_build.NVCC}) is not available to build it")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    sizes = (C.c_int32 * 6)()
    lib.das_abi_struct_sizes(sizes)
    mine = [C.sizeof(Levels), C.sizeof(DecodeCfg), C.sizeof(Buffers), C.sizeof(RowCache), C.sizeof(RefineScratch), C.sizeof(PeerBlocks)]
    if list(sizes) != mine:
        raise DasError(f"ABI mismatch: library struct sizes {list(sizes)} != ctypes mirrors {mine}; rebuild with "
                       "python -m das_b200.build --force")
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().das_last_error().decode(errors="replace")
        raise DasError(f"{what or 'das call'} failed with status {status}: {msg}")
