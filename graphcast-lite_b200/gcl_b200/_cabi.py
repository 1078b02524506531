"""ctypes binding of libgcl_b200.so (the C ABI declared in include/gcl_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, a RuntimeError is
raised (BASELINE.json north_star: "no CPU fallback").  The GIL is released during every call.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_LIB_NAME = "libgcl_b200.so"
_lib = None

P, I64, I32, F32, SZ = c_void_p, c_int64, c_int, c_float, c_size_t

# name -> (restype, argtypes); mirrors include/gcl_b200.h one to one.
_PROTOS = {
    "gcl_version": (c_int, []),
    "gcl_last_error": (c_char_p, []),
    "gcl_launch_count": (ctypes.c_longlong, []),
    "gcl_csr_workspace_bytes": (SZ, [I64, I64]),
    "gcl_csr_build": (c_int, [P, P, I64, I64, I32, P, P, P, P, P, P, P, P, P, P, P, SZ, P]),
    "gcl_csr_weights": (c_int, [P, P, P, P, P, I64, I64, I32, P, P, P, P]),
    "gcl_spmm_f32": (c_int, [P, P, P, P, P, I64, I64, I64, I64, I64, I64, P, P, P, I64, P]),
    "gcl_tile_plan_host": (c_int, [P, P, P, I64, I64, I64, I32, I32, I32, I32, P, P, P, P, P, P, P, P, P, P]),
    "gcl_spmm_tiled_f32": (c_int, [P, P, P, P, P, P, P, I64, I64, I64, I64, I64, P, P, P, P]),
    "gcl_spmm_tiled_bf16": (c_int, [P, P, P, P, P, P, P, I64, I64, I64, I64, I64, P, P, P, P]),
    "gcl_gat_fwd_tiled_f32": (c_int, [P] * 11 + [I64, I64, I64, I64, F32, P]),
    "gcl_gat_bwd_tiled_f32": (c_int, [P] * 14 + [I64, I64, I64, I64, F32, P]),
    "gcl_gat_ws_supported": (c_int, [P, P, I64, I64]),
    "gcl_gat_fwd_ws_f32": (c_int, [P] * 16 + [I64, I64, I64, I64, I64, I64, F32, P]),
    "gcl_gat_bwd_ws_f32": (c_int, [P] * 16 + [I64, I64, I64, I64, I64, P]),
    "gcl_linear_fwd_f32": (c_int, [P, P, P, P, I64, I64, I64, P, P, P, P]),
    "gcl_linear_fwd_scores_f32": (c_int, [P, P, P, P, P, P, P, I64, I64, I64, P, P]),
    "gcl_linear_bwd_dx_f32": (c_int, [P, P, P, I64, I64, I64, P, P]),
    "gcl_linear_bwd_dx_prelu_workspace_bytes": (c_size_t, [I64, I64]),
    "gcl_linear_bwd_dx_prelu_f32": (c_int, [P, P, P, P, P, P, P, I64, I64, I64, P, P, c_size_t, P]),
    "gcl_set_dense_mode": (c_int, [I32]),
    "gcl_get_dense_mode": (c_int, []),
    "gcl_linear_bwd_dw_workspace_bytes": (SZ, [I64, I64, I64]),
    "gcl_linear_bwd_dw_f32": (c_int, [P, P, P, P, I64, I64, I64, P, SZ, P]),
    "gcl_colsum_workspace_bytes": (SZ, [I64, I64]),
    "gcl_colsum_f32": (c_int, [P, P, I64, I64, P, SZ, P]),
    "gcl_prelu_fwd_f32": (c_int, [P, P, P, I64, P]),
    "gcl_prelu_bwd_workspace_bytes": (SZ, [I64]),
    "gcl_prelu_bwd_f32": (c_int, [P, P, P, P, P, I64, P, SZ, P]),
    "gcl_prelu_bwd_colsum_workspace_bytes": (SZ, [I64, I64]),
    "gcl_prelu_bwd_colsum_f32": (c_int, [P, P, P, P, P, P, I64, I64, P, SZ, P]),
    "gcl_layernorm_fwd_f32": (c_int, [P, P, P, P, P, P, I64, I64, F32, P]),
    "gcl_layernorm_bwd_workspace_bytes": (SZ, [I64, I64]),
    "gcl_layernorm_bwd_f32": (c_int, [P, P, P, P, P, P, P, P, I64, I64, P, SZ, P]),
    "gcl_gat_scores_f32": (c_int, [P, P, P, P, P, I64, I64, I64, P]),
    "gcl_gat_fwd_f32": (c_int, [P] * 12 + [I64, I64, I64, I64, I64, I32, F32, P]),
    "gcl_gat_bwd_f32": (c_int, [P] * 16 + [I64, I64, I64, I64, I64, I32, F32, P]),
    "gcl_gat_datt_workspace_bytes": (SZ, [I64, I64, I64]),
    "gcl_gat_datt_f32": (c_int, [P, P, P, P, P, I64, I64, I64, P, SZ, P]),
    "gcl_edge_prune_workspace_bytes": (SZ, [I64]),
    "gcl_edge_prune": (c_int, [P, P, I64, I64, F32, P, I64, P, P, SZ, P]),
    "gcl_radius_query_workspace_bytes": (SZ, [I64]),
    "gcl_radius_query_count": (c_int, [P, P, I64, I64, ctypes.c_double, P, P, SZ, P]),
    "gcl_radius_query_fill": (c_int, [P, P, I64, I64, ctypes.c_double, P, P, I64, I64, P]),
    "gcl_closest_face": (c_int, [P, P, P, I64, I64, I64, ctypes.c_double, P, P]),
    "gcl_assemble_input_f32": (c_int, [P, P, P, P, I64, I64, I64, I64, I64, I64, P]),
    "gcl_rows_concat_f32": (c_int, [P, P, P, I64, I64, I64, I64, P]),
    "gcl_rows_split_f32": (c_int, [P, P, P, I64, I64, I64, I64, P]),
    "gcl_rows_block_copy_f32": (c_int, [P, P, I64, I64, I64, I64, I64, I64, I64, P]),
    "gcl_wmse_workspace_bytes": (SZ, [I64, I64, I64]),
    "gcl_wmse_f32": (c_int, [P, P, I64, P, I64, P, F32, P, P, P, I32, F32, I64, I64, I64, P, SZ, P]),
    "gcl_ar_step_f32": (c_int, [P, P, P, I64, P, P, P, I32, F32, F32, P, P, P, I32, I64, I64, I64, I64, P, SZ, P]),
    "gcl_ar_step_bwd_f32": (c_int, [P, P, P, P, I32, P, P, I64, I64, I64, I64, P]),
    "gcl_window_assemble": (c_int, [P, I32, P, P, P, P, I64, I64, I64, I64, I64, I64, I64, I32, P]),
    "gcl_forecast_metrics_workspace_bytes": (SZ, [I64, I64]),
    "gcl_forecast_metrics_f32": (c_int, [P, P, P, I64, I64, I64, P, SZ, P]),
    "gcl_resize_channels_f32": (c_int, [P, P, I64, I64, I64, P]),
    "gcl_act_fwd_f32": (c_int, [P, P, I64, I32, P]),
    "gcl_act_bwd_f32": (c_int, [P, P, P, I64, I32, P]),
    "gcl_add_f32": (c_int, [P, P, P, I64, P]),
    "gcl_layernorm_graph_workspace_bytes": (SZ, [I64]),
    "gcl_layernorm_graph_fwd_f32": (c_int, [P, P, P, P, P, I64, I64, I64, F32, P, SZ, P]),
    "gcl_layernorm_graph_bwd_f32": (c_int, [P, P, P, P, P, P, P, I64, I64, I64, P, SZ, P]),
    "gcl_adam_f32": (c_int, [P, P, P, P, I64, F32, F32, F32, F32, F32, P, P]),
}

ABI_VERSION = 4


class TilePlanStruct(ctypes.Structure):
    """gcl_tile_plan of include/gcl_b200.h: device arrays of one tile plan (host struct, passed by pointer)."""
    _fields_ = [("tile_rowptr", c_void_p), ("tile_uptr", c_void_p), ("rows", c_void_p), ("eptr", c_void_p),
                ("lidx", c_void_p), ("ek", c_void_p), ("usrc", c_void_p), ("heavy_rows", c_void_p),
                ("tile_desc", c_void_p),
                ("n_tiles", ctypes.c_int32), ("n_heavy", ctypes.c_int32), ("max_rows", ctypes.c_int32),
                ("max_union", ctypes.c_int32), ("max_entries", ctypes.c_int32), ("pad_entries", ctypes.c_int32)]


def lib_path() -> str:
    override = os.environ.get("GCL_LIB_PATH")          # development: A/B builds of the kernels
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load():
    """Load (once) and return the CDLL.  Raises RuntimeError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"gcl_b200: {path} not found -- build it with graphcast-lite_b200/csrc/build.sh "
            "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.gcl_version() != ABI_VERSION:
        raise RuntimeError(f"gcl_b200: ABI version {lib.gcl_version()} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def exported_names():
    return sorted(_PROTOS)


def last_error() -> str:
    return load().gcl_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")
