"""ctypes binding of libwflow_b200.so (include/wflow_b200.h). Fails loudly when the CUDA
library is missing: there is no CPU fallback on the product path."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libwflow_b200.so")
_lib = None


class Config(C.Structure):
    _fields_ = [("n", C.c_int64), ("nriv", C.c_int64), ("n_layers", C.c_int32),
                ("device", C.c_int32), ("gash", C.c_int32), ("has_lai", C.c_int32),
                ("snow", C.c_int32), ("glacier", C.c_int32),
                ("soil_infiltration_reduction", C.c_int32), ("kv_profile", C.c_int32),
                ("adaptive", C.c_int32), ("nthreads", C.c_int32),
                ("land_streamorder_min", C.c_int32), ("river_streamorder_min", C.c_int32),
                ("dt_land", C.c_double), ("dt_river", C.c_double), ("dt_ssf", C.c_double),
                ("ssf_alpha_coefficient", C.c_double), ("kin_wave_min_flow_qroot", C.c_double),
                ("wave_piece_depth_land", C.c_int32), ("vertical_slices", C.c_int32),
                ("unsat_inline_iters", C.c_int32), ("snow_gravitational_transport", C.c_int32),
                ("river_routing", C.c_int32), ("li_froude_limit", C.c_int32),
                ("li_ghost_nodes", C.c_int32), ("reserved_", C.c_int32),
                ("li_alpha", C.c_double), ("li_h_thresh", C.c_double),
                ("fp_levels", C.c_int32), ("reserved2_", C.c_int32), ("fp_depth", C.c_double * 16),
                ("land_routing", C.c_int32), ("li_land_froude_limit", C.c_int32),
                ("li_land_alpha", C.c_double), ("li_land_theta", C.c_double),
                ("li_land_h_thresh", C.c_double)]


class Domain(C.Structure):
    _fields_ = [("d1", C.c_int64), ("d2", C.c_int64), ("indices", C.c_void_p),
                ("ldd", C.c_void_p), ("river_land_indices", C.c_void_p), ("nres", C.c_int64),
                ("reservoir_river_indices", C.c_void_p),
                # cut edges of a shard that is part of a basin
                ("n_land_imports", C.c_int64), ("land_import_dst", C.c_void_p),
                ("land_import_pos", C.c_void_p), ("n_land_exports", C.c_int64),
                ("land_export_src", C.c_void_p), ("n_river_imports", C.c_int64),
                ("river_import_dst", C.c_void_p), ("river_import_pos", C.c_void_p),
                ("n_river_exports", C.c_int64), ("river_export_src", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in (
        "newton_calls_land", "newton_iters_land", "newton_maxit_land", "newton_calls_river",
        "newton_iters_river", "newton_maxit_river", "substeps_land", "substeps_river",
        "substeps_ssf", "wave_levels_land", "wave_levels_river", "kernel_launches")] + [
        (k, C.c_double) for k in ("ms_land_hydrology", "ms_subsurface", "ms_soil_storage",
                                  "ms_overland", "ms_river", "ms_total_storage", "ms_glue")] + [
        ("timed_steps", C.c_int64)]


ARTIFACTS = dict(order=0, streamorder=1, upstream_ptr=2, upstream_idx=3, subdomain_level_ptr=4,
                 subdomain_level_idx=5, subdomain_ptr=6, subdomain_order=7, subdomain_indices=8,
                 ldd=9, wave_level_ptr=10, wave_perm=11, wave_node_level=12, wave_chunk_ptr=13,
                 wave_chunk_outlet=14)
# EdgeConnectivity of the land network (land_routing = 1 only)
EDGE_ARTIFACTS = dict(edge_x_up=15, edge_x_down=16, edge_y_up=17, edge_y_down=18)


def header_symbols():
    """Every entry point declared in include/wflow_b200.h."""
    src = open(os.path.join(_ROOT, "include", "wflow_b200.h")).read()
    return sorted(set(re.findall(r"\b(wflowb200_[a-z0-9_]+)\s*\(", src)))


def build(verbose: bool = False) -> str:
    """Compile libwflow_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j4"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libwflow_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
            "(wflow.jl_b200 has no CPU fallback)")
    L = C.CDLL(os.environ.get("WFB_LIB", LIB_PATH))  # WFB_LIB: a build variant (scripts/build_variants.sh)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.wflowb200_create.argtypes = [C.POINTER(Config), C.POINTER(Domain), C.POINTER(vp)]
    L.wflowb200_destroy.argtypes = [vp]
    L.wflowb200_destroy.restype = None
    L.wflowb200_last_error.argtypes = [vp]
    L.wflowb200_last_error.restype = C.c_char_p
    L.wflowb200_field_name.argtypes = [i32]
    L.wflowb200_field_name.restype = C.c_char_p
    L.wflowb200_field_kind.argtypes = [i32]
    L.wflowb200_field_id.argtypes = [C.c_char_p]
    L.wflowb200_set_field.argtypes = [vp, i32, vp, i64, i64]
    L.wflowb200_get_field.argtypes = [vp, i32, vp, i64, i64]
    L.wflowb200_set_field_i64.argtypes = [vp, i32, vp]
    L.wflowb200_get_field_i64.argtypes = [vp, i32, vp]
    L.wflowb200_set_forcing.argtypes = [vp, vp, vp, vp]
    for f in ("update_land_hydrology_model", "update_subsurface_flow_model",
              "update_soil_water_storage", "update_overland_flow_model",
              "update_river_flow_model", "update_model"):
        getattr(L, "wflowb200_" + f).argtypes = [vp, dbl]
    for f in ("exchange_recharge", "update_lateral_inflow_overland", "update_lateral_inflow_river",
              "update_inflow_reservoir", "update_total_water_storage", "synchronize",
              "update_bc_overland_flow_model"):
        getattr(L, "wflowb200_" + f).argtypes = [vp]
    L.wflowb200_get_artifact.argtypes = [vp, i32, i32, vp, i64, C.POINTER(i64)]
    L.wflowb200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.wflowb200_set_timing.argtypes = [vp, i32]
    L.wflowb200_timer_start.argtypes = [vp]
    L.wflowb200_timer_stop.argtypes = [vp, C.POINTER(C.c_double)]
    L.wflowb200_forcing_ring_create.argtypes = [vp, i32]
    L.wflowb200_forcing_ring_put.argtypes = [vp, i32, vp, vp, vp]
    L.wflowb200_forcing_ring_use.argtypes = [vp, i32]
    L.wflowb200_set_cyclic_lai.argtypes = [vp, vp, i32]
    L.wflowb200_use_cyclic_lai.argtypes = [vp, i32]
    L.wflowb200_get_fields.argtypes = [vp, vp, i32, vp]
    L.wflowb200_get_fields_async.argtypes = [vp, vp, i32, vp]
    L.wflowb200_wait_outputs.argtypes = [vp]
    L.wflowb200_comm_unique_id.argtypes = [C.c_char_p]
    L.wflowb200_comm_init_nccl.argtypes = [vp, i32, i32, C.c_char_p]
    L.wflowb200_group_create.argtypes = [i32, C.POINTER(vp)]
    L.wflowb200_group_join.argtypes = [vp, vp]
    L.wflowb200_exchange_prepare.argtypes = [vp, C.c_double, C.POINTER(C.c_uint64), vp,
                                             C.POINTER(C.c_int64)]
    L.wflowb200_exchange_open_peer.argtypes = [vp, i32, C.c_uint64, i32, vp, C.c_int64, C.c_int64]
    L.wflowb200_exchange_bind.argtypes = [vp, i32, C.c_int64, i32, C.c_int64]
    L.wflowb200_group_destroy.argtypes = [vp]
    L.wflowb200_group_destroy.restype = None
    L.wflowb200_set_option.argtypes = [vp, C.c_char_p, i32]
    L.wflowb200_get_vertical_timeline.argtypes = [vp, C.POINTER(C.c_double), i32]
    L.wflowb200_get_unsat_buckets.argtypes = [vp, C.POINTER(C.c_int64), i32]
    L.wflowb200_newton_trace.argtypes = [vp, i32]
    L.wflowb200_get_newton_trace.argtypes = [vp, i32, vp]
    L.wflowb200_selftest_math.argtypes = [i32, i64, C.POINTER(C.c_double)]
    _lib = L
    return L


def field_table():
    L = lib()
    return [(L.wflowb200_field_name(i).decode(), L.wflowb200_field_kind(i))
            for i in range(L.wflowb200_num_fields())]
