"""
oracle.py -- ctypes front-end of the CPU ORACLE (oracle/wfo*.c + oracle/network.py).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs. The product (wflow.jl_b200/) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libwfo.so")
_ALT_PATH = os.path.join(_HERE, "_build", "libwfo_alt.so")  # noisy-libm variant (wfo_math.h)
_libs = {}


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("wfo_vertical.c", "wfo_routing.c", "wfo.h", "wfo_math.h")]
    stale = any((not os.path.exists(p)) or any(os.path.getmtime(s) > os.path.getmtime(p)
                                                for s in srcs) for p in (_LIB_PATH, _ALT_PATH))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "all"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Net(C.Structure):
    _fields_ = [("n", C.c_int64), ("up_ptr", C.c_void_p), ("up_idx", C.c_void_p),
                ("n_levels", C.c_int64), ("level_ptr", C.c_void_p), ("level_sub", C.c_void_p),
                ("n_sub", C.c_int64), ("sub_ptr", C.c_void_p), ("sub_nodes", C.c_void_p),
                ("sub_pos", C.c_void_p), ("down", C.c_void_p), ("order", C.c_void_p),
                ("in_ptr", C.c_void_p), ("in_idx", C.c_void_p)]


class _Cfg(C.Structure):
    _fields_ = [("n", C.c_int64), ("nriv", C.c_int64), ("N", C.c_int64), ("nres", C.c_int64),
                ("gash", C.c_int32), ("has_lai", C.c_int32), ("snow", C.c_int32),
                ("glacier", C.c_int32), ("soil_infiltration_reduction", C.c_int32),
                ("kv_profile", C.c_int32), ("adaptive", C.c_int32), ("snow_transport", C.c_int32),
                ("river_routing", C.c_int32), ("li_froude_limit", C.c_int32),
                ("li_ghost_nodes", C.c_int32), ("li_alpha", C.c_double), ("li_h_thresh", C.c_double),
                ("nthreads", C.c_int32),
                ("dt_land", C.c_double), ("dt_river", C.c_double), ("dt_ssf", C.c_double),
                ("ssf_alpha_coefficient", C.c_double), ("fp_levels", C.c_int32),
                ("fp_depth", C.c_double * 16),
                ("land_routing", C.c_int32), ("li_land_froude_limit", C.c_int32),
                ("li_land_alpha", C.c_double), ("li_land_theta", C.c_double),
                ("li_land_h_thresh", C.c_double)]


def lib(variant: str = ""):
    """variant "alt": the same oracle on a noisy libm (tolerance calibration)."""
    if variant not in _libs:
        build()
        L = C.CDLL(_ALT_PATH if variant == "alt" else _LIB_PATH)
        L.wfo_field_name.restype = C.c_char_p
        L.wfo_new.restype = C.c_void_p
        L.wfo_cfg.restype = C.POINTER(_Cfg)
        L.wfo_cfg.argtypes = [C.c_void_p]
        L.wfo_free.argtypes = [C.c_void_p]
        L.wfo_set_ptr.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.wfo_set_iptr.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.wfo_set_network.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Net)]
        for f in ("wfo_update_land_hydrology_model", "wfo_update_subsurface_flow_model",
                  "wfo_update_soil_water_storage", "wfo_surface_routing",
                  "wfo_update_overland_flow_model", "wfo_update_river_flow_model",
                  "wfo_update_model"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_double]
            getattr(L, f).restype = None
        for f in ("wfo_exchange_recharge", "wfo_update_total_water_storage",
                  "wfo_update_diagnostic_vars", "wfo_update_lateral_inflow_overland",
                  "wfo_update_lateral_inflow_river", "wfo_update_inflow_reservoir"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = None
        L.wfo_kinwave_river_update.argtypes = [C.c_void_p, C.c_double]
        L.wfo_kinwave_river_update.restype = None
        L.wfo_li_stable_timestep.argtypes = [C.c_void_p]
        L.wfo_li_stable_timestep.restype = C.c_double
        for f in ("wfo_li_update_river_channel_flow", "wfo_li_update_bc_reservoir_model",
                  "wfo_li_update_water_depth_and_storage", "wfo_li_update_floodplain_flow",
                  "wfo_li_update_floodplain_water_depth_and_storage",
                  "wfo_river_channel_floodplain_exchange", "wfo_update_floodplain_model"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_double]
            getattr(L, f).restype = None
        # 2-D local-inertial overland flow (test hooks of the reference's unit tests)
        L.wfo_lil_stable_timestep.argtypes = [C.c_void_p]
        L.wfo_lil_stable_timestep.restype = C.c_double
        L.wfo_lil_update_directional_flow.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_int]
        L.wfo_lil_update_directional_flow.restype = None
        for f in ("wfo_lil_compute_river_storage_change", "wfo_lil_compute_land_storage_change"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int64, C.c_double]
            getattr(L, f).restype = C.c_double
        L.wfo_lil_compute_water_depths.argtypes = [C.c_void_p, C.c_double, C.c_int64, C.c_int64,
                                                   C.POINTER(C.c_double)]
        L.wfo_lil_compute_water_depths.restype = None
        for f in ("wfo_lil_update_river_and_land_storage_and_depth",
                  "wfo_lil_update_land_storage_and_depth"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int64, C.c_double]
            getattr(L, f).restype = None
        for f in ("wfo_lil_update_fluxes", "wfo_lil_update_water_depth",
                  "wfo_lil_update_overland_flow_model"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_double]
            getattr(L, f).restype = None
        for f in ("wfo_lil_update_inflow_reservoir", "wfo_update_bc_overland_flow_model"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = None
        L.wfo_update_reservoir_model.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double]
        L.wfo_update_reservoir_model.restype = None
        L.wfo_update_reservoir_at_node.argtypes = [C.c_void_p, C.c_int64, C.c_double]
        L.wfo_update_reservoir_at_node.restype = None
        L.wfo_local_inertial_flow.argtypes = [C.c_double] * 8 + [C.c_int, C.c_double]
        L.wfo_local_inertial_flow.restype = C.c_double
        L.wfo_local_inertial_flow_rect.argtypes = [C.c_double] * 10 + [C.c_int, C.c_double]
        L.wfo_local_inertial_flow_rect.restype = C.c_double
        L.wfo_get_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.wfo_get_stats.restype = None
        L.wfo_set_num_threads.argtypes = [C.c_int]
        L.wfo_set_num_threads.restype = C.c_int
        L.wfo_sweep.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.wfo_sweep.restype = C.c_int
        d = C.c_double
        pd = C.POINTER(C.c_double)

        def sig(name, args, res=None):
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res

        sig("wfo_rainfall_interception_gash", [d] * 7 + [pd])
        sig("wfo_rainfall_interception_modrut", [d] * 6 + [pd])
        sig("wfo_precipitation_hbv", [d] * 4 + [pd])
        sig("wfo_snowpack_hbv", [d] * 9 + [pd])
        sig("wfo_glacier_hbv", [d] * 9 + [pd])
        sig("wfo_infiltration", [d] * 7 + [pd])
        sig("wfo_unsatzone_flow_layer", [d] * 5 + [pd])
        sig("wfo_vwc_brooks_corey", [d] * 5, d)
        sig("wfo_head_brooks_corey", [d] * 5, d)
        sig("wfo_feddes_h3", [d] * 3, d)
        sig("wfo_rwu_reduction_feddes", [d] * 6, d)
        sig("wfo_soil_temperature", [d] * 3, d)
        sig("wfo_infiltration_reduction_factor", [d, d, C.c_int, C.c_int], d)
        sig("wfo_soil_evaporation_unsaturated_store", [d, d, d, C.c_int64, d, d], d)
        sig("wfo_soil_evaporation_saturated_store", [d, C.c_int64, d, d, d, d], d)
        sig("wfo_actual_infiltration_soil_path", [d] * 6 + [pd])
        sig("wfo_scurve", [d] * 4, d)
        sig("wfo_kinematic_wave", [d] * 6 + [pd, C.POINTER(C.c_int64)])
        sig("wfo_kw_ssf_newton_raphson", [d] * 5, d)
        sig("wfo_ssf_celerity", [d] * 6 + [C.c_int], d)
        sig("wfo_kinematic_wave_ssf", [C.c_void_p] + [d] * 11 + [C.c_int64, pd])
        sig("wfo_water_table_change", [C.c_void_p, d, d, C.c_int64, d, pd])
        sig("wfo_stable_timestep_surface", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, d,
                                            C.c_void_p], d)
        sig("wfo_round_sigdigits12", [d], d)
        sig("wfo_cld", [d, d], d)
        sig("wfo_accucapacityflux", [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, d])
        _libs[variant] = L
    return _libs[variant]


def field_table():
    L = lib()
    return [(L.wfo_field_name(i).decode(), L.wfo_field_kind(i)) for i in range(L.wfo_num_fields())]


# Non-NaN defaults of the reference's structs (canopy.jl:11, runoff.jl:26, soil.jl:71-83,
# lateral_subsurface_flow.jl:9-38, boundary_conditions.jl:204-213, surface_kinwave.jl:5-29,
# 154-185, surface_flow.jl:9-34); every other field starts as MISSING_VALUE (NaN).
ZERO_DEFAULTS = (
    "li_error", "li_zs_at_edge", "li_water_depth_at_edge", "snow_in", "snow_out", "canopy_storage", "waterdepth_river", "unsaturated_store_depth", "total_storage",
    "ssf_exfiltwater_cumulative", "ssf_exfiltwater_average", "ssf_q_cumulative", "ssf_q_average",
    "ssf_q_in_cumulative", "ssf_q_in_average", "ssf_to_river_cumulative", "ssf_to_river_average",
    "ssf_q_net_cumulative", "ssf_q_net_average", "recharge_flux", "recharge_flux_cumulative",
    "olf_inwater", "olf_q", "olf_qlat", "olf_qin", "olf_qin_cumulative", "olf_qin_average",
    "olf_q_cumulative", "olf_q_average", "olf_storage", "olf_h", "olf_to_river_cumulative",
    "olf_to_river_average", "riv_external_inflow", "riv_abstraction",
    "riv_actual_external_abstraction_cumulative", "riv_actual_external_abstraction_average",
    "riv_inwater", "riv_q", "riv_qlat", "riv_qin", "riv_qin_cumulative", "riv_qin_average",
    "riv_q_cumulative", "riv_q_average", "riv_storage", "riv_h")
VALUE_DEFAULTS = {"f_infiltration_reduction": 1.0, "soil_surface_temperature": 10.0 + 273.15}

INT_FIELDS = ("number_of_layers", "n_unsatlayers", "nlayers_kv", "river_land_indices",
              "reservoir_river_indices", "edge_x_up", "edge_x_down", "edge_y_up", "edge_y_down",
              "land_river_indices")
# reservoir defaults (reservoir.jl:200-272): cumulative / average variables start at zero
ZERO_DEFAULTS = ZERO_DEFAULTS + (
    "fp_h", "fp_storage", "fp_q", "fp_q_cumulative", "fp_q_average", "fp_error",
    "fp_water_depth_at_edge", "riv_q_channel_average", "fp_flow_capacity", "fp_qin",
    "fp_qin_cumulative", "fp_qin_average", "riv_floodplain_water_exchange",
    "res_inflow_cumulative", "res_inflow_average", "res_external_inflow",
    "res_actual_external_abstraction_cumulative", "res_actual_external_abstraction_average",
    "res_outflow_cumulative", "res_outflow_average", "res_actevap_cumulative",
    # LocalInertialOverlandFlowVariables / BC (surface_staggered_scheme.jl:840-865,963-968)
    "li_land_runoff", "li_land_qx0", "li_land_qy0", "li_land_qx", "li_land_qy",
    "li_land_qx_cumulative", "li_land_qy_cumulative", "li_land_qx_average", "li_land_qy_average",
    "li_land_error")


def call_out(fn_name, nout, *args):
    """Call a scalar oracle kernel that writes `nout` doubles to its trailing out[] argument."""
    out = (C.c_double * nout)()
    getattr(lib(), fn_name)(*args, out)
    return tuple(out)


def _csr(lists):
    ptr = np.zeros(len(lists) + 1, dtype=np.int64)
    for k, l in enumerate(lists):
        ptr[k + 1] = ptr[k] + len(l)
    idx = np.concatenate([np.asarray(l, dtype=np.int64) for l in lists]) if lists else np.zeros(0, np.int64)
    return ptr, np.ascontiguousarray(idx, dtype=np.int64)


class OracleModel:
    """Owns numpy arrays for every field and walks them with the C oracle.

    `fields`: dict name -> ndarray (land scalars (n,), layered (n, N) cell-major like Julia's
    Vector{SVector{N}}, river (nriv,)); missing fields are allocated as NaN (MISSING_VALUE).
    `net_land` / `net_river`: dicts from oracle.network.build_domain_network (1-based).
    """

    def __init__(self, cfg: dict, fields: dict, net_land: dict, net_river: dict,
                 variant: str = ""):
        L = lib(variant)
        self._L = L
        self.h = L.wfo_new()
        c = L.wfo_cfg(self.h).contents
        n, nriv, N = int(cfg["n"]), int(cfg["nriv"]), int(cfg["N"])
        nres = int(cfg.get("nres", 0))
        c.n, c.nriv, c.N, c.nres = n, nriv, N, nres
        for k in ("gash", "has_lai", "snow", "glacier", "soil_infiltration_reduction",
                  "kv_profile", "adaptive", "snow_transport", "river_routing", "li_froude_limit",
                  "li_ghost_nodes"):
            setattr(c, k, int(cfg.get(k, 0)))
        c.li_alpha = float(cfg.get("li_alpha", 0.7))
        c.li_h_thresh = float(cfg.get("li_h_thresh", 1.0e-3))
        c.land_routing = int(cfg.get("land_routing", 0))
        c.li_land_froude_limit = int(cfg.get("li_land_froude_limit", 1))
        c.li_land_alpha = float(cfg.get("li_land_alpha", 0.7))
        c.li_land_theta = float(cfg.get("li_land_theta", 1.0))
        c.li_land_h_thresh = float(cfg.get("li_land_h_thresh", 1.0e-3))
        c.nthreads = int(cfg.get("nthreads", 0))
        c.dt_land = float(cfg.get("dt_land", 3600.0))
        c.dt_river = float(cfg.get("dt_river", 900.0))
        c.dt_ssf = float(cfg.get("dt_ssf", 86400.0))
        c.ssf_alpha_coefficient = float(cfg.get("ssf_alpha_coefficient", 1.0))
        fp_depth = [float(x) for x in cfg.get("fp_depth", [])]
        c.fp_levels = len(fp_depth)
        for k, x in enumerate(fp_depth):
            c.fp_depth[k] = x
        P = max(len(fp_depth), 1)
        self.cfg = dict(cfg)
        self.f = {}
        n6 = n if int(cfg.get("land_routing", 0)) == 1 else 0
        shapes = {0: (n,), 1: (n, N), 2: (n, N + 1), 3: (nriv,), 4: (nres,), 5: (nriv, P), 6: (n6,)}
        for name, kind in field_table():
            if name in fields and fields[name] is not None:
                a = np.ascontiguousarray(np.array(fields[name], dtype=np.float64, copy=True))
                assert a.shape == shapes[kind], (name, a.shape, shapes[kind])
            elif name in ZERO_DEFAULTS:
                a = np.zeros(shapes[kind])
            elif name in VALUE_DEFAULTS:
                a = np.full(shapes[kind], VALUE_DEFAULTS[name])
            else:
                a = np.full(shapes[kind], np.nan)
            self.f[name] = a
            L.wfo_set_ptr(self.h, name.encode(), a.ctypes.data)
        for name in INT_FIELDS:
            size = {"river_land_indices": nriv, "reservoir_river_indices": nres}.get(name, n)
            if name.startswith("edge_") or name == "land_river_indices":
                if name not in fields:      # only the 2-D local-inertial overland flow reads them
                    continue
            if name in fields and fields[name] is not None:
                a = np.ascontiguousarray(np.array(fields[name], dtype=np.int64, copy=True))
            else:
                a = np.zeros(size, dtype=np.int64)
            self.f[name] = a
            L.wfo_set_iptr(self.h, name.encode(), a.ctypes.data)
        self._keep = []
        for which, net in ((0, net_land), (1, net_river)):
            s = _Net()
            s.n = len(net["order"])
            arrs = {}
            arrs["up_ptr"] = np.ascontiguousarray(net["up_ptr"], dtype=np.int64)
            arrs["up_idx"] = np.ascontiguousarray(net["up_idx"] - 1, dtype=np.int64)
            lp, ls = _csr(net["order_of_subdomains"])
            arrs["level_ptr"], arrs["level_sub"] = lp, ls - 1
            sp, sn = _csr(net["order_subdomain"])
            _, si = _csr(net["subdomain_indices"])
            arrs["sub_ptr"], arrs["sub_nodes"], arrs["sub_pos"] = sp, sn - 1, si - 1
            if "graph" in net:   # only the reservoirs need the downstream node
                arrs["down"] = np.asarray(net["graph"].down, dtype=np.int64) - 1   # 0 (pit) -> -1
            else:
                arrs["down"] = np.full(max(int(s.n), 1), -1, dtype=np.int64)
            arrs["order"] = np.asarray(net["order"], dtype=np.int64) - 1
            dn = arrs["down"][:int(s.n)]            # in-neighbours by node id (full graph)
            src = np.nonzero(dn >= 0)[0]
            o = np.argsort(dn[src], kind="stable")
            arrs["in_ptr"] = np.concatenate([[0], np.cumsum(np.bincount(dn[src], minlength=int(s.n)))])
            arrs["in_idx"] = src[o]
            s.n_levels = len(net["order_of_subdomains"])
            s.n_sub = len(net["order_subdomain"])
            for k, a in arrs.items():
                a = np.ascontiguousarray(a, dtype=np.int64)
                arrs[k] = a
                setattr(s, k, a.ctypes.data)
            self._keep.append((s, arrs))
            L.wfo_set_network(self.h, which, C.byref(s))

    def __del__(self):
        try:
            self._L.wfo_free(self.h)
        except Exception:
            pass

    # names mirror the reference's update functions
    def update_land_hydrology_model(self, dt): self._L.wfo_update_land_hydrology_model(self.h, dt)
    def exchange_recharge(self): self._L.wfo_exchange_recharge(self.h)
    def update_subsurface_flow_model(self, dt): self._L.wfo_update_subsurface_flow_model(self.h, dt)
    def update_soil_water_storage(self, dt): self._L.wfo_update_soil_water_storage(self.h, dt)
    def surface_routing(self, dt): self._L.wfo_surface_routing(self.h, dt)
    def update_lateral_inflow_overland(self): self._L.wfo_update_lateral_inflow_overland(self.h)
    def update_lateral_inflow_river(self): self._L.wfo_update_lateral_inflow_river(self.h)
    def update_inflow_reservoir(self): self._L.wfo_update_inflow_reservoir(self.h)
    def update_overland_flow_model(self, dt): self._L.wfo_update_overland_flow_model(self.h, dt)
    def update_river_flow_model(self, dt): self._L.wfo_update_river_flow_model(self.h, dt)
    def update_total_water_storage(self): self._L.wfo_update_total_water_storage(self.h)
    def update_diagnostic_vars(self): self._L.wfo_update_diagnostic_vars(self.h)
    def update_model(self, dt): self._L.wfo_update_model(self.h, dt)

    def newton_stats(self):
        """Counters kept by wfo_routing.c (the struct tail after the network blocks)."""
        out = (C.c_int64 * 9)()
        self._L.wfo_get_stats(self.h, out)
        keys = ("newton_iters_land", "newton_iters_river", "newton_calls_land",
                "newton_calls_river", "newton_maxit_land", "newton_maxit_river",
                "substeps_land", "substeps_river", "substeps_ssf")
        return dict(zip(keys, list(out)))

    def newton_trace(self, enable: bool):
        """Per-node Newton iteration totals of the kinematic-wave solves (zeroed on enable)."""
        for dom, size in (("land", self.cfg["n"]), ("river", self.cfg["nriv"])):
            a = np.zeros(int(size), dtype=np.int64) if enable else None
            self._trace = getattr(self, "_trace", {})
            self._trace[dom] = a
            self._L.wfo_set_iptr(self.h, ("newton_trace_" + dom).encode(),
                                 a.ctypes.data if enable else None)

    def newton_trace_get(self, dom: str) -> np.ndarray:
        return self._trace[dom].copy()

    def sweep(self, name, dt=0.0):
        rc = self._L.wfo_sweep(self.h, name.encode(), dt)
        if rc != 0:
            raise KeyError(name)
