"""Host-side mirror of the reference's per-timestep interface for the `sbm` model type, on top
of the C ABI (include/wflow_b200.h). Method names follow the Julia functions they stand for
(`update_land_hydrology_model!` -> `update_land_hydrology_model`, ...; all under
/root/reference/Wflow/src, cited in the header). Julia itself is not available in this image,
so this Python class plays the role of the Julia shim shown in INTEGRATION.md.

The model state lives in HBM; `get`/`set` move one array (BMI get_value_ptr / set_value,
bmi.jl:208-248). Layered arrays cross this interface cell-major, shape (n, N), like Julia's
Vector{SVector{N,Float64}}.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

INT_FIELDS = {"number_of_layers": 0, "n_unsatlayers": 1, "nlayers_kv": 2}


class WflowB200Error(RuntimeError):
    pass


class ShardGroup:
    """The shards of one domain as handles of ONE process (several shards on one GPU, tests):
    their adaptive time-step statistics are reduced through a host rendezvous, so the handles
    must be stepped from `n` threads at the same time (`step_all`)."""

    def __init__(self, models):
        self._L = _lib.lib()
        self._g = C.c_void_p()
        self.models = list(models)
        rc = self._L.wflowb200_group_create(len(self.models), C.byref(self._g))
        if rc != 0:
            raise WflowB200Error(f"wflowb200_group_create failed ({rc})")
        for m in self.models:
            m._check(self._L.wflowb200_group_join(self._g, m._h))
            m._has_comm = True

    def step_all(self, dt):
        import threading
        errs = []

        def run(m):
            try:
                m.update_model(dt)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        ts = [threading.Thread(target=run, args=(m,)) for m in self.models]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            raise errs[0]

    def close(self):
        if self._g.value:
            self._L.wflowb200_group_destroy(self._g)
            self._g = C.c_void_p()


class SbmModel:
    """One handle = one GPU. `cfg` keys follow WflowB200Config; `domain` holds d1, d2,
    indices (n, 2) 1-based CartesianIndex pairs in column-major order, ldd (n,) uint8 and
    river_land_indices (nriv,) 1-based."""

    def __init__(self, cfg: dict, domain: dict, fields: dict | None = None, device: int = 0):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        self._table = _lib.field_table()
        self._ids = {name: i for i, (name, _) in enumerate(self._table)}
        self._kinds = {name: k for name, k in self._table}
        idx = np.ascontiguousarray(domain["indices"], dtype=np.int64)
        ldd = np.ascontiguousarray(domain["ldd"], dtype=np.uint8)
        rli = np.ascontiguousarray(domain["river_land_indices"], dtype=np.int64)
        rri = np.ascontiguousarray(domain.get("reservoir_river_indices", np.zeros(0)), dtype=np.int64)
        self.n, self.nriv, self.N, self.nres = len(ldd), len(rli), int(cfg["n_layers"]), len(rri)
        c = _lib.Config()
        c.n, c.nriv, c.n_layers, c.device = self.n, self.nriv, self.N, device
        for k in ("gash", "has_lai", "snow", "glacier", "soil_infiltration_reduction",
                  "kv_profile", "adaptive"):
            setattr(c, k, int(cfg.get(k, 0)))
        c.nthreads = int(cfg.get("nthreads", 1))
        c.land_streamorder_min = int(cfg.get("land_streamorder_min", 5))
        c.river_streamorder_min = int(cfg.get("river_streamorder_min", 6))
        c.dt_land = float(cfg.get("dt_land", 3600.0))
        c.dt_river = float(cfg.get("dt_river", 900.0))
        c.dt_ssf = float(cfg.get("dt_ssf", 86400.0))
        c.ssf_alpha_coefficient = float(cfg.get("ssf_alpha_coefficient", 1.0))
        c.kin_wave_min_flow_qroot = float(cfg.get("kin_wave_min_flow_qroot", 1e-30 ** 0.2))
        c.snow_gravitational_transport = int(cfg.get("snow_transport", 0))
        c.river_routing = int(cfg.get("river_routing", 0))
        c.li_froude_limit = int(cfg.get("li_froude_limit", 1))
        c.li_ghost_nodes = int(cfg.get("li_ghost_nodes", 1))
        c.li_alpha = float(cfg.get("li_alpha", 0.7))
        c.li_h_thresh = float(cfg.get("li_h_thresh", 1.0e-3))
        fp_depth = [float(x) for x in cfg.get("fp_depth", [])]
        c.fp_levels = len(fp_depth)
        for k, x in enumerate(fp_depth):
            c.fp_depth[k] = x
        self.P = max(len(fp_depth), 1)
        c.land_routing = int(cfg.get("land_routing", 0))
        c.li_land_froude_limit = int(cfg.get("li_land_froude_limit", 1))
        c.li_land_alpha = float(cfg.get("li_land_alpha", 0.7))
        c.li_land_theta = float(cfg.get("li_land_theta", 1.0))
        c.li_land_h_thresh = float(cfg.get("li_land_h_thresh", 1.0e-3))
        self.n6 = self.n if c.land_routing == 1 else 0   # li_land_* fields exist with land_routing = 1
        for k in ("wave_piece_depth_land", "vertical_slices", "unsat_inline_iters"):  # 0 = automatic
            setattr(c, k, int(cfg.get(k, 0)))
        d = _lib.Domain(int(domain["d1"]), int(domain["d2"]), idx.ctypes.data, ldd.ctypes.data,
                        rli.ctypes.data, len(rri), rri.ctypes.data if len(rri) else None)
        keep = []   # cut edges of a shard that is part of a basin (partition.cut_basin)
        for dom_name in ("land", "river"):
            for what in ("import_dst", "import_pos", "export_src"):
                a = np.ascontiguousarray(domain.get(f"{dom_name}_{what}", np.zeros(0)), dtype=np.int64)
                keep.append(a)
                setattr(d, f"{dom_name}_{what}", a.ctypes.data if len(a) else None)
                if what != "import_pos":
                    setattr(d, f"n_{dom_name}_{what.split('_')[0]}s", len(a))
        self.n_imports = (int(d.n_land_imports), int(d.n_river_imports))
        self.n_exports = (int(d.n_land_exports), int(d.n_river_exports))
        self.device = device
        rc = self._L.wflowb200_create(C.byref(c), C.byref(d), C.byref(self._h))
        if rc != 0:
            msg = self._L.wflowb200_last_error(None).decode()
            self._h = C.c_void_p()
            raise WflowB200Error(f"wflowb200_create failed ({rc}): {msg}")
        self.cfg = dict(cfg)
        if fields:
            for name, a in fields.items():
                if a is not None and (name in self._ids or name in INT_FIELDS):
                    self.set(name, a)

    # ---- lifetime ----------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.wflowb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise WflowB200Error(f"libwflow_b200 error {rc}: "
                                 f"{self._L.wflowb200_last_error(self._h).decode()}")

    # ---- state transfer ----------------------------------------------------------------
    def _shape(self, name):
        k = self._kinds[name]
        return {0: (self.n,), 1: (self.n, self.N), 2: (self.n, self.N + 1), 3: (self.nriv,),
                4: (self.nres,), 5: (self.nriv, self.P), 6: (self.n6,)}[k]

    def set(self, name: str, a) -> None:
        if name in INT_FIELDS:
            a = np.ascontiguousarray(a, dtype=np.int64)
            assert a.shape == (self.n,), (name, a.shape)
            self._check(self._L.wflowb200_set_field_i64(self._h, INT_FIELDS[name], a.ctypes.data))
            return
        a = np.ascontiguousarray(a, dtype=np.float64)
        shape = self._shape(name)
        if a.shape != shape:
            raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
        layers = shape[1] if len(shape) == 2 else 1
        self._check(self._L.wflowb200_set_field(self._h, self._ids[name], a.ctypes.data,
                                                layers, 1))

    def get(self, name: str) -> np.ndarray:
        if name in INT_FIELDS:
            a = np.zeros(self.n, dtype=np.int64)
            self._check(self._L.wflowb200_get_field_i64(self._h, INT_FIELDS[name], a.ctypes.data))
            return a
        shape = self._shape(name)
        a = np.empty(shape, dtype=np.float64)
        layers = shape[1] if len(shape) == 2 else 1
        self._check(self._L.wflowb200_get_field(self._h, self._ids[name], a.ctypes.data,
                                                layers, 1))
        return a

    def field_names(self):
        return [n for n, _ in self._table] + list(INT_FIELDS)

    def set_forcing(self, precipitation, potential_evaporation, temperature) -> None:
        """AtmosphericForcing hand-off (forcing.jl:2-10), asynchronous H2D."""
        p = np.ascontiguousarray(precipitation, dtype=np.float64)
        e = np.ascontiguousarray(potential_evaporation, dtype=np.float64)
        t = np.ascontiguousarray(temperature, dtype=np.float64)
        assert p.shape == e.shape == t.shape == (self.n,)
        self._check(self._L.wflowb200_set_forcing(self._h, p.ctypes.data, e.ctypes.data,
                                                  t.ctypes.data))

    # ---- staging either side of the path (io.jl:108-227, 815-899) ------------------------
    def forcing_ring_create(self, depth: int):
        self._check(self._L.wflowb200_forcing_ring_create(self._h, int(depth)))

    def forcing_ring_put(self, slot: int, p, e, t):
        """The arrays must stay alive (and unchanged) until the slab has been used."""
        for a in (p, e, t):
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape == (self.n,)
        self._check(self._L.wflowb200_forcing_ring_put(self._h, int(slot), p.ctypes.data,
                                                       e.ctypes.data, t.ctypes.data))

    def forcing_ring_use(self, slot: int):
        self._check(self._L.wflowb200_forcing_ring_use(self._h, int(slot)))

    def set_cyclic_lai(self, table):
        t = np.ascontiguousarray(table, dtype=np.float64)
        assert t.ndim == 2 and t.shape[1] == self.n
        self._check(self._L.wflowb200_set_cyclic_lai(self._h, t.ctypes.data, t.shape[0]))

    def use_cyclic_lai(self, slab: int):
        self._check(self._L.wflowb200_use_cyclic_lai(self._h, int(slab)))

    def get_fields(self, names) -> dict:
        """Several output vectors with one device-to-host copy (write_output, io.jl:815-899)."""
        ids = np.array([self._ids[n] for n in names], dtype=np.int32)
        sizes = [int(np.prod(self._shape(n))) for n in names]
        buf = np.empty(sum(sizes), dtype=np.float64)
        self._check(self._L.wflowb200_get_fields(self._h, ids.ctypes.data, len(ids), buf.ctypes.data))
        out, off = {}, 0
        for n, sz in zip(names, sizes):
            out[n] = buf[off:off + sz].reshape(self._shape(n))
            off += sz
        return out

    def output_size(self, names) -> int:
        return int(sum(np.prod(self._shape(n)) for n in names))

    def get_fields_async(self, names, out_pinned: np.ndarray):
        """Non-blocking get_fields into a page-locked float64 buffer of output_size(names)."""
        ids = np.array([self._ids[n] for n in names], dtype=np.int32)
        assert out_pinned.dtype == np.float64 and out_pinned.size >= self.output_size(names)
        self._check(self._L.wflowb200_get_fields_async(self._h, ids.ctypes.data, len(ids),
                                                       out_pinned.ctypes.data))

    def wait_outputs(self):
        self._check(self._L.wflowb200_wait_outputs(self._h))

    # ---- the hot path (names of the reference functions) ---------------------------------
    def update_land_hydrology_model(self, dt):
        self._check(self._L.wflowb200_update_land_hydrology_model(self._h, dt))

    def exchange_recharge(self):
        self._check(self._L.wflowb200_exchange_recharge(self._h))

    def update_subsurface_flow_model(self, dt):
        self._check(self._L.wflowb200_update_subsurface_flow_model(self._h, dt))

    def update_soil_water_storage(self, dt):
        self._check(self._L.wflowb200_update_soil_water_storage(self._h, dt))

    def update_lateral_inflow_overland(self):
        self._check(self._L.wflowb200_update_lateral_inflow_overland(self._h))

    def update_overland_flow_model(self, dt):
        self._check(self._L.wflowb200_update_overland_flow_model(self._h, dt))

    def update_lateral_inflow_river(self):
        self._check(self._L.wflowb200_update_lateral_inflow_river(self._h))

    def update_inflow_reservoir(self):
        self._check(self._L.wflowb200_update_inflow_reservoir(self._h))

    def update_river_flow_model(self, dt):
        self._check(self._L.wflowb200_update_river_flow_model(self._h, dt))

    def update_bc_overland_flow_model(self):
        self._check(self._L.wflowb200_update_bc_overland_flow_model(self._h))

    def surface_routing(self, dt):
        """surface_routing! (routing/surface/surface_routing.jl:7-46; :62-86 with local-inertial
        land and river routing)."""
        if self.cfg.get("land_routing", 0) == 1:
            self.update_bc_overland_flow_model()
            self.update_inflow_reservoir()
            self.update_overland_flow_model(dt)   # overland and river flow, one scheme
            return
        self.update_lateral_inflow_overland()
        self.update_overland_flow_model(dt)
        self.update_lateral_inflow_river()
        self.update_inflow_reservoir()
        self.update_river_flow_model(dt)

    def update_total_water_storage(self):
        self._check(self._L.wflowb200_update_total_water_storage(self._h))

    def update_model(self, dt):
        """update_model!(model::AbstractModel{<:SbmModel}) (sbm_model.jl:60-92)."""
        if self.cfg.get("sharded") and self.cfg.get("adaptive") and not self._has_comm:
            raise WflowB200Error(
                "a shard of a domain with adaptive internal time steps needs the statistics of "
                "ALL shards: attach a communicator first (comm_init_nccl / ShardGroup.join)")
        self._check(self._L.wflowb200_update_model(self._h, dt))

    # ---- sharded domains: reductions of the adaptive time-step statistics ------------------
    _has_comm = False

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = _lib.lib().wflowb200_comm_unique_id(buf)
        if rc != 0:
            raise WflowB200Error(f"wflowb200_comm_unique_id failed ({rc})")
        return buf.raw

    def comm_init_nccl(self, rank: int, world: int, unique_id: bytes):
        """One process per GPU: every rank passes the 128 bytes rank 0 got from comm_unique_id."""
        self._check(self._L.wflowb200_comm_init_nccl(self._h, int(rank), int(world), unique_id))
        self._has_comm = True

    # ---- cut edges (wflowb200_exchange_*) -------------------------------------------------
    def exchange_prepare(self, dt: float):
        """-> (device pointer, 64-byte CUDA IPC handle, bytes) of this shard's import buffer"""
        ptr, nbytes = C.c_uint64(0), C.c_int64(0)
        ipc = C.create_string_buffer(64)
        self._check(self._L.wflowb200_exchange_prepare(self._h, float(dt), C.byref(ptr), ipc,
                                                       C.byref(nbytes)))
        return int(ptr.value), ipc.raw, int(nbytes.value)

    def exchange_open_peer(self, peer: int, n_imports, device_ptr: int = 0, peer_device: int = 0,
                           ipc_handle: bytes | None = None):
        buf = C.create_string_buffer(ipc_handle, 64) if ipc_handle is not None else None
        self._check(self._L.wflowb200_exchange_open_peer(self._h, int(peer), int(device_ptr),
                                                         int(peer_device), buf,
                                                         int(n_imports[0]), int(n_imports[1])))

    def exchange_bind(self, domain: int, export_index: int, peer: int, peer_import_index: int):
        self._check(self._L.wflowb200_exchange_bind(self._h, int(domain), int(export_index),
                                                    int(peer), int(peer_import_index)))

    def synchronize(self):
        self._check(self._L.wflowb200_synchronize(self._h))

    def set_option(self, name: str, value: int):
        """Select between kernel organisations with identical results (wflow_b200.h)."""
        self._check(self._L.wflowb200_set_option(self._h, name.encode(), int(value)))

    def vertical_timeline(self):
        buf = (C.c_double * 4)()
        self._check(self._L.wflowb200_get_vertical_timeline(self._h, buf, 4))
        return list(buf)

    def unsat_buckets(self):
        """cells handed to the loop engine in the last step, by log2 of the suspended loop's trips"""
        buf = (C.c_int64 * 16)()
        self._check(self._L.wflowb200_get_unsat_buckets(self._h, buf, 16))
        return list(buf)[:8], list(buf)[8:]

    def newton_trace(self, enable: bool):
        self._check(self._L.wflowb200_newton_trace(self._h, int(enable)))

    def newton_trace_get(self, domain: str) -> np.ndarray:
        dom = {"land": 0, "river": 1}[domain]
        a = np.zeros(self.n if dom == 0 else self.nriv, dtype=np.int64)
        self._check(self._L.wflowb200_get_newton_trace(self._h, dom, a.ctypes.data))
        return a

    # ---- artefacts / statistics ----------------------------------------------------------
    def artifact(self, domain: str, name: str) -> np.ndarray:
        dom = {"land": 0, "river": 1}[domain]
        n = C.c_int64()
        aid = _lib.ARTIFACTS[name] if name in _lib.ARTIFACTS else _lib.EDGE_ARTIFACTS[name]
        self._check(self._L.wflowb200_get_artifact(self._h, dom, aid, None, 0, C.byref(n)))
        a = np.zeros(n.value, dtype=np.int64)
        self._check(self._L.wflowb200_get_artifact(self._h, dom, aid, a.ctypes.data, n.value,
                                                   C.byref(n)))
        return a

    def artifacts(self, domain: str) -> dict:
        return {k: self.artifact(domain, k) for k in _lib.ARTIFACTS}

    def set_timing(self, enabled: bool):
        self._check(self._L.wflowb200_set_timing(self._h, int(enabled)))

    def timer_start(self):
        self._check(self._L.wflowb200_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._check(self._L.wflowb200_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def stats(self) -> dict:
        s = _lib.Stats()
        self._check(self._L.wflowb200_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}
