"""Host-only access to the indexing artefacts built by libwflow_b200 (no GPU needed):
topological order, Strahler order, upstream CSR, sub-domain partition, wavefront levels."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def build_network_artifacts(cfg: dict, domain: dict) -> dict:
    """Returns {"land": {...}, "river": {...}} with the 1-based int64 artefacts of
    include/wflow_b200.h (WFLOWB200_A_*)."""
    L = _lib.lib()
    L.wflowb200_network_build.argtypes = [C.POINTER(_lib.Config), C.POINTER(_lib.Domain),
                                          C.POINTER(C.c_void_p)]
    L.wflowb200_network_get.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                        C.POINTER(C.c_int64)]
    L.wflowb200_network_destroy.argtypes = [C.c_void_p]
    L.wflowb200_network_destroy.restype = None
    idx = np.ascontiguousarray(domain["indices"], dtype=np.int64)
    ldd = np.ascontiguousarray(domain["ldd"], dtype=np.uint8)
    rli = np.ascontiguousarray(domain["river_land_indices"], dtype=np.int64)
    c = _lib.Config()
    c.n, c.nriv, c.n_layers = len(ldd), len(rli), int(cfg.get("n_layers", 4))
    c.nthreads = int(cfg.get("nthreads", 1))
    c.land_streamorder_min = int(cfg.get("land_streamorder_min", 5))
    c.river_streamorder_min = int(cfg.get("river_streamorder_min", 6))
    c.land_routing = int(cfg.get("land_routing", 0))     # 1: also EdgeConnectivity of the land
    c.river_routing = int(cfg.get("river_routing", 0))
    d = _lib.Domain(int(domain["d1"]), int(domain["d2"]), idx.ctypes.data, ldd.ctypes.data,
                    rli.ctypes.data, 0, None)
    net = C.c_void_p()
    rc = L.wflowb200_network_build(C.byref(c), C.byref(d), C.byref(net))
    if rc != 0:
        raise RuntimeError(f"wflowb200_network_build failed ({rc}): "
                           f"{L.wflowb200_last_error(None).decode()}")
    out = {}
    try:
        for dom_name, dom_id in (("land", 0), ("river", 1)):
            o = {}
            ids = dict(_lib.ARTIFACTS)
            if dom_id == 0 and c.land_routing == 1:
                ids.update(_lib.EDGE_ARTIFACTS)
            for name, aid in ids.items():
                m = C.c_int64()
                L.wflowb200_network_get(net, dom_id, aid, None, 0, C.byref(m))
                a = np.zeros(m.value, dtype=np.int64)
                L.wflowb200_network_get(net, dom_id, aid, a.ctypes.data, m.value, C.byref(m))
                o[name] = a
            out[dom_name] = o
    finally:
        L.wflowb200_network_destroy(net)
    return out
