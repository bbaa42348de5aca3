"""Shared helpers of the parity tests: drive the CUDA library (through the C ABI, via the host
mirror) and the CPU oracle on the same seeded synthetic basin and compare every field."""
from __future__ import annotations

import numpy as np

from oracle import network as onw
from oracle import oracle as orc

RTOL = 1e-10  # north_star: storage, discharge and flux fields within 1e-10 relative (Float64)


def oracle_networks(cfg, dom):
    """The oracle builds its OWN artefacts (oracle/network.py), independent of the product."""
    land = onw.build_domain_network(dom["ldd"], dom["indices"], dom["d1"],
                                    cfg["land_streamorder_min"], cfg["nthreads"])
    rl = dom["river_land_indices"]
    river = onw.build_domain_network(dom["ldd"][rl - 1], dom["indices"][rl - 1], dom["d1"],
                                     cfg["river_streamorder_min"], cfg["nthreads"],
                                     streamorder=land["streamorder"][rl - 1])
    return land, river


def make_oracle(cfg, dom, fields, nets=None):
    land, river = nets if nets is not None else oracle_networks(cfg, dom)
    f = dict(fields)
    f["river_land_indices"] = dom["river_land_indices"] - 1
    return orc.OracleModel(cfg, f, land, river)


def run_pair(pkg, d1, d2, steps=2, seed=42, fine_grained=False, **kw):
    cfg, dom, fields = pkg.synthetic.make_basin(d1, d2, seed=seed, **kw)
    dt = cfg["dt"]
    gpu = pkg.SbmModel(cfg, dom, fields)
    ora = make_oracle(cfg, dom, fields)
    for step in range(steps):
        p, e, t = pkg.synthetic.make_forcing(seed, step, dom["gid"], dt)
        gpu.set_forcing(p, e, t)
        ora.f["precipitation"][:] = p
        ora.f["potential_evaporation"][:] = e
        ora.f["temperature"][:] = t
        if fine_grained:
            for m in (gpu, ora):
                m.update_land_hydrology_model(dt)
                m.exchange_recharge()
                m.update_subsurface_flow_model(dt)
                m.update_soil_water_storage(dt)
                m.surface_routing(dt)
                m.update_total_water_storage()
        else:
            gpu.update_model(dt)
            ora.update_model(dt)
    gpu.synchronize()
    return gpu, ora, cfg


# Fields that are exact-cancellation residues of larger operands: their own magnitude is
# rounding noise (|x| ~ eps * operand), so they are judged against the operand's scale.
OPERAND_SCALE = {
    "saturation_excess_water": "soil_water_flux_surface",  # (flux - act_infilt) - infilt_excess
    "excess_water_soil": "soil_water_flux_surface",
    "excess_water_compacted_soil": "soil_water_flux_surface",
    "runoff": "soil_water_flux_surface",
    "net_runoff": "soil_water_flux_surface",
    "olf_inwater": "riv_inwater",
    "recharge": "transfer",
    "recharge_rate": "transfer",
}


def compare_models(gpu, ora, rtol=RTOL, skip=(), verbose=False):
    """Every Float64 field and both integer fields. NaN (MISSING_VALUE) must match NaN.
    |g - o| <= rtol * max(|o|, scale) with scale = the field's largest magnitude, so that
    exact-cancellation residues of O(eps * scale) do not count as relative errors.
    Returns the worst scaled relative difference."""
    worst, worst_name = 0.0, ""
    for name in gpu.field_names():
        if name in skip:
            continue
        g = gpu.get(name)
        o = ora.f[name]
        assert g.shape == o.shape, (name, g.shape, o.shape)
        if g.dtype.kind == "i":
            assert np.array_equal(g, o), f"{name}: integer field differs"
            continue
        gn, on = np.isnan(g), np.isnan(o)
        assert np.array_equal(gn, on), f"{name}: NaN pattern differs ({gn.sum()} vs {on.sum()})"
        m = ~on
        if not m.any():
            continue
        gi, oi = np.isinf(g[m]), np.isinf(o[m])
        assert np.array_equal(gi, oi) and np.array_equal(g[m][gi], o[m][oi]), f"{name}: inf differs"
        fin = ~oi
        if not fin.any():
            continue
        gv, ov = g[m][fin], o[m][fin]
        scale = float(np.max(np.abs(ov)))
        if name in OPERAND_SCALE:
            ref = ora.f[OPERAND_SCALE[name]]
            if ref.size and np.isfinite(ref).any():
                scale = max(scale, float(np.nanmax(np.abs(ref))))
        if scale == 0.0:
            assert np.all(gv == 0.0), f"{name}: expected all zeros"
            continue
        rel = np.abs(gv - ov) / np.maximum(np.abs(ov), scale)
        w = float(rel.max())
        if verbose:
            print(f"{name:48s} {w:.3e}")
        if w > worst:
            worst, worst_name = w, name
        assert w <= rtol, f"{name}: scaled relative difference {w:.3e} > {rtol:g}"
    return worst
