"""Shared helpers of the parity tests: drive the CUDA library (through the C ABI, via the host
mirror) and the CPU oracle on the same seeded synthetic basin and compare every field.

The comparison is ELEMENTWISE:  |g - o| <= RTOL * |o| + atol_f[i]  with RTOL = 1e-10 (north_star:
storage, discharge and flux fields within 1e-10 relative in Float64). atol_f is zero for most
fields. It is non-zero only where the reference's own expression makes a pure relative test
meaningless, and it is then derived PER ELEMENT from the operands of that expression (never from
the field's maximum):

* cancellation residues -- a field computed as a difference of like-sized terms carries the
  rounding noise of its operands: atol_f[i] = RTOL * sum |operand_k[i]| (ATOL_OPERANDS below
  lists the reference expression for every such field);
* the kinematic-wave Newton iteration stops on an ABSOLUTE residual |f(u)| <= 1e-12
  (surface_process.jl:52-57), so two faithful evaluations may stop one iterate apart:
  |dq| <= 5 u^4 * 2e-12 / f'(u) <= 2e-12 * dx / dt_sub  [m3 s-1], |dA| <= 2e-12 [m2] per solve,
  and a model step is S internal sub-steps, each starting from the previous one's result:
  S * 2e-12 m2 (NEWTON_FLOOR below turns that into a floor for q, h, storage, their
  cumulatives and what is computed from h).

tests/test_tolerances.py calibrates this table on the CPU: the oracle against itself on a libm
that is noisy by +-1 ulp (oracle/wfo_math.h: WFO_ALT_LIBM) must pass with it, at 1000 x 1000.
The old normwise number (difference over the field's largest magnitude) is still printed as a
second diagnostic.
"""
from __future__ import annotations

import numpy as np

from oracle import network as onw
from oracle import oracle as orc

RTOL = 1e-10  # north_star: storage, discharge and flux fields within 1e-10 relative (Float64)
# Values this small (SI: m, m s-1, m3 s-1) are noise of the kinematic wave's own absolute
# tolerance carried into other fields: 2e-12 m2 of cross-section over a ~1 km cell and a day is
# ~2e-20 m s-1 of open-water evaporation, lateral inflow, ... . One yoctometre per second.
ATOL_FLOOR = 1e-18


def oracle_networks(cfg, dom):
    """The oracle builds its OWN artefacts (oracle/network.py), independent of the product."""
    rl = dom["river_land_indices"]
    pits_land = pits_river = None
    rr = dom.get("reservoir_river_indices")
    if rr is not None and len(rr):      # reservoir outlets (domain.jl:96-109)
        pits_river = np.zeros(len(rl), dtype=bool)
        pits_river[rr - 1] = True
        pits_land = np.zeros(len(dom["ldd"]), dtype=bool)
        pits_land[rl[rr - 1] - 1] = True
    land = onw.build_domain_network(dom["ldd"], dom["indices"], dom["d1"],
                                    cfg["land_streamorder_min"], cfg["nthreads"],
                                    pits_mask=pits_land)
    river = onw.build_domain_network(dom["ldd"][rl - 1], dom["indices"][rl - 1], dom["d1"],
                                     cfg["river_streamorder_min"], cfg["nthreads"],
                                     streamorder=land["streamorder"][rl - 1], pits_mask=pits_river)
    return land, river


def make_oracle(cfg, dom, fields, nets=None, variant=""):
    land, river = nets if nets is not None else oracle_networks(cfg, dom)
    f = dict(fields)
    f["river_land_indices"] = dom["river_land_indices"] - 1
    if cfg.get("nres", 0):
        f["reservoir_river_indices"] = dom["reservoir_river_indices"] - 1
    if cfg.get("land_routing", 0) == 1:   # the oracle's OWN EdgeConnectivity (oracle/network.py)
        e = onw.edge_connectivity(dom["indices"], dom["d1"], dom["d2"])
        for k in ("x_up", "x_down", "y_up", "y_down"):
            f["edge_" + k] = e["ind_" + k] - 1
        lri = np.full(len(dom["ldd"]), -1, dtype=np.int64)     # network_land.river_indices
        lri[dom["river_land_indices"] - 1] = np.arange(len(dom["river_land_indices"]))
        f["land_river_indices"] = lri
    return orc.OracleModel(cfg, f, land, river, variant=variant)


def step_models(pkg, models, dom, cfg, seed, steps, fine_grained=False, first_step=0):
    """Advance every model (GPU handle or oracle: same method names) over the same forcing."""
    dt = cfg["dt"]
    for step in range(first_step, first_step + steps):
        p, e, t = pkg.synthetic.make_forcing(seed, step, dom["gid"], dt)
        for m in models:
            if hasattr(m, "set_forcing"):
                m.set_forcing(p, e, t)
            else:
                m.f["precipitation"][:] = p
                m.f["potential_evaporation"][:] = e
                m.f["temperature"][:] = t
            if fine_grained:
                m.update_land_hydrology_model(dt)
                m.exchange_recharge()
                m.update_subsurface_flow_model(dt)
                m.update_soil_water_storage(dt)
                m.surface_routing(dt)
                m.update_total_water_storage()
            else:
                m.update_model(dt)


def run_pair(pkg, d1, d2, steps=2, seed=42, fine_grained=False, cfg_over=None, options=None,
             newton_trace=False, **kw):
    """cfg_over: WflowB200Config tuning fields fixed at create; options: wflowb200_set_option
    pairs; kw: synthetic.make_basin arguments."""
    cfg, dom, fields = pkg.synthetic.make_basin(d1, d2, seed=seed, **kw)
    gcfg = dict(cfg)
    gcfg.update(cfg_over or {})
    gpu = pkg.SbmModel(gcfg, dom, fields)
    for k, v in (options or {}).items():
        gpu.set_option(k, v)
    ora = make_oracle(cfg, dom, fields)
    if newton_trace:
        gpu.newton_trace(True)
        ora.newton_trace(True)
    step_models(pkg, (gpu, ora), dom, cfg, seed, steps, fine_grained)
    gpu.synchronize()
    return gpu, ora, cfg


# ---------------------------------------------------------------------------------------------
# per-element absolute tolerances
# ---------------------------------------------------------------------------------------------
# field -> operands of the reference expression that produces it (file:line), per element.
# A trailing "*area" / "*dt" scales the operand like the expression does.
_SOIL_FLUX = ("soil_water_flux_surface",)
_RECHARGE = ("transfer", "actual_capillary_flux", "actual_leakage",
             "actual_evaporation_saturated_zone", "soil_evaporation_saturated_zone")
_SSF_NET = ("ssf_q_in", "ssf_q", "ssf_q_net_bnds")
ATOL_OPERANDS = {
    # (flux - actual_infiltration) - infiltration_excess                     soil.jl:1180-1183
    "saturation_excess_water": _SOIL_FLUX,
    # max(flux * (1 - pathfrac) - actinf_soil, 0), max(flux * pathfrac - actinf_path, 0)  :1185-1192
    "excess_water_soil": _SOIL_FLUX,
    "excess_water_compacted_soil": _SOIL_FLUX,
    # infiltration - ustoredepth_excess / dt                                       soil.jl:1015
    "actual_infiltration": ("infiltration",),
    "actual_infiltration_soil": ("infiltration",),
    "actual_infiltration_compacted_soil": ("infiltration",),
    # max(0, exfilt + saturation_excess + runoff_land + infiltration_excess)     soil.jl:1318-1324
    "runoff": _SOIL_FLUX + ("ssf_exfiltwater_average",),
    # runoff - actual_open_water_evaporation_land                               soil.jl:1390
    "net_runoff": _SOIL_FLUX + ("ssf_exfiltwater_average", "actual_open_water_evaporation_land"),
    # net_runoff * area                                                 surface_kinwave.jl:757-765
    "olf_inwater": ("soil_water_flux_surface*area", "ssf_exfiltwater_average*area",
                    "actual_open_water_evaporation_land*area"),
    "olf_qlat": ("soil_water_flux_surface*area/flow_length", "ssf_exfiltwater_average*area/flow_length",
                 "actual_open_water_evaporation_land*area/flow_length"),
    # runoff_river - actual_open_water_evaporation_river                        runoff.jl:108
    "net_runoff_river": ("runoff_river", "actual_open_water_evaporation_river"),
    # transfer - capillary flux - leakage - ae_sat - soilevap_sat               soil.jl:1201-1204
    "recharge": _RECHARGE,
    "recharge_rate": _RECHARGE,
    "recharge_flux": tuple(o + "*area" for o in _RECHARGE),
    "recharge_flux_average": tuple(o + "*area" for o in _RECHARGE),
    "recharge_flux_cumulative": tuple(o + "*area*dt" for o in _RECHARGE),
    "ssf_q_net_bnds": tuple(o + "*area" for o in _RECHARGE),
    # (q_in + q_net_bnds - q) / (dw dx) * area                  subsurface_process.jl:127, lsf.jl:262
    "ssf_q_net_average": _SSF_NET,
    "ssf_q_net_cumulative": tuple(o + "*dt" for o in _SSF_NET),
    # soil_water_capacity - satwaterdepth - ustoredepth                   soil.jl:1196,1365
    "unsaturated_store_capacity": ("soil_water_capacity",),
    # zi = max(0, soil_thickness - satwaterdepth / (theta_s - theta_r))   soil.jl:1424
    # (and everything that is a difference against zi)
    "water_table_depth": ("soil_thickness",),
    "ssf_water_table_depth": ("soil_thickness",),
    "ssf_head": ("ssf_top",),
    "unsaturated_layer_thickness": ("soil_thickness",),
    "drainable_water_depth": ("soil_thickness",),
    # sy * (d - zi) * area                                         lateral_subsurface_flow.jl:267
    "ssf_storage": ("soil_thickness*area",),
    # (d - zi) * theta_e                                                     soil.jl:1360
    "saturated_water_depth": ("soil_water_capacity",),
    # snow - melt * dt, snowwater - refreeze, ...                        snow_process.jl:40-66
    "snow_storage": ("snow_water_equivalent",),
    "snow_water": ("snow_water_equivalent",),
    # throughfall = P - interception - stemflow                  rainfall_interception.jl:58
    "throughfall": ("precipitation",),
    "effective_precip": ("precipitation",),
    "liquid_precip": ("precipitation",),
    "snow_precip": ("precipitation",),
    # to_river + net_runoff_river * area                         surface_kinwave.jl:720-733
    "riv_inwater": ("runoff_river*area", "actual_open_water_evaporation_river*area",
                    "ssf_to_river_average", "olf_to_river_average"),
    "riv_qlat": ("riv_inwater/riv_flow_length",),
}
# Newton floor (see the module docstring): per element, in the field's own unit.
NEWTON_FLOOR = {
    "olf_q": "2e-12*flow_length/dt_land*S_land", "olf_q_average": "2e-12*flow_length/dt_land*S_land",
    "olf_qin": "2e-12*flow_length/dt_land*S_land", "olf_qin_average": "2e-12*flow_length/dt_land*S_land",
    "olf_q_cumulative": "2e-12*flow_length/dt_land*dt*S_land",
    "olf_qin_cumulative": "2e-12*flow_length/dt_land*dt*S_land",
    "olf_to_river_cumulative": "2e-12*flow_length/dt_land*dt*S_land",
    "olf_to_river_average": "2e-12*flow_length/dt_land*S_land",
    "olf_h": "2e-12/surface_flow_width*S_land", "olf_storage": "2e-12*flow_length*S_land",
    "waterdepth_land": "2e-12/surface_flow_width*S_land",
    "riv_q": "2e-12*riv_flow_length/dt_river*S_river", "riv_q_average": "2e-12*riv_flow_length/dt_river*S_river",
    "riv_qin": "2e-12*riv_flow_length/dt_river*S_river",
    "riv_qin_average": "2e-12*riv_flow_length/dt_river*S_river",
    "riv_q_cumulative": "2e-12*riv_flow_length/dt_river*dt*S_river",
    "riv_qin_cumulative": "2e-12*riv_flow_length/dt_river*dt*S_river",
    "riv_h": "2e-12/riv_flow_width*S_river", "riv_storage": "2e-12*riv_flow_length*S_river",
    # ... and what is computed from the water depths h = A / width of the previous step:
    # open-water evaporation wf * min(h / dt, PET) (runoff.jl:104-107) and the lateral inflows
    # built on it (soil.jl:1390, surface_kinwave.jl:720-765); river width >= 1 m
    "waterdepth_river": "2e-12*S_river",
    "actual_open_water_evaporation_land": "2e-12/surface_flow_width/dt*S_land",
    "actual_open_water_evaporation_river": "2e-12/dt*S_river",
    "net_runoff": "2e-12/surface_flow_width/dt*S_land",
    "olf_inwater": "2e-12/surface_flow_width/dt*area*S_land",
    "olf_qlat": "2e-12/surface_flow_width/dt*area/flow_length*S_land",
    "net_runoff_river": "2e-12/dt*S_river",
    "riv_inwater": "2e-12/dt*area*S_river",
    "riv_qlat": "2e-12/dt*area/riv_flow_length*S_river",
}


# Storages [m] that the reference updates by differencing against their own previous value
# (snow - min(pot, snow / dt) * dt, usd - st * dt, canopy storage - evaporation * dt, ...): an
# emptied store keeps eps * (its metre-scale previous value) instead of 0. One femtometre.
STATE_FLOOR = {name: 1e-15 for name in (
    "snow_storage", "snow_water", "snow_water_equivalent", "canopy_storage",
    "unsaturated_layer_depth", "unsaturated_store_depth", "glacier_store")}


def _term(expr, get, cfg):
    """|a * b / c ...| for an expression of field names, cfg keys and numbers."""
    out = None
    tok, op = "", "*"
    for ch in expr + "*":
        if ch in "*/":
            if tok in ("dt", "dt_land", "dt_river", "dt_ssf", "S_land", "S_river"):
                v = float(cfg[tok])
            else:
                try:
                    v = float(tok)
                except ValueError:
                    v = np.abs(np.nan_to_num(get(tok), nan=0.0, posinf=0.0, neginf=0.0))
            if out is None:
                out = v
            elif op == "*":
                out = out * v
            else:
                with np.errstate(divide="ignore", invalid="ignore"):
                    out = np.where(v > 0, out / np.where(v > 0, v, 1.0), 0.0)
            tok, op = "", ch
        else:
            tok += ch
    return out


def atol_of(name, get, cfg, shape, rtol=RTOL, rli=None):
    """Per-element absolute tolerance of a field (zeros unless listed above). rli: 0-based land
    index of every river cell (land operands of a river field are taken at the river's cell)."""
    a = np.full(shape, max(ATOL_FLOOR, STATE_FLOOR.get(name, 0.0)))

    def fit(t):
        if np.ndim(t) == 1 and len(shape) == 2:
            return t[:, None]
        if np.ndim(t) == 1 and len(t) != shape[0]:
            return t[rli]
        return t
    def product(expr):
        t = 1.0
        for fac in expr.split("*"):          # each factor fitted to the field's shape first
            num, *den = fac.split("/")
            t = t * fit(_term(num, get, cfg))
            for dn in den:
                v = fit(_term(dn, get, cfg))
                with np.errstate(divide="ignore", invalid="ignore"):
                    t = np.where(v > 0, t / np.where(v > 0, v, 1.0), 0.0)
        return t
    for expr in ATOL_OPERANDS.get(name, ()):
        a = a + rtol * product(expr)
    if name in NEWTON_FLOOR:
        a = a + product(NEWTON_FLOOR[name])
    return a


class Report(dict):
    """field -> dict(rel=worst |g-o|/|o| over elements above their atol, need_atol=number of
    elements that pass only thanks to atol, norm=the old normwise number)."""

    @property
    def worst_rel(self):
        return max((v["rel"] for v in self.values()), default=0.0)

    def summary(self):
        k = max(self, key=lambda n: self[n]["rel"]) if self else ""
        n_atol = sum(v["need_atol"] for v in self.values())
        n_outl = sum(v.get("outliers", 0) for v in self.values())
        extra = f"; {n_outl} threshold outliers" if n_outl else ""
        return (f"{len(self)} fields: worst elementwise rel diff {self.worst_rel:.3e} ({k}){extra}; "
                f"{n_atol} elements inside their absolute tolerance only; worst normwise "
                f"{max((v['norm'] for v in self.values()), default=0.0):.3e}")


def _getter(m):
    return m.get if hasattr(m, "field_names") else (lambda name: m.f[name])


def compare_models(gpu, ora, rtol=RTOL, skip=(), verbose=False, names=None, outliers=None):
    """Every Float64 field and the integer fields. NaN (MISSING_VALUE) must match NaN, +-Inf must
    match exactly, integers must be equal; Float64: |g - o| <= rtol |o| + atol_f (module
    docstring). `gpu` may be a second oracle (tolerance calibration). Returns a Report.
    outliers = (fraction, rtol_out): at most that fraction of a field's elements may miss the test
    as long as they pass it with rtol_out -- only for schemes whose own thresholds turn a last-bit
    difference into a visible one (the local-inertial river flow switches an edge's discharge on
    where the depth crosses h_thresh exactly, surface_staggered_scheme.jl:359-377); the number of
    such elements is reported in Report[...]["outliers"]."""
    gget, oget = _getter(gpu), _getter(ora)
    cfg = dict(ora.cfg)
    st = ora.newton_stats()   # sub-steps of the last model step (fixed or adaptive)
    cfg["S_land"], cfg["S_river"] = max(st["substeps_land"], 1), max(st["substeps_river"], 1)
    if cfg.get("adaptive"):      # mean sub-step length instead of the fixed one
        cfg["dt_land"], cfg["dt_river"] = cfg["dt"] / cfg["S_land"], cfg["dt"] / cfg["S_river"]
    rli = np.asarray(ora.f["river_land_indices"], dtype=np.int64)
    rep = Report()
    if names is None:
        names = gpu.field_names() if hasattr(gpu, "field_names") else list(ora.f)
    if cfg.get("land_routing", 0) == 1:
        # LocalInertialOverlandFlow has no kinematic-wave variables (surface_staggered_scheme.jl:
        # 840-865): of the olf_* fields only h and storage exist in the reference's model
        skip = tuple(skip) + tuple(n for n in names if n.startswith("olf_") and
                                   n not in ("olf_h", "olf_storage"))
    for name in names:
        if name in skip or name in ("river_land_indices", "land_river_indices") or name.startswith("edge_"):
            continue
        g = gget(name)
        o = oget(name)
        assert g.shape == o.shape, (name, g.shape, o.shape)
        if g.dtype.kind == "i":
            assert np.array_equal(g, o), f"{name}: integer field differs ({(g != o).sum()} cells)"
            continue
        gn, on = np.isnan(g), np.isnan(o)
        assert np.array_equal(gn, on), f"{name}: NaN pattern differs ({gn.sum()} vs {on.sum()})"
        gi, oi = np.isinf(g), np.isinf(o)
        assert np.array_equal(gi, oi) and np.array_equal(g[gi], o[oi]), f"{name}: inf differs"
        fin = ~(on | oi)
        if not fin.any():
            continue
        atol = atol_of(name, oget, cfg, o.shape, rtol, rli)
        diff = np.abs(np.where(fin, g - o, 0.0))
        mag = np.abs(np.where(fin, o, 0.0))
        bad = diff > rtol * mag + atol
        n_out = 0
        if outliers is not None and bad.any():
            frac, rtol_out = outliers
            really_bad = diff > rtol_out * mag + atol * (rtol_out / rtol)
            n_out = int((bad & ~really_bad).sum())
            if n_out <= frac * max(int(fin.sum()), 1):
                bad = really_bad
        above = fin & (diff > atol)          # judged by the relative term
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = np.where(above & (mag > 0), diff / np.where(mag > 0, mag, 1.0), 0.0)
        need = fin & (diff > rtol * mag) & ~bad
        scale = float(mag.max())
        if n_out:
            rel = np.where(bad | ~(diff > rtol * mag + atol), rel, 0.0)  # outliers reported apart
        rep[name] = dict(rel=float(rel.max()), need_atol=int(need.sum()), outliers=n_out,
                         norm=float(diff.max() / scale) if scale > 0 else float(diff.max()))
        if verbose:
            print(f"{name:48s} rel {rep[name]['rel']:.3e}  atol-only {rep[name]['need_atol']:8d}  "
                  f"normwise {rep[name]['norm']:.3e}  max atol {float(np.max(atol)):.2e}")
        if bad.any():
            k = np.unravel_index(np.argmax(np.where(bad, diff - rtol * mag - atol, -1.0)), o.shape)
            raise AssertionError(
                f"{name}: {int(bad.sum())} of {int(fin.sum())} elements outside "
                f"|g-o| <= {rtol:g}|o| + atol; worst at {k}: g={g[k]!r} o={o[k]!r} "
                f"diff={diff[k]:.3e} atol={float(np.broadcast_to(atol, o.shape)[k]):.3e}")
    return rep


def newton_parity(gpu, ora):
    """Per-node Newton iteration totals of the kinematic-wave solves since newton_trace(True):
    returns dict(domain -> (nodes, nodes whose totals differ, sum gpu, sum oracle))."""
    out = {}
    for dom in ("land", "river"):
        g = gpu.newton_trace_get(dom)
        o = ora.newton_trace_get(dom)
        assert g.shape == o.shape
        out[dom] = (int(g.size), int((g != o).sum()), int(g.sum()), int(o.sum()))
    return out
