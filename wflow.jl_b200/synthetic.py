"""Synthetic wflow_sbm basins for tests and benchmarks: a Scheidegger-type
random D8 forest, parameters drawn per cell +-20 % around the Moselle means asserted in
/root/reference/Wflow/test/run_sbm.jl:98-406, cold-start states as in
Wflow/src/soil/soil.jl:164-198, and counter-based forcing. Everything is a pure function of
(seed, cell id, step), so any shard of a raster can be generated independently.

The init-time derivations that the Julia model constructors perform are restated here with
the reference's formulas (cited inline); they are host-side set-up, not part of the hot path.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays."""
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def u01(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    """Counter-based uniform [0, 1): a pure function of (seed, stream, cell id)."""
    with np.errstate(over="ignore"):
        k = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
             + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03))
        x = _mix(idx.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + k)
        x = _mix(x + np.uint64(0x632BE59BD9B4E019))
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _pm20(seed, stream, idx, mean):
    return mean * (0.8 + 0.4 * u01(seed, stream, idx))


def jl_pow(x, y):
    """Wflow's pow(x, y) = exp(y * log(x)) (utils.jl:470)."""
    with np.errstate(divide="ignore"):
        return np.exp(y * np.log(x))


def scheidegger_ldd(d1: int, d2: int, seed: int, mask: np.ndarray | None = None,
                    catchment_length: int = 0):
    """Random D8 forest: cell (i, j) drains to (i + delta, j + 1), delta in {-1, 0, +1};
    the last row and cells draining out of the active domain are pits (LDD 5). With
    catchment_length > 0 every column that is a multiple of it is an outlet line as well: the
    raster is a mosaic of catchments that are catchment_length cells long.
    Returns (indices (n,2) 1-based column-major, ldd uint8 (n,), down (n,) 1-based, 0 = pit,
    gid (n,) global cell id)."""
    if mask is None:
        lin = np.arange(d1 * d2, dtype=np.int64)
    else:
        lin = np.nonzero(np.asarray(mask, dtype=bool).ravel(order="F"))[0].astype(np.int64)
    i = lin % d1 + 1
    j = lin // d1 + 1
    delta = np.floor(u01(seed, 1, lin) * 3.0).astype(np.int64) - 1
    ti = i + delta
    delta = np.where((ti < 1) | (ti > d1), 0, delta)
    ti = i + delta
    tlin = j * d1 + (ti - 1)  # (ti, j + 1)
    pos = np.searchsorted(lin, tlin)
    posc = np.minimum(pos, len(lin) - 1)
    ok = (j < d2) & (pos < len(lin)) & (lin[posc] == tlin)
    if catchment_length > 0:
        ok &= (j % catchment_length) != 0
    ldd = np.where(ok, 8 + delta, 5).astype(np.uint8)  # (-1,+1) = 7, (0,+1) = 8, (+1,+1) = 9
    down = np.where(ok, posc + 1, 0).astype(np.int64)
    indices = np.stack([i, j], axis=1).astype(np.int64)
    return indices, ldd, down, lin


# PCRaster LDD code of the CartesianIndex offset (di, dj)   (utils.jl:2-12)
_LDD_OF = {(-1, -1): 1, (0, -1): 2, (1, -1): 3, (-1, 0): 4, (1, 0): 6, (-1, 1): 7, (0, 1): 8,
           (1, 1): 9}


def dendritic_ldd(d1: int, d2: int, seed: int, n_active: int | None = None):
    """A single-outlet dendritic D8 network with all eight drain directions: priority-flood
    (Barnes et al. 2014) from one outlet on the raster edge over a smooth random surface (a
    tilted plane + a few low-frequency waves + white noise): every cell drains to the neighbour
    it was flooded from, i.e. along the steepest-descent path of the pit-filled surface. With
    `n_active` the basin is the first n_active cells reached (a connected, basin-shaped mask).
    Cell ids are the column-major ranks of the active cells, so they are unrelated to the
    drainage order. Same return values as scheidegger_ldd."""
    import heapq
    ii, jj = np.meshgrid(np.arange(d1, dtype=np.float64), np.arange(d2, dtype=np.float64),
                         indexing="ij")
    lin_all = (jj * d1 + ii).astype(np.int64)
    ph = u01(seed, 70, np.arange(16))
    z = 0.9 * (d2 - 1 - jj) / max(d2 - 1, 1) + 0.25 * np.abs(ii / max(d1 - 1, 1) - 0.5)
    for k in range(4):
        fx, fy = 1.0 + 2.0 * ph[4 * k], 1.0 + 2.0 * ph[4 * k + 1]
        z += 0.08 / (k + 1) * np.sin(2 * np.pi * (fx * ii / d1 + ph[4 * k + 2])) \
            * np.cos(2 * np.pi * (fy * jj / d2 + ph[4 * k + 3]))
    z += 0.02 * u01(seed, 71, lin_all.ravel()).reshape(d1, d2)
    want = d1 * d2 if n_active is None else int(n_active)
    oi, oj = d1 // 2, d2 - 1                        # the outlet: middle of the last column
    closed = np.zeros((d1, d2), dtype=bool)
    down_ij = np.full((d1, d2, 2), -1, dtype=np.int64)
    heap = [(float(z[oi, oj]), oi, oj)]
    closed[oi, oj] = True
    count = 1
    nbrs = [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]
    while heap and count < want:
        zc, ci, cj = heapq.heappop(heap)
        for di, dj in nbrs:
            ni, nj = ci + di, cj + dj
            if 0 <= ni < d1 and 0 <= nj < d2 and not closed[ni, nj]:
                closed[ni, nj] = True
                down_ij[ni, nj] = (ci, cj)
                heapq.heappush(heap, (max(float(z[ni, nj]), zc), ni, nj))
                count += 1
                if count >= want:
                    break
    lin = np.nonzero(closed.ravel(order="F"))[0].astype(np.int64)
    i = lin % d1
    j = lin // d1
    rank = np.full(d1 * d2, -1, dtype=np.int64)
    rank[lin] = np.arange(len(lin))
    di = down_ij[i, j, 0] - i
    dj = down_ij[i, j, 1] - j
    pit = down_ij[i, j, 0] < 0
    code = np.array([[_LDD_OF.get((a, b), 5) for b in (-1, 0, 1)] for a in (-1, 0, 1)], dtype=np.uint8)
    ldd = np.where(pit, 5, code[np.clip(di, -1, 1) + 1, np.clip(dj, -1, 1) + 1]).astype(np.uint8)
    down = np.where(pit, 0, rank[np.where(pit, 0, down_ij[i, j, 1] * d1 + down_ij[i, j, 0])] + 1)
    indices = np.stack([i + 1, j + 1], axis=1).astype(np.int64)
    return indices, ldd, down.astype(np.int64), lin


def upstream_cells(indices: np.ndarray, down: np.ndarray) -> np.ndarray:
    """Number of cells draining through each cell (itself included): Kahn's algorithm, one
    vectorised sweep per topological level."""
    n = len(down)
    acc = np.ones(n, dtype=np.int64)
    indeg = np.bincount(down[down > 0] - 1, minlength=n)
    front = np.nonzero(indeg == 0)[0]
    while len(front):
        d = down[front]
        m = d > 0
        tgt = d[m] - 1
        np.add.at(acc, tgt, acc[front[m]])
        np.subtract.at(indeg, tgt, 1)
        cand = np.unique(tgt)
        front = cand[indeg[cand] == 0]
    return acc


def set_layerthickness(ref_depth, cum_depth, thickness):
    """utils.jl:390-404, vectorised over cells: (n,), (N+1,) or (n,N+1), (N,) or (n,N)."""
    n = len(ref_depth)
    thickness = np.broadcast_to(thickness, (n, thickness.shape[-1]))
    cum_depth = np.broadcast_to(cum_depth, (n, cum_depth.shape[-1]))
    out = np.full(thickness.shape, np.nan)
    with np.errstate(invalid="ignore"):
        for k in range(thickness.shape[1]):
            full = ref_depth > cum_depth[:, k + 1]
            part = ~full & (ref_depth - cum_depth[:, k] > 0.0)
            out[full, k] = thickness[full, k]
            out[part, k] = (ref_depth - cum_depth[:, k])[part]
    return out


def _kh_layered(profile, F, nlayers, zi, N):
    """kh_layered_profile! (utils.jl:792-895) for the cold state: n_unsatlayers from zi."""
    n = len(zi)
    kv, cld, alt = F["kv"], F["cumulative_layer_depth"], F["actual_layer_thickness"]
    d, ratio = F["soil_thickness"], F["ssf_khfrac"]
    ult = set_layerthickness(zi, cld, alt)
    nunsat = N - np.isnan(ult).sum(axis=1)
    kh = np.zeros(n)
    for i in range(n):
        m = int(nlayers[i])
        if d[i] > zi[i]:
            t = 0.0
            k = max(int(nunsat[i]), 1)
            if profile == 3 and zi[i] >= F["z_layered"][i]:
                f, zl, j = F["hydraulic_conductivity_scale_parameter"][i], F["z_layered"][i], int(F["nlayers_kv"][i])
                t += kv[i, j - 1] / f * (np.exp(-f * (zi[i] - zl)) - np.exp(-f * (d[i] - zl)))
                k = m
            else:
                t += (cld[i, k] - zi[i]) * kv[i, k - 1]
            k += 1
            while k <= m:
                if profile == 3 and k > F["nlayers_kv"][i]:
                    f, zl, j = F["hydraulic_conductivity_scale_parameter"][i], F["z_layered"][i], int(F["nlayers_kv"][i])
                    t += kv[i, j - 1] / f * (1.0 - np.exp(-f * (d[i] - zl)))
                    k = m
                else:
                    t += alt[i, k - 1] * kv[i, k - 1]
                k += 1
            kh[i] = (t / (d[i] - zi[i])) * ratio[i]
        else:
            kh[i] = kv[i, m - 1] * ratio[i]
    return kh


def edge_connectivity(indices: np.ndarray, d1: int, d2: int) -> dict:
    """EdgeConnectivity(network::NetworkLand) (network.jl:136-153), 1-based, n + 1 = no active
    neighbour; the model holds it in domain.land.network.edge_indices."""
    n = len(indices)
    rev = np.zeros((d1 + 2, d2 + 2), dtype=np.int64)
    rev[indices[:, 0], indices[:, 1]] = np.arange(1, n + 1)
    out = {}
    for name, (di, dj) in (("ind_y_down", (0, -1)), ("ind_x_down", (-1, 0)), ("ind_x_up", (1, 0)),
                           ("ind_y_up", (0, 1))):
        r = rev[indices[:, 0] + di, indices[:, 1] + dj]
        out[name] = np.where(r != 0, r, n + 1).astype(np.int64)
    return out


_PCR_DIR = [(-1, -1), (0, -1), (1, -1), (-1, 0), (0, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]


def _set_effective_flowwidth(we_x, we_y, rldd, rdown_land, ridx, river, flow_width, res_outlet,
                             x_down, y_down, n, racc):
    """set_effective_flowwidth! (utils.jl:604-669): the river width is taken off the cell edges
    the river crosses (half of it off two edges for a diagonal direction). rdown_land: downstream
    LAND cell (1-based, 0 = none) of every river cell; x_down / y_down 0-based (n = none)."""
    riv_of_land = np.full(n, -1, dtype=np.int64)
    riv_of_land[ridx] = np.arange(len(ridx))
    sub = lambda a, k, w: max(a[k] - w, 0.0)
    for v in np.argsort(racc, kind="stable"):              # any topological order gives the same result
        dl = rdown_land[v]
        if dl <= 0 or not river[dl - 1]:
            continue
        w = min(flow_width[v], flow_width[riv_of_land[dl - 1]])
        di, dj = _PCR_DIR[int(rldd[v]) - 1]
        idx, res = ridx[v], bool(res_outlet[v])
        xd, yd = x_down[idx], y_down[idx]
        if (di, dj) == (1, 1):
            we_x[idx] = 0.0 if res else sub(we_x, idx, 0.5 * w)
            we_y[idx] = 0.0 if res else sub(we_y, idx, 0.5 * w)
        elif (di, dj) == (-1, -1):
            if xd < n:
                we_y[xd] = 0.0 if res else sub(we_y, xd, 0.5 * w)
            if yd < n:
                we_x[yd] = 0.0 if res else sub(we_x, yd, 0.5 * w)
        elif (di, dj) == (1, 0):
            we_y[idx] = 0.0 if res else sub(we_y, idx, w)
        elif (di, dj) == (0, 1):
            we_x[idx] = 0.0 if res else sub(we_x, idx, w)
        elif (di, dj) == (-1, 0):
            if xd < n:
                we_y[xd] = 0.0 if res else sub(we_y, xd, w)
        elif (di, dj) == (0, -1):
            if yd < n:
                we_x[yd] = 0.0 if res else sub(we_x, yd, w)
        elif (di, dj) == (1, -1):
            we_y[idx] = sub(we_y, idx, 0.5 * w)
            if yd < n:
                we_x[yd] = 0.0 if res else sub(we_x, yd, 0.5 * w)
        elif (di, dj) == (-1, 1):
            if xd < n:
                we_y[xd] = 0.0 if res else sub(we_y, xd, 0.5 * w)
            we_x[idx] = 0.0 if res else sub(we_x, idx, 0.5 * w)


def make_basin(d1: int, d2: int, seed: int = 42, dt: float = 86400.0,
               soil_layer_thickness_mm=(100, 300, 800), mask=None, river_fraction_target=0.116,
               cell_length: float = 1000.0, snow: bool = True, glacier: bool = False,
               kv_profile: int = 0, nthreads: int = 8, adaptive: bool = False,
               soil_infiltration_reduction: bool = False, id_offset: int = 0,
               external_inflow: bool = False, network: str = "scheidegger",
               n_active: int | None = None, n_river: int | None = None, reservoirs: int = 0,
               snow_transport: bool = False, river_routing: int = 0, catchment_length: int = 0,
               floodplain: bool = False, land_routing: int = 0):
    """Returns (cfg, domain, fields). `fields` holds every input array of the hot path under
    the reference's field names; layered arrays are cell-major (n, N). network: "scheidegger"
    (a forest of many small basins, codes 5/7/8/9) or "dendritic" (one outlet, all eight
    directions; n_active cells, n_river river cells by upstream area)."""
    if network == "dendritic":
        indices, ldd, down, lin = dendritic_ldd(d1, d2, seed, n_active)
    else:
        indices, ldd, down, lin = scheidegger_ldd(d1, d2, seed, mask, catchment_length)
    gid = lin + np.int64(id_offset)
    n = len(ldd)
    N = len(soil_layer_thickness_mm) + 1
    acc = upstream_cells(indices, down)
    # river mask: largest upstream areas, fraction ~ Moselle's 5809 / 50063
    if n_river:
        thr = np.sort(acc)[::-1][min(int(n_river), n) - 1]
        river = acc >= max(thr, 2)
    elif river_fraction_target > 0:
        thr = np.quantile(acc, 1.0 - river_fraction_target)
        river = acc >= max(thr, 2)
    else:
        river = np.zeros(n, dtype=bool)
    river_land_indices = np.nonzero(river)[0].astype(np.int64) + 1
    nriv = len(river_land_indices)

    F = {}
    r = lambda stream, mean: _pm20(seed, stream, gid, mean)
    # ---- vegetation / interception / snow (run_sbm.jl:98-142) ------------------------------
    F["leaf_area_index"] = r(10, 1.06)
    F["storage_specific_leaf"] = r(11, 8.94e-5)
    F["storage_wood"] = r(12, 1.96e-4)
    F["light_extinction_coefficient"] = r(13, 0.674)
    F["crop_coefficient"] = np.ones(n)
    F["canopy_gap_fraction"] = np.exp(-F["light_extinction_coefficient"] * F["leaf_area_index"])
    F["maximum_canopy_storage"] = (F["storage_specific_leaf"] * F["leaf_area_index"]
                                   + F["storage_wood"])
    F["evaporation_to_precipitation_ratio"] = r(14, 0.1)
    F["canopy_storage"] = np.zeros(n)
    F["temperature_threshold_snowfall"] = np.full(n, 273.15)
    F["temperature_interval_snowfall"] = np.full(n, 2.0)
    F["temperature_threshold_melt"] = np.full(n, 273.15)
    F["degree_day_factor"] = r(15, 4.348e-8)
    F["water_holding_capacity"] = np.full(n, 0.1)
    F["snow_storage"] = np.zeros(n)
    F["snow_water"] = np.zeros(n)
    if snow_transport:   # a snow pack to move: up to 1.5 m on a third of the cells
        pack = u01(seed, 18, gid) < 0.33
        F["snow_storage"] = np.where(pack, 1.5 * u01(seed, 19, gid), 0.0)
        F["snow_water"] = 0.08 * F["snow_storage"]
    if glacier:
        F["glacier_temperature_threshold_melt"] = np.full(n, 273.15)
        F["glacier_degree_day_factor"] = r(16, 3.0e-3 / 86400.0)
        F["glacier_snow_to_ice_fraction"] = np.full(n, 0.001 / 86400.0)
        F["glacier_fraction"] = np.where(u01(seed, 17, gid) < 0.05, 0.4, 0.0)
        F["glacier_store"] = np.full(n, 5.5)
    # ---- soil (run_sbm.jl:196-324, App. D) ---------------------------------------------------
    theta_s = r(20, 0.4409)
    theta_r = r(21, 0.1657)
    frac_fc = r(22, 0.512)
    theta_fc = theta_r + (theta_s - theta_r) * frac_fc
    soil_thickness = r(23, 1.838)
    F["theta_s"], F["theta_r"], F["theta_fc"] = theta_s, theta_r, theta_fc
    F["soil_thickness"] = soil_thickness
    cfg_thick = np.array([t * 1e-3 for t in soil_layer_thickness_mm] + [np.nan])
    cum_cfg = np.concatenate([[0.0], np.cumsum(cfg_thick)])  # soil.jl:327-332
    alt = set_layerthickness(soil_thickness, cum_cfg, cfg_thick)
    F["actual_layer_thickness"] = alt
    F["cumulative_layer_depth"] = np.concatenate([np.zeros((n, 1)), np.cumsum(alt, axis=1)], axis=1)
    nlayers = (N - np.isnan(alt).sum(axis=1)).astype(np.int64)
    F["number_of_layers"] = nlayers
    bc_means = [9.43, 9.82, 10.24, 10.24, 10.24, 10.24, 10.24, 10.24]
    F["brooks_corey_exponent"] = np.stack([r(30 + k, bc_means[k]) for k in range(N)], axis=1)
    F["vertical_hydraulic_conductivity_factor"] = np.ones((n, N))
    F["air_entry_pressure"] = np.full(n, -0.1)
    F["h1"], F["h2"] = np.zeros(n), np.full(n, -1.0)
    F["h3_high"], F["h3_low"], F["h4"] = np.full(n, -4.0), np.full(n, -10.0), np.full(n, -160.0)
    F["alpha_h1"] = np.ones(n)
    F["w_soil"], F["cf_soil"] = np.full(n, 0.1125), np.full(n, 0.038)
    # mostly Moselle-like (0.013), with 5 % "urban" cells that shed infiltration-excess runoff
    urban = u01(seed, 46, gid) < 0.05
    F["compacted_soil_area_fraction"] = np.where(urban, r(47, 0.6), r(40, 0.013))
    F["infiltration_capacity_compacted_soil"] = r(41, 5.787e-8)
    F["kv_0"] = r(42, 4.846e-6)
    F["infiltration_capacity_soil"] = F["kv_0"] * F["vertical_hydraulic_conductivity_factor"][:, 0]
    F["hydraulic_conductivity_scale_parameter"] = r(43, 3.304)
    if kv_profile == 1:
        F["z_exp"] = r(44, 0.6)
    if kv_profile >= 2:   # layered profiles (soil.jl:272-316): kv per layer, decreasing with depth
        dec = np.array([1.0, 0.55, 0.3, 0.18, 0.12, 0.08, 0.05, 0.03])[:N]
        F["kv"] = np.stack([r(90 + k, 4.846e-6 * dec[k]) for k in range(N)], axis=1)
        F["infiltration_capacity_soil"] = F["kv"][:, 0] * F["vertical_hydraulic_conductivity_factor"][:, 0]
    if kv_profile == 3:   # KvLayeredExponential: z_layered snapped to a layer boundary
        cld_ = F["cumulative_layer_depth"]
        zl = np.full(n, 0.4)
        nk = np.zeros(n, dtype=np.int64)
        for i in range(n):
            layers = cld_[i, 1:nlayers[i]]        # cumulative_layer_depth[i][2:number_of_layers[i]]
            if len(layers) == 0:                  # a single layer: keep the first boundary
                layers = cld_[i, 1:2]
            k = int(np.argmin(np.abs(zl[i] - layers)))
            nk[i] = k + 1
            zl[i] = layers[k]
        F["z_layered"] = zl
        F["nlayers_kv"] = nk
    F["maximum_leakage"] = np.zeros(n)
    F["cap_hmax"], F["cap_n"] = np.full(n, 2.0), np.full(n, 2.0)
    F["wet_root_distribution_parameter"] = np.full(n, -5.0e5)
    rd = np.minimum(soil_thickness * 0.99, r(45, 0.370))  # sbm.jl:57-58
    F["rooting_depth"] = rd
    # default root fraction (soil.jl:537-563)
    cld = F["cumulative_layer_depth"]
    rootfraction = np.zeros((n, N))
    with np.errstate(invalid="ignore"):
        for k in range(N):
            full = (rd - cld[:, k]) >= alt[:, k]
            part = np.maximum((rd - cld[:, k]) / rd, 0.0)
            rootfraction[:, k] = np.where(full, alt[:, k] / rd, part)
    rootfraction[rd <= 0.0] = 0.0
    F["rootfraction"] = np.nan_to_num(rootfraction, nan=0.0)
    F["soil_water_capacity"] = soil_thickness * (theta_s - theta_r)
    # cold states (soil.jl:164-198)
    F["saturated_water_depth"] = 0.85 * F["soil_water_capacity"]
    F["unsaturated_layer_depth"] = np.zeros((n, N))
    F["soil_surface_temperature"] = np.full(n, 10.0 + 273.15)
    zi = np.maximum(0.0, soil_thickness - F["saturated_water_depth"] / (theta_s - theta_r))
    F["water_table_depth"] = zi
    # ---- shared land parameters (domain.jl:218-261, utils.jl:418-466) ------------------------
    xl = yl = cell_length
    area = np.full(n, xl * yl)
    diag = ~np.isin(ldd, (2, 8, 4, 6))
    F["area"] = area
    F["flow_length"] = np.where(diag, np.hypot(xl, yl), yl)
    F["flow_width"] = np.where(diag, (xl * yl) / np.hypot(xl, yl), xl)
    F["slope"] = np.maximum(r(50, 0.0947), 1e-5)
    riv_len_land = _pm20(seed, 51, gid, cell_length) * 0.5 + 0.5 * cell_length
    riv_wid_land = np.maximum(2.0, 0.8 * np.sqrt(acc.astype(np.float64)))
    F["river_fraction"] = np.where(river, np.minimum(riv_len_land * riv_wid_land / area, 1.0), 0.0)
    # open water (lakes, ponds) on 30 % of the cells: land runoff feeding the overland wave
    wf = np.where(u01(seed, 55, gid) < 0.3, 0.05 * u01(seed, 56, gid), 0.0)
    F["water_fraction"] = np.maximum(wf - F["river_fraction"], 0.0)  # domain.jl:296
    land_area = (1.0 - F["river_fraction"]) * area
    F["surface_flow_width"] = np.where(river, land_area / F["flow_length"], F["flow_width"])
    # get_flow_fraction_to_river (utils.jl:493-510)
    f2r = np.zeros(n)
    has_down = down > 0
    dn = np.where(has_down, down - 1, 0)
    sel = has_down & river[dn] & (ldd != ldd[dn])
    f2r[sel] = F["slope"][sel] / (F["slope"][dn[sel]] + F["slope"][sel])
    F["flow_fraction_to_river"] = f2r
    # ---- overland flow (surface_kinwave.jl:32-44,188-219) ------------------------------------
    mannings_land = r(52, 0.484)
    F["olf_alpha"] = (jl_pow(mannings_land / np.sqrt(F["slope"]), 0.6)
                      * jl_pow(F["surface_flow_width"], (2.0 / 3.0) * 0.6))
    for k in ("olf_q", "olf_h", "olf_storage", "olf_qin", "olf_qlat", "olf_inwater"):
        F[k] = np.zeros(n)
    # ---- lateral subsurface flow (lateral_subsurface_flow.jl:84-141, utils.jl:916-937) -------
    khfrac = r(53, 100.0)
    F["kh_0"] = khfrac * F["kv_0"]
    F["ssf_soil_thickness"] = soil_thickness.copy()
    F["specific_yield"] = np.maximum(theta_s - theta_fc, 0.02)
    F["ssf_top"] = 100.0 + 50.0 * u01(seed, 54, gid)
    fpar = F["hydraulic_conductivity_scale_parameter"]
    if kv_profile >= 2:   # initialize_lateral_ssf_model!(..., ::KvLayered / ::KvLayeredExponential)
        F["ssf_khfrac"] = khfrac                                        # utils.jl:988-1050
        F["ssf_kh"] = _kh_layered(kv_profile, F, nlayers, zi, N)
        F["ssf_q"] = F["ssf_kh"] * (soil_thickness - zi) * F["slope"] * F["flow_width"]
        kh_max = np.zeros(n)
        for i in range(n):
            acc_ = 0.0
            for j in range(1, nlayers[i] + 1):
                if kv_profile == 2 or j <= F["nlayers_kv"][i]:
                    acc_ += F["kv"][i, j - 1] * alt[i, j - 1]
                else:
                    zt = soil_thickness[i] - F["z_layered"][i]
                    k = max(j - 1, 1)
                    acc_ += F["kv"][i, k - 1] / fpar[i] * (1.0 - np.exp(-fpar[i] * zt))
                    break
            kh_max[i] = acc_ * khfrac[i]
        F["ssf_q_max"] = kh_max * F["slope"]
    elif kv_profile == 0:
        F["ssf_q_max"] = ((F["kh_0"] * F["slope"]) / fpar) * (1.0 - np.exp(-fpar * soil_thickness))
        F["ssf_q"] = (((F["kh_0"] * F["slope"]) / fpar)
                      * (np.exp(-fpar * zi) - np.exp(-fpar * soil_thickness)) * F["flow_width"])
    else:  # exponential_constant (utils.jl:940-985)
        z_exp = F["z_exp"]
        ssf_constant = F["kh_0"] * np.exp(-fpar * z_exp) * F["slope"] * (soil_thickness - z_exp)
        F["ssf_q_max"] = (((F["kh_0"] * F["slope"]) / fpar) * (1.0 - np.exp(-fpar * z_exp))
                          + ssf_constant)
        q_exp = (((F["kh_0"] * F["slope"]) / fpar) * (np.exp(-fpar * zi) - np.exp(-fpar * z_exp))
                 + ssf_constant)
        q_const = F["kh_0"] * np.exp(-fpar * zi) * F["slope"] * (soil_thickness - zi)
        F["ssf_q"] = np.where(zi < z_exp, q_exp, q_const) * F["flow_width"]
    F["ssf_water_table_depth"] = zi.copy()
    F["ssf_head"] = F["ssf_top"] - zi
    F["ssf_storage"] = F["specific_yield"] * (soil_thickness - zi) * area
    # ---- river (surface_kinwave.jl:67-90, domain.jl:263-277) ---------------------------------
    ridx = river_land_indices - 1
    rg = gid[ridx]
    F["riv_flow_length"] = riv_len_land[ridx]
    F["riv_flow_width"] = riv_wid_land[ridx]
    riv_slope = np.maximum(_pm20(seed, 60, rg, 0.0031), 1e-5)
    riv_n = _pm20(seed, 61, rg, 0.0305)
    bankfull = _pm20(seed, 62, rg, 1.10)
    F["riv_alpha"] = (jl_pow(riv_n / np.sqrt(riv_slope), 0.6)
                      * jl_pow(F["riv_flow_width"] + bankfull, (2.0 / 3.0) * 0.6))
    F["riv_external_inflow"] = np.zeros(nriv)
    if external_inflow and nriv:
        F["riv_external_inflow"] = np.where(u01(seed, 63, rg) < 0.02, -0.05, 0.0)
    F["riv_abstraction"] = np.zeros(nriv)
    for k in ("riv_q", "riv_h", "riv_storage", "riv_qin", "riv_qlat", "riv_inwater"):
        F[k] = np.zeros(nriv)

    # ---- local-inertial river flow: the staggered grid (surface_staggered_scheme.jl:36-156) ------
    if river_routing == 1 and nriv:
        rdown = np.zeros(nriv, dtype=np.int64)          # downstream river node (1-based, 0 = pit)
        riv_of_land = np.zeros(n, dtype=np.int64)
        riv_of_land[ridx] = np.arange(1, nriv + 1)
        dl = down[ridx]
        rdown[dl > 0] = riv_of_land[dl[dl > 0] - 1]
        L, W = F["riv_flow_length"], F["riv_flow_width"]
        # bed level: falls along the network with the river slope, 10 m at the pits
        zb = np.full(nriv, 10.0)
        order = np.argsort(-acc[ridx], kind="stable")   # downstream nodes first
        for r in order:
            if rdown[r] > 0:
                zb[r] = zb[rdown[r] - 1] + riv_slope[r] * L[r]
        d = np.where(rdown > 0, rdown - 1, np.arange(nriv))   # ghost node copies the pit
        F["li_zb"] = zb
        F["li_zb_at_edge"] = np.maximum(zb, zb[d])
        F["li_flow_width_at_edge"] = np.minimum(W, W[d])
        F["li_flow_length_at_edge"] = (L + L[d]) / 2.0          # Statistics.mean of the pair
        n_at_edge = (riv_n[d] * L[d] + riv_n * L) / (L[d] + L)
        F["li_mannings_n_sq_at_edge"] = n_at_edge * n_at_edge
        F["li_ghost_h"] = np.zeros(nriv)                        # riverdepth_bc
        F["li_error"] = np.zeros(nriv)
    if floodplain and nriv:
        # 1-D floodplain (floodplain.jl:50-147): a six-level profile per node whose first
        # width is the channel's and whose widths grow with the depth; flow area, wetted
        # perimeter and storage are its cumulative sums. A shallow bankfull depth so that
        # the synthetic rivers do go over bank.
        fp_depth = np.array([0.0, 0.5, 1.0, 1.5, 2.0, 2.5])
        P = len(fp_depth)
        bankfull_depth = 0.02 + 0.06 * u01(seed, 95, rg)
        grow = 1.5 + 2.0 * u01(seed, 96, rg)               # widening per level
        width = np.empty((nriv, P))
        width[:, 0] = F["riv_flow_width"]
        for l in range(1, P):
            width[:, l] = width[:, l - 1] * (1.0 + (grow - 1.0) / l)
        dh = np.diff(fp_depth)
        area = np.zeros((nriv, P))
        perim = np.empty((nriv, P))
        perim[:, 0] = F["riv_flow_width"] + 2.0 * bankfull_depth
        perim[:, 1] = perim[:, 0] + 2.0 * dh[0]
        for l in range(1, P):
            area[:, l] = area[:, l - 1] + width[:, l] * dh[l - 1]
            if l > 1:
                perim[:, l] = perim[:, l - 1] + (width[:, l] - width[:, l - 1]) + 2.0 * dh[l - 1]
        F["fp_profile_width"] = width
        F["fp_profile_flow_area"] = area
        F["fp_profile_wetted_perimeter"] = perim
        F["fp_profile_storage"] = area * F["riv_flow_length"][:, None]
        F["li_bankfull_depth"] = bankfull_depth
        F["li_bankfull_storage"] = bankfull_depth * F["riv_flow_width"] * F["riv_flow_length"]
        n_fp = 0.072
        if river_routing == 1:      # local inertial: edge parameters (floodplain.jl:164-215)
            zb_fp = F["li_zb"] + bankfull_depth
            rd = np.where(rdown > 0, rdown - 1, np.arange(nriv))
            F["fp_zb_at_edge"] = np.maximum(zb_fp, zb_fp[rd])
            F["fp_mannings_n_sq_at_edge"] = np.full(nriv, n_fp * n_fp)
        else:                       # kinematic wave: Manning flow capacity (floodplain.jl:238-246)
            F["fp_mannings_n"] = np.full(nriv, n_fp)
            F["fp_slope"] = riv_slope.copy()
        for k in ("fp_h", "fp_storage", "fp_q", "fp_error"):
            F[k] = np.zeros(nriv)
    # ---- reservoirs on the river (reservoir.jl; Moselle has two, test/sbm_config.toml:126) -----
    reservoir_river_indices = np.zeros(0, dtype=np.int64)
    nres = 0
    if reservoirs and nriv:
        racc = acc[ridx]
        has_down_river = river[np.maximum(down[ridx], 1) - 1] & (down[ridx] > 0)
        # mid-sized river nodes with a downstream river node, one per downstream node
        cand = np.nonzero(has_down_river & (racc >= np.quantile(racc, 0.5)))[0]
        pick, used = [], set()
        for c in cand[np.argsort(u01(seed, 80, rg[cand]))]:
            if int(down[ridx][c]) not in used:
                pick.append(int(c))
                used.add(int(down[ridx][c]))
            if len(pick) == reservoirs:
                break
        reservoir_river_indices = np.sort(np.array(pick, dtype=np.int64)) + 1
        nres = len(reservoir_river_indices)
        k = np.arange(nres)
        rid = rg[reservoir_river_indices - 1]
        F["res_area"] = 1.0e6 * (0.5 + u01(seed, 81, rid))
        # outflow types cycle: simple, modified_puls, free_weir, simple + observed outflow
        typ = np.array([4.0, 3.0, 2.0, 4.0])[k % 4]
        F["res_outflow_curve_type"] = typ
        F["res_maximum_storage"] = F["res_area"] * 40.0
        F["res_threshold"] = np.where(typ == 2.0, 9.5, 0.0)
        F["res_rating_curve_coefficient"] = np.where(typ == 3.0, 8.0, 12.0) * (0.8 + 0.4 * u01(seed, 82, rid))
        F["res_rating_curve_exponent"] = np.full(nres, 1.5)
        F["res_maximum_release"] = np.full(nres, 24.0)
        F["res_demand"] = 0.5 + u01(seed, 83, rid)
        F["res_target_minimum_fraction"] = np.full(nres, 0.075)
        F["res_target_full_fraction"] = np.full(nres, 0.75)
        F["res_waterlevel"] = np.where(typ == 4.0, 30.0, 10.0)
        F["res_storage"] = F["res_area"] * F["res_waterlevel"]   # initialize_storage, linear
        F["res_outflow_obs"] = np.where(k % 4 == 3, 0.8, np.nan)
        F["res_external_inflow"] = np.where(k % 2 == 0, -0.1, 0.05)
        F["res_precipitation"] = np.full(nres, 2.0e-8)
        F["res_evaporation"] = np.full(nres, 1.0e-8)
        for name in ("res_inflow_overland", "res_inflow_subsurface", "res_outflow"):
            F[name] = np.zeros(nres)

    # ---- 2-D local-inertial overland flow coupled to the local-inertial river ------------------
    # (LocalInertialOverlandFlowParameters surface_staggered_scheme.jl:894-961; the host prepares
    # the edge parameters like the Julia model does: EdgeConnectivity network.jl:136-153,
    # set_effective_flowwidth! utils.jl:604-669)
    edges = None
    if land_routing == 1:
        assert river_routing == 1 and not floodplain, "land_routing = 1 needs river_routing = 1, no 1-D floodplain"
        edges = edge_connectivity(indices, d1, d2)
        x_up, y_up = edges["ind_x_up"] - 1, edges["ind_y_up"] - 1       # 0-based, n = none
        x_down, y_down = edges["ind_x_down"] - 1, edges["ind_y_down"] - 1
        x_len = cell_length * (0.9 + 0.2 * u01(seed, 100, gid))
        y_len = F["area"] / x_len
        # ground elevation: falls along the drainage paths (2 per mille), pits at 10 .. 12 m
        z = 10.0 + 2.0 * u01(seed, 101, gid)
        rise = 0.002 * F["flow_length"] * (0.5 + u01(seed, 102, gid))
        for v in np.argsort(-acc, kind="stable"):                   # downstream cells first
            if down[v] > 0:
                z[v] = z[down[v] - 1] + rise[v]
        zx_max, zy_max = np.zeros(n), np.zeros(n)
        sx, sy = x_up < n, y_up < n
        zx_max[sx] = np.maximum(z[sx], z[x_up[sx]])
        zy_max[sy] = np.maximum(z[sy], z[y_up[sy]])
        mann = _pm20(seed, 103, gid, 0.072)
        we_x, we_y = x_len.copy(), y_len.copy()
        res_outlet = np.zeros(nriv, dtype=bool)
        if nres:
            res_outlet[reservoir_river_indices - 1] = True
        if nriv:
            _set_effective_flowwidth(we_x, we_y, ldd[ridx], down[ridx], ridx, river,
                                     F["riv_flow_width"], res_outlet, x_down, y_down, n, acc[ridx])
            bankfull_depth = 0.05 + 0.5 * u01(seed, 104, rg)
            F["li_bankfull_depth"] = bankfull_depth
            F["li_bankfull_storage"] = bankfull_depth * F["riv_flow_width"] * F["riv_flow_length"]
            zb = z[ridx] - bankfull_depth                           # bankfull elevation - depth
            rd = np.where(rdown > 0, rdown - 1, np.arange(nriv))
            F["li_zb"] = zb
            F["li_zb_at_edge"] = np.maximum(zb, zb[rd])
        F["li_land_xwidth_at_edge"], F["li_land_ywidth_at_edge"] = we_x, we_y
        F["li_land_zx_max_at_edge"], F["li_land_zy_max_at_edge"] = zx_max, zy_max
        F["li_land_mannings_n_sq_at_edge"] = mann * mann
        F["li_land_z"] = z
        F["li_land_x_length"], F["li_land_y_length"] = x_len, y_len
        for k in ("li_land_qx", "li_land_qy", "li_land_qx0", "li_land_qy0", "li_land_error",
                  "li_land_runoff"):
            F[k] = np.zeros(n)

    cfg = dict(n=n, nriv=nriv, nres=nres, n_layers=N, N=N, gash=int(dt >= 23 * 3600.0), has_lai=1,
               snow=int(snow), glacier=int(glacier),
               soil_infiltration_reduction=int(soil_infiltration_reduction),
               kv_profile=kv_profile, adaptive=int(adaptive), nthreads=nthreads,
               snow_transport=int(snow_transport), river_routing=int(river_routing),
               li_froude_limit=1, li_ghost_nodes=1, li_alpha=0.7, li_h_thresh=1.0e-3,
               land_streamorder_min=5, river_streamorder_min=6, dt_land=3600.0, dt_river=900.0,
               dt_ssf=86400.0, ssf_alpha_coefficient=1.0, dt=dt,
               kin_wave_min_flow_qroot=1e-30 ** 0.2)
    if floodplain and nriv:
        cfg["fp_depth"] = [0.0, 0.5, 1.0, 1.5, 2.0, 2.5]
    domain = dict(d1=d1, d2=d2, indices=indices, ldd=ldd, river_land_indices=river_land_indices,
                  down=down, gid=gid, upstream_cells=acc,
                  reservoir_river_indices=reservoir_river_indices)
    if land_routing == 1:
        cfg.update(land_routing=1, li_land_alpha=0.7, li_land_theta=0.9, li_land_h_thresh=1.0e-3,
                   li_land_froude_limit=1)
        domain["edges"] = edges
    return cfg, domain, F


def make_forcing(seed: int, step: int, gid: np.ndarray, dt: float = 86400.0):
    """Forcing of one model step in SI (m s-1, m s-1, K): P = 0 w.p. 0.6 else Exp(3 mm d-1);
    PET ~ U(0.2, 1.2) mm d-1; T ~ N(275.5, 4) K (Box-Muller)."""
    s = 1000 + 8 * step
    wet = u01(seed, s, gid) >= 0.6
    p_mm_day = np.where(wet, -3.0 * np.log1p(-u01(seed, s + 1, gid)), 0.0)
    pet_mm_day = 0.2 + u01(seed, s + 2, gid)
    u1 = np.maximum(u01(seed, s + 3, gid), 1e-300)
    u2 = u01(seed, s + 4, gid)
    temp = 275.5 + 4.0 * np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    mm_per_day = (1.0 / 86400.0) * 1e-3
    return p_mm_day * mm_per_day, pet_mm_day * mm_per_day, temp
