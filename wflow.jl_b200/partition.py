"""Multi-GPU sharding of one model domain by whole drainage basins.

The vertical update is independent per cell and the kinematic wave is independent per
drainage basin (a connected component of the D8 forest = one pit), which is also how the
reference splits its work between threads (subdomains.jl:177-199). A shard therefore gets
WHOLE basins: no drainage edge is cut, no flux crosses a rank boundary and the hot path needs
no data-path collective -- each rank runs its own `SbmModel` on its own GPU.

`partition_basins` assigns basins to ranks by greedy longest-processing-time bin packing on
a work estimate (cells + river cells weighted by their larger number of internal sub-steps);
`shard_domain` / `shard_fields` cut the host-side arrays; `Shard.cells` / `Shard.river_cells`
are the global (0-based) ids a rank owns, in ascending order, so `out[shard.cells] = local`
gathers results back. All arrays follow the conventions of `SbmModel` (1-based indices on the
ABI, column-major CartesianIndex pairs).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Shard:
    rank: int
    cells: np.ndarray          # global 0-based land ids owned by the rank, ascending
    river_cells: np.ndarray    # global 0-based positions in river_land_indices, ascending
    basins: np.ndarray         # pit ids (global 0-based land ids) of the rank's basins
    weight: float


def downstream_ids(domain: dict) -> np.ndarray:
    """1-based downstream land id per cell (0 = pit) from the gridded LDD, as flowgraph does
    (routing/utils.jl:6-33): PCRaster codes 1..9, 5 = pit, a cell draining out of the active
    domain becomes a pit."""
    if "down" in domain:
        return np.asarray(domain["down"], dtype=np.int64)
    idx = np.asarray(domain["indices"], dtype=np.int64)
    ldd = np.asarray(domain["ldd"], dtype=np.int64)
    d1, d2 = int(domain["d1"]), int(domain["d2"])
    rev = np.zeros((d1, d2), dtype=np.int64)
    rev[idx[:, 0] - 1, idx[:, 1] - 1] = np.arange(1, len(ldd) + 1)
    # PCR_DIR (utils.jl:2-12): CartesianIndex offsets of the LDD codes 1..9
    drow = np.array([0, -1, 0, 1, -1, 0, 1, -1, 0, 1])
    dcol = np.array([0, -1, -1, -1, 0, 0, 0, 1, 1, 1])
    r = idx[:, 0] - 1 + drow[ldd]
    c = idx[:, 1] - 1 + dcol[ldd]
    ok = (ldd != 5) & (r >= 0) & (r < d1) & (c >= 0) & (c < d2)
    down = np.zeros(len(ldd), dtype=np.int64)
    down[ok] = rev[r[ok], c[ok]]
    return down


def basin_of_cells(down: np.ndarray) -> np.ndarray:
    """Pit (0-based id) that every cell drains to, by pointer jumping (O(n log depth))."""
    n = len(down)
    b = np.where(down > 0, down - 1, np.arange(n))
    while True:
        nb = b[b]
        if np.array_equal(nb, b):
            return b
        b = nb


def partition_basins(domain: dict, world: int, river_weight: float = 3.0) -> list[Shard]:
    down = downstream_ids(domain)
    n = len(down)
    basin = basin_of_cells(down)
    rli = np.asarray(domain["river_land_indices"], dtype=np.int64) - 1
    w = np.ones(n)
    w[rli] += river_weight
    pits, inv = np.unique(basin, return_inverse=True)
    bw = np.bincount(inv, weights=w)
    order = np.lexsort((pits, -bw))            # heaviest first, ties by pit id (deterministic)
    load = np.zeros(world)
    owner = np.empty(len(pits), dtype=np.int64)
    for k in order:
        r = int(np.argmin(load))               # lowest rank wins ties
        owner[k] = r
        load[r] += bw[k]
    cell_owner = owner[inv]
    is_river = np.zeros(n, dtype=bool)
    is_river[rli] = True
    shards = []
    for r in range(world):
        cells = np.nonzero(cell_owner == r)[0]
        riv = np.nonzero(cell_owner[rli] == r)[0]
        shards.append(Shard(r, cells, riv, pits[owner == r], float(load[r])))
    return shards


def shard_domain(domain: dict, shard: Shard) -> dict:
    """The rank's own domain: same raster, its cells only (column-major order is kept because
    `cells` is ascending)."""
    cells = shard.cells
    local_of = np.zeros(len(domain["ldd"]), dtype=np.int64)
    local_of[cells] = np.arange(1, len(cells) + 1)
    rli = np.asarray(domain["river_land_indices"], dtype=np.int64)
    out = dict(d1=domain["d1"], d2=domain["d2"],
               indices=np.ascontiguousarray(np.asarray(domain["indices"])[cells]),
               ldd=np.ascontiguousarray(np.asarray(domain["ldd"])[cells]),
               river_land_indices=local_of[rli[shard.river_cells] - 1])
    assert np.all(out["river_land_indices"] > 0)
    down = downstream_ids(domain)[cells]
    assert np.all(local_of[down[down > 0] - 1] > 0), "a drainage edge leaves the shard"
    out["down"] = np.where(down > 0, local_of[np.maximum(down, 1) - 1], 0)
    out["reservoir_river_indices"] = np.zeros(0, dtype=np.int64)
    for k in ("gid", "upstream_cells"):
        if k in domain:
            out[k] = np.asarray(domain[k])[cells]
    return out


def shard_fields(fields: dict, table: dict, shard: Shard) -> dict:
    """`table`: field name -> kind (0 land, 1/2 land layered, 3 river), e.g.
    dict(_lib.field_table()); integer land fields are cut like kind 0."""
    out = {}
    for name, a in fields.items():
        if a is None:
            continue
        kind = table.get(name, 0)
        a = np.asarray(a)
        out[name] = np.ascontiguousarray(a[shard.river_cells] if kind == 3 else a[shard.cells])
    return out


def shard_config(cfg: dict, shard: Shard) -> dict:
    """The shard's configuration. With adaptive internal time steps a shard cannot be advanced on
    its own: the sub-step length is a statistic of the WHOLE domain (surface_kinwave.jl:674-704,
    lateral_subsurface_flow.jl:314-344). `sharded` makes SbmModel.update_model refuse to run until
    a communicator is attached (comm_init_nccl, ShardGroup)."""
    if cfg.get("nres", 0):
        raise ValueError("sharding a domain with reservoirs is not implemented")
    c = dict(cfg)
    c["n"] = int(len(shard.cells))
    c["nriv"] = int(len(shard.river_cells))
    c["sharded"] = True
    return c
