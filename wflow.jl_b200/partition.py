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
        out[name] = np.ascontiguousarray(a[shard.river_cells] if kind in (3, 5) else a[shard.cells])
    return out


def shard_config(cfg: dict, shard: Shard) -> dict:
    """The shard's configuration. With adaptive internal time steps a shard cannot be advanced on
    its own: the sub-step length is a statistic of the WHOLE domain (surface_kinwave.jl:674-704,
    lateral_subsurface_flow.jl:314-344). `sharded` makes SbmModel.update_model refuse to run until
    a communicator is attached (comm_init_nccl, ShardGroup)."""
    if cfg.get("nres", 0):
        raise ValueError("sharding a domain with reservoirs is not implemented")
    if cfg.get("river_routing", 0) or cfg.get("land_routing", 0):
        # the staggered schemes take ONE sub-step length for the whole domain (the minimum Courant
        # step, surface_staggered_scheme.jl:1004-1043) and the 2-D overland flow crosses basin
        # divides: basin-aligned shards are not independent
        raise ValueError("sharding a domain with local-inertial routing is not implemented")
    c = dict(cfg)
    c["n"] = int(len(shard.cells))
    c["nriv"] = int(len(shard.river_cells))
    c["sharded"] = True
    return c


# ---------------------------------------------------------------------------------------------
# Cutting ONE basin across GPUs (second half). When a basin is larger than a GPU (or
# than its fair share of the work) it is cut at confluences, like the reference cuts its basins
# into sub-domains for its threads (subdomains.jl:169-255): a part owns whole upstream subtrees,
# discharge crosses the cut on CUT EDGES. On the device a cut edge is one more inlet slot of the
# consuming chunk, written by the producing GPU straight through NVLink (wflowb200_exchange_*),
# so the parts run as one skewed wavefront; nothing here touches the data path.

def bfs_levels(down: np.ndarray):
    """(order, dist): nodes in downstream-to-upstream order (pits first) and their distance to
    the pit; vectorised level by level."""
    n = len(down)
    src = np.nonzero(down > 0)[0]
    by_dst = src[np.argsort(down[src] - 1, kind="stable")]
    cnt = np.bincount(down[src] - 1, minlength=n)
    ptr = np.concatenate([[0], np.cumsum(cnt)])
    dist = np.zeros(n, dtype=np.int64)
    frontier = np.nonzero(down == 0)[0]
    order = [frontier]
    d = 0
    while len(frontier):
        d += 1
        lens = cnt[frontier]
        if lens.sum() == 0:
            break
        starts = np.repeat(ptr[frontier], lens)
        offs = np.arange(lens.sum()) - np.repeat(np.cumsum(lens) - lens, lens)
        frontier = by_dst[starts + offs]
        dist[frontier] = d
        order.append(frontier)
    return np.concatenate(order), dist


def split_by_subtrees(down: np.ndarray, parts: int, weights: np.ndarray | None = None) -> np.ndarray:
    """Part of every cell: parts - 1 times the upstream subtree whose (still unassigned) weight is
    closest to the fair share is carved off; part 0 keeps the trunk. Every cut edge leads from a
    carved subtree DOWN into what was left, so the parts form a DAG with part 0 at the bottom."""
    n = len(down)
    w = np.ones(n) if weights is None else np.asarray(weights, dtype=float)
    order, _ = bfs_levels(down)
    owner = np.zeros(n, dtype=np.int64)
    target = w.sum() / parts
    for p in range(1, parts):
        free = owner == 0
        acc = np.where(free, w, 0.0)
        for v in order[::-1]:            # upstream first (small basins: the tests and the bench)
            if down[v] > 0 and free[v] and free[down[v] - 1]:
                acc[down[v] - 1] += acc[v]
        cand = np.nonzero(free & (down > 0))[0]
        if len(cand) == 0:
            break
        root = cand[np.argmin(np.abs(acc[cand] - target))]
        take = np.zeros(n, dtype=bool)
        take[root] = True
        for v in order:                  # downstream first: a cell follows its downstream cell
            if down[v] > 0 and free[v] and take[down[v] - 1]:
                take[v] = True
        owner[take] = p
    return owner


def _cut_one_graph(down: np.ndarray, owner: np.ndarray, parts: int):
    """Cut edges of one drainage graph (1-based `down`, 0 = pit). Per part: node ids (global,
    ascending), import (dst local 1-based, pos, global source) and export (src local 1-based,
    peer, index of the peer's import) lists."""
    n = len(down)
    src = np.nonzero(down > 0)[0]
    dst = down[src] - 1
    # position of every edge among the in-edges of its destination, ascending source id
    o = np.lexsort((src, dst))
    pos = np.zeros(len(src), dtype=np.int64)
    ds = dst[o]
    first = np.r_[True, ds[1:] != ds[:-1]]
    run_start = np.maximum.accumulate(np.where(first, np.arange(len(o)), 0))
    pos[o] = np.arange(len(o)) - run_start
    cut = owner[src] != owner[dst]
    local = np.zeros(n, dtype=np.int64)
    nodes = []
    for p in range(parts):
        ids = np.nonzero(owner == p)[0]
        local[ids] = np.arange(1, len(ids) + 1)
        nodes.append(ids)
    out = []
    import_index = {}
    for p in range(parts):
        e = np.nonzero(cut & (owner[dst] == p))[0]
        e = e[np.lexsort((pos[e], dst[e]))]
        for k, ee in enumerate(e):
            import_index[int(src[ee])] = (p, k)
        out.append(dict(nodes=nodes[p], import_dst=local[dst[e]], import_pos=pos[e],
                        import_src_global=src[e]))
    for p in range(parts):
        e = np.nonzero(cut & (owner[src] == p))[0]
        e = e[np.argsort(src[e])]
        out[p]["export_src"] = local[src[e]]
        out[p]["export_src_global"] = src[e]
        out[p]["export_to"] = [import_index[int(u)] for u in src[e]]
    return out


def cut_basin(domain: dict, owner: np.ndarray, parts: int) -> list[dict]:
    """Per part: `shard` (Shard), `domain` (for SbmModel, with the cut-edge arrays) and the
    `links` of its exports: links[domain_id][export] = (peer part, import index of the peer)."""
    down = downstream_ids(domain)
    n = len(down)
    owner = np.asarray(owner, dtype=np.int64)
    rli = np.asarray(domain["river_land_indices"], dtype=np.int64) - 1
    riv_of_land = np.full(n, -1, dtype=np.int64)
    riv_of_land[rli] = np.arange(len(rli))
    dl = down[rli]                      # downstream land cell of every river node
    rdown = np.where(dl > 0, riv_of_land[np.maximum(dl, 1) - 1] + 1, 0)
    land = _cut_one_graph(down, owner, parts)
    river = _cut_one_graph(rdown, owner[rli], parts)
    ldd = np.asarray(domain["ldd"]).copy()
    plans = []
    for p in range(parts):
        sh = Shard(p, land[p]["nodes"], river[p]["nodes"], np.zeros(0, dtype=np.int64),
                   float(len(land[p]["nodes"])))
        local_of = np.zeros(n, dtype=np.int64)
        local_of[sh.cells] = np.arange(1, len(sh.cells) + 1)
        l = ldd[sh.cells].copy()
        l[land[p]["export_src"] - 1] = 5        # the upstream end of a cut edge is a pit here
        dn = down[sh.cells].copy()
        dn[land[p]["export_src"] - 1] = 0
        d = dict(d1=domain["d1"], d2=domain["d2"],
                 indices=np.ascontiguousarray(np.asarray(domain["indices"])[sh.cells]),
                 ldd=np.ascontiguousarray(l),
                 river_land_indices=local_of[rli[sh.river_cells]] ,
                 down=np.where(dn > 0, local_of[np.maximum(dn, 1) - 1], 0),
                 reservoir_river_indices=np.zeros(0, dtype=np.int64),
                 land_import_dst=land[p]["import_dst"], land_import_pos=land[p]["import_pos"],
                 land_export_src=land[p]["export_src"],
                 river_import_dst=river[p]["import_dst"], river_import_pos=river[p]["import_pos"],
                 river_export_src=river[p]["export_src"])
        assert np.all(d["river_land_indices"] > 0)
        for k in ("gid", "upstream_cells"):
            if k in domain:
                d[k] = np.asarray(domain[k])[sh.cells]
        plans.append(dict(shard=sh, domain=d, links=[land[p]["export_to"], river[p]["export_to"]]))
    return plans


def connect_in_process(models: list, plans: list[dict], dt: float) -> None:
    """Handles of ONE process (one per GPU, or several on one GPU): exchange the import buffers as
    device pointers and bind every export."""
    bufs = [m.exchange_prepare(dt) for m in models]
    for p, (m, plan) in enumerate(zip(models, plans)):
        peers = sorted({q for links in plan["links"] for (q, _) in links})
        for q in peers:
            m.exchange_open_peer(q, models[q].n_imports, device_ptr=bufs[q][0],
                                 peer_device=models[q].device)
        for dom_id, links in enumerate(plan["links"]):
            for e, (q, k) in enumerate(links):
                m.exchange_bind(dom_id, e, q, k)


def connect_distributed(model, plan: dict, dt: float, dist) -> None:
    """One process per GPU: the import buffers travel as CUDA IPC handles (torch.distributed
    all_gather_object)."""
    ptr, ipc, nbytes = model.exchange_prepare(dt)
    mine = dict(ipc=ipc, n_imports=model.n_imports, bytes=nbytes)
    every = [None] * dist.get_world_size()
    dist.all_gather_object(every, mine)
    peers = sorted({q for links in plan["links"] for (q, _) in links})
    for q in peers:
        model.exchange_open_peer(q, every[q]["n_imports"], ipc_handle=every[q]["ipc"])
    for dom_id, links in enumerate(plan["links"]):
        for e, (q, k) in enumerate(links):
            model.exchange_bind(dom_id, e, q, k)
