"""
network.py -- CPU ORACLE (test infrastructure): indexing artefacts of the D8 drainage network.

TEST INFRASTRUCTURE ONLY -- nothing in the product imports this. It restates, in plain
Python/numpy, the reference's integer artefacts that must be reproduced BIT-EXACTLY:

  * active_indices                 Wflow/src/utils.jl:85-99
  * flowgraph                      Wflow/src/routing/utils.jl:6-33
  * topological_sort_by_dfs        Graphs.jl 1.14.0 (un-vendored dependency, Manifest.toml:797-801):
                                   iterative DFS over vertices 1..n, first white out-neighbour,
                                   post-order reversed. Call sites network.jl:99,244, utils.jl:66.
  * stream_order / subbasins / fillnodata_upstream / graph_from_nodes / subbasins_order /
    kinwave_set_subdomains         Wflow/src/subdomains.jl:1-255
  * filter_upstream_nodes          Wflow/src/utils.jl:61-71
  * get_flow_fraction_to_river     Wflow/src/utils.jl:493-510
  * EdgeConnectivity               Wflow/src/network.jl:27-33,136-153

Pinned against the reference's own golden vectors (Wflow/test/subdomains.jl:48-87) in
tests/test_oracle_golden.py.

All node ids are 1-based (Julia `Int`) in this module's inputs and outputs, so they can be
compared literally with the reference's numbers. Every graph on this path is a forest with
out-degree <= 1, so a graph is represented by `down` (downstream node id, 0 = none) plus a
CSR of in-neighbours sorted ascending (Graphs.jl keeps adjacency lists sorted).
"""
from __future__ import annotations

import numpy as np

LDD_PIT = 5
# Wflow/src/utils.jl:2-12 : PCRaster LDD value (1..9) -> CartesianIndex offset (d1, d2)
PCR_DIR = np.array(
    [(-1, -1), (0, -1), (1, -1), (-1, 0), (0, 0), (1, 0), (-1, 1), (0, 1), (1, 1)],
    dtype=np.int64,
)


class DiGraph1:
    """Forest digraph: down[v-1] = downstream id (0 = none); in-neighbours via CSR."""

    def __init__(self, down: np.ndarray):
        self.down = np.asarray(down, dtype=np.int64)
        n = self.n = len(self.down)
        src = np.nonzero(self.down > 0)[0]  # 0-based sources, ascending
        dst = self.down[src] - 1
        order = np.argsort(dst, kind="stable")  # ascending source ids within a destination
        counts = np.bincount(dst, minlength=n)
        self.in_ptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(counts, out=self.in_ptr[1:])
        self.in_idx = (src[order] + 1).astype(np.int64)

    def inneighbors(self, v: int) -> np.ndarray:
        return self.in_idx[self.in_ptr[v - 1]:self.in_ptr[v]]

    def outneighbors(self, v: int) -> list:
        d = int(self.down[v - 1])
        return [d] if d else []


def active_indices(mask2d: np.ndarray):
    """utils.jl:85-99. mask2d: bool (d1, d2) in Julia dimension order. Returns
    (indices (n,2) 1-based CartesianIndex in column-major order, reverse_indices (d1,d2))."""
    d1, d2 = mask2d.shape
    lin = np.nonzero(mask2d.ravel(order="F"))[0]
    i = lin % d1 + 1
    j = lin // d1 + 1
    indices = np.stack([i, j], axis=1).astype(np.int64)
    rev = np.zeros((d1, d2), dtype=np.int64)
    rev[i - 1, j - 1] = np.arange(1, len(lin) + 1)
    return indices, rev


def flowgraph(ldd: np.ndarray, indices: np.ndarray, d1: int):
    """routing/utils.jl:6-33. Returns (DiGraph1, possibly modified ldd). Out-of-domain
    targets become pits; cycles raise."""
    ldd = np.array(ldd, dtype=np.uint8, copy=True)
    n = len(ldd)
    lin = (indices[:, 1] - 1) * np.int64(d1) + (indices[:, 0] - 1)  # column-major, ascending
    # CartesianIndex ordering == column-major linear order, so searchsortedfirst == this:
    off = PCR_DIR[ldd.astype(np.int64) - 1]
    ti = indices[:, 0] + off[:, 0]
    tj = indices[:, 1] + off[:, 1]
    tlin = (tj - 1) * np.int64(d1) + (ti - 1)
    pos = np.searchsorted(lin, tlin)
    valid = (ti >= 1) & (ti <= d1) & (tj >= 1) & (pos < n)
    posc = np.minimum(pos, n - 1)
    valid &= lin[posc] == tlin
    notpit = ldd != LDD_PIT
    bad = notpit & ~valid
    ldd[bad] = LDD_PIT
    down = np.where(notpit & valid, posc + 1, 0).astype(np.int64)
    g = DiGraph1(down)
    topological_sort_by_dfs(g)  # raises on cycles (is_cyclic)
    return g, ldd


def topological_sort_by_dfs(g: DiGraph1) -> np.ndarray:
    """Graphs.jl topological_sort_by_dfs, specialised to out-degree <= 1 but keeping the
    literal colour/stack structure."""
    n = g.n
    down = g.down.tolist()
    color = [0] * (n + 1)
    verts = []
    for v in range(1, n + 1):
        if color[v] != 0:
            continue
        stack = [v]
        color[v] = 1
        while stack:
            u = stack[-1]
            w = 0
            d = down[u - 1]
            if d:
                if color[d] == 1:
                    raise ValueError("The input graph contains at least one loop.")
                if color[d] == 0:
                    w = d
            if w:
                color[w] = 1
                stack.append(w)
            else:
                color[u] = 2
                verts.append(u)
                stack.pop()
    return np.array(verts[::-1], dtype=np.int64)


def stream_order(g: DiGraph1, toposort: np.ndarray) -> np.ndarray:
    """subdomains.jl:32-47 (Strahler)."""
    n = len(toposort)
    strord = [1] * (n + 1)
    ptr, idx = g.in_ptr.tolist(), g.in_idx.tolist()
    for v in toposort.tolist():
        a, b = ptr[v - 1], ptr[v]
        if b > a:
            sto_up = [strord[u] for u in idx[a:b]]
            mx = max(sto_up)
            strord[v] = mx + 1 if sto_up.count(mx) > 1 else mx
    return np.array(strord[1:], dtype=np.int64)


def subbasins(g: DiGraph1, streamorder, toposort, min_sto: int) -> np.ndarray:
    """subdomains.jl:55-82."""
    n = len(toposort)
    subbas = np.zeros(n, dtype=np.int64)
    down = g.down
    i = 1
    for v in toposort.tolist():
        if streamorder[v - 1] < min_sto:
            continue
        d = down[v - 1]
        if d:
            if streamorder[v - 1] != streamorder[d - 1]:
                subbas[v - 1] = i
                i += 1
        else:
            subbas[v - 1] = i
            i += 1
    return subbas


def fillnodata_upstream(g: DiGraph1, toposort, data, nodata: int) -> np.ndarray:
    """subdomains.jl:8-24."""
    out = np.array(data, dtype=np.int64, copy=True).tolist()
    down = g.down.tolist()
    for v in toposort[::-1].tolist():
        d = down[v - 1]
        if d:
            if out[v - 1] == nodata and out[d - 1] != nodata:
                out[v - 1] = out[d - 1]
    return np.array(out, dtype=np.int64)


def graph_from_nodes(g: DiGraph1, subbas, subbas_fill) -> DiGraph1:
    """subdomains.jl:127-143."""
    n = int(subbas.max())
    down = np.zeros(n, dtype=np.int64)
    node_of = np.zeros(n + 1, dtype=np.int64)
    nz = np.nonzero(subbas > 0)[0]
    node_of[subbas[nz]] = nz + 1
    for i in range(1, n + 1):
        d = g.down[node_of[i] - 1]
        if d:
            down[i - 1] = subbas_fill[d - 1]
    return DiGraph1(down)


def distances_undirected(g: DiGraph1, s: int) -> np.ndarray:
    """Graphs.Experimental.Traversals.distances(Graph(g), s): BFS hop counts."""
    n = g.n
    dist = np.full(n, -1, dtype=np.int64)
    dist[s - 1] = 0
    frontier = [s]
    while frontier:
        nxt = []
        for u in frontier:
            nbrs = list(g.inneighbors(u)) + g.outneighbors(u)
            for w in nbrs:
                if dist[w - 1] < 0:
                    dist[w - 1] = dist[u - 1] + 1
                    nxt.append(int(w))
        frontier = nxt
    return dist


def subbasins_order(g: DiGraph1, outlet: int, max_dist: int):
    """subdomains.jl:93-120, including Julia's iterate-while-filter! semantics: when the
    current element is removed, the element that slides into its slot is skipped."""
    order = [None] * (max_dist + 1)
    order[0] = [outlet]
    for i in range(max_dist):
        v = []
        for n in order[i]:
            ups = g.inneighbors(n)
            if len(ups):
                v.extend(int(x) for x in ups)
        order[i + 1] = v
    for i in range(max_dist):
        lst = order[i]
        k = 0
        while k < len(lst):  # Julia: iterate(A, k) re-checks length(A) every step
            s = lst[k]
            k += 1
            if len(g.inneighbors(s)) == 0:
                order[max_dist].append(s)
                lst[:] = [e for e in lst if e != s]
        order[i] = lst
    return order[::-1]


def _induced(down_parent: np.ndarray, nodes: np.ndarray) -> DiGraph1:
    """induced_subgraph(g, nodes) with vmap = nodes (ascending 1-based ids)."""
    d = down_parent[nodes - 1]
    pos = np.searchsorted(nodes, d)
    posc = np.minimum(pos, len(nodes) - 1)
    inside = (d > 0) & (pos < len(nodes)) & (nodes[posc] == d)
    return DiGraph1(np.where(inside, posc + 1, 0))


def kinwave_set_subdomains(g: DiGraph1, toposort, index_pit, streamorder, min_sto: int,
                           nthreads: int):
    """subdomains.jl:169-255. Returns (subbas_order, indices_subbas, topo_subbas) as lists of
    int64 arrays (1-based)."""
    n = len(toposort)
    if nthreads <= 1 or n == 0:  # (n == 0 is not reachable in the reference: no river cells)
        return ([np.array([1], dtype=np.int64)], [np.arange(1, n + 1, dtype=np.int64)],
                [np.asarray(toposort, dtype=np.int64)])
    index_pit = np.asarray(index_pit, dtype=np.int64)
    n_pits = len(index_pit)
    basin = np.zeros(n, dtype=np.int64)
    basin[index_pit - 1] = np.arange(1, n_pits + 1)
    basin_fill = fillnodata_upstream(g, toposort, basin, 0)
    index_toposort = np.zeros(n, dtype=np.int64)
    index_toposort[toposort - 1] = np.arange(1, n + 1)

    # findall(x -> x == i, basin_fill) for every i at once (ascending ids inside a group)
    grp = np.argsort(basin_fill, kind="stable")
    starts = np.searchsorted(basin_fill[grp], np.arange(1, n_pits + 2))

    order_subbas, indices_subbas, topo_subbas, index = [], [], [], []
    total_subbas = 0
    for i in range(1, n_pits + 1):
        bas = (grp[starts[i - 1]:starts[i]] + 1).astype(np.int64)
        gb = _induced(g.down, bas)
        toposort_b = topological_sort_by_dfs(gb)
        so_b = streamorder[bas - 1]
        subbas = subbasins(gb, so_b, toposort_b, min_sto)
        subbas_fill = fillnodata_upstream(gb, toposort_b, subbas, 0)
        n_subbas = max(int((subbas > 0).sum()), 1)
        if n_subbas > 1:
            graph_subbas = graph_from_nodes(gb, subbas, subbas_fill)
            toposort_subbas = topological_sort_by_dfs(graph_subbas)
            dist = distances_undirected(graph_subbas, int(toposort_subbas[-1]))
            max_dist = max(int(dist.max()), 1)
            v_subbas = subbasins_order(graph_subbas, int(toposort_subbas[-1]), max_dist)
        else:
            v_subbas = [[1]]
        v_subbas = [[x + total_subbas for x in grp_] for grp_ in v_subbas]
        total_subbas += n_subbas
        order_subbas.extend(v_subbas)
        index.extend(range(1, len(v_subbas) + 1))
        if n_subbas > 1:
            sg_grp = np.argsort(subbas_fill, kind="stable")
            sg_starts = np.searchsorted(subbas_fill[sg_grp], np.arange(1, n_subbas + 2))
            for s in range(1, n_subbas + 1):
                subbas_s = (sg_grp[sg_starts[s - 1]:sg_starts[s]] + 1).astype(np.int64)
                sg = _induced(gb.down, subbas_s)
                toposort_sg = topological_sort_by_dfs(sg)
                nodes = bas[subbas_s[toposort_sg - 1] - 1]
                topo_subbas.append(nodes)
                indices_subbas.append(index_toposort[nodes - 1])
        else:
            nodes = bas[toposort_b - 1]
            topo_subbas.append(nodes)
            indices_subbas.append(index_toposort[nodes - 1])
    index = np.array(index, dtype=np.int64)
    subbas_order = []
    for m in range(1, int(index.max()) + 1):
        parts = [order_subbas[k] for k in np.nonzero(index == m)[0]]
        subbas_order.append(np.array([x for p in parts for x in p], dtype=np.int64))
    return subbas_order, indices_subbas, topo_subbas


def filter_upstream_nodes(g: DiGraph1, toposort, vec_logical=None):
    """utils.jl:61-71: upstream lists INDEXED BY TOPOSORT POSITION. Returns CSR
    (ptr (n+1), idx) with 1-based node ids, ascending within a list."""
    n = g.n
    cnt = g.in_ptr[1:] - g.in_ptr[:-1]
    if vec_logical is not None and np.any(vec_logical):
        keep = ~np.asarray(vec_logical, dtype=bool)[g.in_idx - 1]
    else:
        keep = np.ones(len(g.in_idx), dtype=bool)
    ptr = np.zeros(n + 1, dtype=np.int64)
    lists = []
    for k, v in enumerate(toposort.tolist()):
        a, b = g.in_ptr[v - 1], g.in_ptr[v]
        ups = g.in_idx[a:b][keep[a:b]]
        lists.append(ups)
        ptr[k + 1] = ptr[k] + len(ups)
    idx = np.concatenate(lists) if lists else np.zeros(0, dtype=np.int64)
    return ptr, idx.astype(np.int64)


def get_flow_fraction_to_river(g: DiGraph1, ldd, inds_river, slope) -> np.ndarray:
    """utils.jl:493-510."""
    fraction = np.zeros(len(slope))
    for i in np.asarray(inds_river).tolist():
        for j in g.inneighbors(i).tolist():
            if ldd[j - 1] != ldd[i - 1]:
                fraction[j - 1] = slope[j - 1] / (slope[i - 1] + slope[j - 1])
    return fraction


# DIRS / NEIGHBORS                  Wflow/src/network.jl:3, routing/subsurface/connectivity.jl:61-66
EDGE_DIRS = (("ind_y_down", (0, -1)), ("ind_x_down", (-1, 0)), ("ind_x_up", (1, 0)),
             ("ind_y_up", (0, 1)))


def edge_connectivity(indices: np.ndarray, d1: int, d2: int) -> dict:
    """EdgeConnectivity(network::NetworkLand)                    Wflow/src/network.jl:136-153:
    for every active cell the index (1-based) of the active neighbour in each of the four
    directions, n + 1 where the neighbour is outside the raster or inactive."""
    n = len(indices)
    rev = np.zeros((d1 + 2, d2 + 2), dtype=np.int64)          # reverse_indices, padded
    rev[indices[:, 0], indices[:, 1]] = np.arange(1, n + 1)
    out = {}
    for name, (di, dj) in EDGE_DIRS:
        r = rev[indices[:, 0] + di, indices[:, 1] + dj]
        out[name] = np.where(r != 0, r, n + 1).astype(np.int64)
    return out


def build_domain_network(ldd, indices, d1, min_sto, nthreads, streamorder=None, pits_mask=None):
    """NetworkLand / NetworkRiver construction (network.jl:87-133, 214-278; domain.jl:80-125).
    pits_mask: the reservoir outlet cells of this domain (domain.jl:96-109): they are dropped
    from the upstream lists AFTER order and sub-domains have been built on the full graph.
    Returns a dict of the reference's artefacts (1-based)."""
    g, ldd2 = flowgraph(ldd, indices, d1)
    order = topological_sort_by_dfs(g)
    so = stream_order(g, order) if streamorder is None else np.asarray(streamorder)
    pits = np.nonzero(ldd2 == LDD_PIT)[0] + 1
    sub_order, sub_indices, sub_topo = kinwave_set_subdomains(g, order, pits, so, min_sto, nthreads)
    up_ptr, up_idx = filter_upstream_nodes(g, order, pits_mask)
    return dict(graph=g, ldd=ldd2, order=order, streamorder=so, pits=pits,
                order_of_subdomains=sub_order, subdomain_indices=sub_indices,
                order_subdomain=sub_topo, up_ptr=up_ptr, up_idx=up_idx)
