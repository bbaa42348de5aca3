"""CPU-side tests of the product: the C-ABI library loads and exports every symbol the header
declares, the host-built indexing artefacts are bit-exact against the oracle's independent
restatement, and the product fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity  # noqa: E402
from oracle.oracle import _csr  # noqa: E402


def test_library_exports_every_header_symbol(pkg):
    L = pkg._lib.lib()
    syms = pkg.header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"libwflow_b200.so does not export {s}"


def test_field_table_matches_oracle_schema(pkg):
    """The product and the oracle keep separate field lists; they must agree on names/kinds for
    every product field (the oracle has a few extra fields for profiles not yet on the GPU)."""
    from oracle import oracle as orc
    prod = dict(pkg._lib.field_table())
    ora = dict(orc.field_table())
    for name, kind in prod.items():
        assert ora.get(name) == kind, name
    L = pkg._lib.lib()
    assert L.wflowb200_field_id(b"olf_q") == list(prod).index("olf_q")
    assert L.wflowb200_field_id(b"nope") == -1


@pytest.mark.parametrize("d1,d2,ml,mr,seed,nthreads", [
    (40, 60, 5, 6, 42, 8), (120, 200, 3, 3, 1, 2), (64, 512, 2, 4, 9, 8), (200, 100, 4, 2, 3, 4),
    (50, 50, 5, 6, 7, 1), (1, 30, 1, 1, 2, 8), (30, 1, 1, 1, 2, 8)])
def test_indexing_artifacts_bit_exact(pkg, d1, d2, ml, mr, seed, nthreads):
    cfg, dom, _ = pkg.synthetic.make_basin(d1, d2, seed=seed, nthreads=nthreads)
    cfg["land_streamorder_min"], cfg["river_streamorder_min"] = ml, mr
    art = pkg.build_network_artifacts(cfg, dom)
    land, river = parity.oracle_networks(cfg, dom)
    for name, o in (("land", land), ("river", river)):
        a = art[name]
        assert np.array_equal(a["order"], o["order"]), name
        assert np.array_equal(a["streamorder"], o["streamorder"]), name
        assert np.array_equal(a["upstream_ptr"], o["up_ptr"]), name
        assert np.array_equal(a["upstream_idx"], o["up_idx"]), name
        assert np.array_equal(a["ldd"], o["ldd"]), name
        lp, li = _csr(o["order_of_subdomains"])
        assert np.array_equal(a["subdomain_level_ptr"], lp), name
        assert np.array_equal(a["subdomain_level_idx"], li), name
        sp, so = _csr(o["order_subdomain"])
        _, si = _csr(o["subdomain_indices"])
        assert np.array_equal(a["subdomain_ptr"], sp), name
        assert np.array_equal(a["subdomain_order"], so), name
        assert np.array_equal(a["subdomain_indices"], si), name


def test_masked_raster_artifacts_and_pit_fixup(pkg):
    mask = np.ones((37, 53), dtype=bool)
    mask[:5, :7] = False
    mask[20:, 40:] = False
    cfg, dom, _ = pkg.synthetic.make_basin(37, 53, seed=4, mask=mask)
    # hand the library the raw LDD of the unmasked raster (it still points out of the mask):
    # flowgraph must turn exactly those cells into pits (routing/utils.jl:20-24)
    _, ldd_full, _, _ = pkg.synthetic.scheidegger_ldd(37, 53, 4)
    raw = ldd_full[np.nonzero(mask.ravel(order="F"))[0]]
    assert (raw != dom["ldd"]).any()
    dom2 = dict(dom, ldd=raw)
    art = pkg.build_network_artifacts(cfg, dom2)
    assert np.array_equal(art["land"]["ldd"], dom["ldd"])
    land, _ = parity.oracle_networks(cfg, dom2)
    assert np.array_equal(art["land"]["order"], land["order"])


@pytest.mark.parametrize("piece_land,piece_river", [(None, None), ("0", "0"), ("6", "4")])
def test_wavefront_levels_and_chunks_are_a_valid_schedule(pkg, monkeypatch, piece_land, piece_river):
    """Every drainage edge spans exactly one wavefront level (the invariant the skewed
    wavefront relies on); a chunk holds at most 32 nodes in at most 32 levels, made of connected
    pieces with one root each; slots are ordered (chunk, level, node id); every edge that leaves
    a chunk leads to a LATER chunk (queue order = topological order of the chunk DAG)."""
    if piece_land is not None:
        monkeypatch.setenv("WFB_PIECE_LAND", piece_land)
        monkeypatch.setenv("WFB_PIECE_RIVER", piece_river)
    cfg, dom, _ = pkg.synthetic.make_basin(90, 140, seed=5)
    art = pkg.build_network_artifacts(cfg, dom)
    rl = dom["river_land_indices"]
    # river forest: downstream river node (1-based river id) of every river node
    riv_of_land = np.zeros(cfg["n"] + 1, dtype=np.int64)
    riv_of_land[rl] = np.arange(1, len(rl) + 1)
    down_land = dom["down"]
    down_riv = riv_of_land[down_land[rl - 1]]
    for name, down in (("land", down_land), ("river", down_riv)):
        a = art[name]
        perm, level, cp, roots = (a["wave_perm"], a["wave_node_level"], a["wave_chunk_ptr"],
                                  a["wave_chunk_outlet"])
        n = len(perm)
        assert sorted(perm.tolist()) == list(range(1, n + 1))
        has = down > 0
        assert np.all(level[down[has] - 1] == level[has] + 1)
        assert np.all(level[~has] == level.max())
        assert np.array_equal(np.bincount(level), np.diff(a["wave_level_ptr"]))
        chunk = np.zeros(n, dtype=np.int64)
        for c in range(len(cp) - 1):
            seg = perm[cp[c]:cp[c + 1]]
            chunk[seg - 1] = c
            assert 1 <= len(seg) <= 32
            assert level[seg - 1].max() - level[seg - 1].min() < 32
            key = level[seg - 1] * (n + 1) + seg
            assert np.all(np.diff(key) > 0)                  # (level, node id) ascending
        leaving = (~has) | (chunk[np.maximum(down, 1) - 1] != chunk)
        assert sorted((np.nonzero(leaving)[0] + 1).tolist()) == sorted(roots.tolist())
        out = has & leaving
        assert np.all(chunk[down[out] - 1] > chunk[out])     # producers come first in the queue
        assert len(cp) - 1 >= 2


def test_cycle_is_rejected(pkg):
    # two cells pointing at each other: 6 (east, +1 in d1) and 4 (west)
    cfg = dict(n_layers=4, nthreads=1)
    dom = dict(d1=2, d2=1, indices=np.array([[1, 1], [2, 1]]), ldd=np.array([6, 4], np.uint8),
               river_land_indices=np.zeros(0, np.int64))
    with pytest.raises(RuntimeError, match="cycle"):
        pkg.build_network_artifacts(cfg, dom)


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg, dom, fields = pkg.synthetic.make_basin(8, 8, seed=1)
    with pytest.raises(pkg.WflowB200Error, match="no CUDA device|CUDA"):
        pkg.SbmModel(cfg, dom, fields)


def test_oracle_water_balance_and_determinism(pkg):
    """The oracle itself: river-reach water balance closes (reference invariant,
    test/run_sbm.jl:1238-1264) and thread count does not change results."""
    cfg, dom, fields = pkg.synthetic.make_basin(48, 64, seed=13)
    dt = cfg["dt"]
    ora = parity.make_oracle(cfg, dom, fields)
    for step in range(3):
        p, e, t = pkg.synthetic.make_forcing(13, step, dom["gid"], dt)
        ora.f["precipitation"][:], ora.f["potential_evaporation"][:], ora.f["temperature"][:] = p, e, t
        r0 = ora.f["riv_storage"].copy()
        ora.update_model(dt)
    f = ora.f
    err = (f["riv_storage"] - r0) - (f["riv_qin_average"] + f["riv_inwater"] - f["riv_q_average"]) * dt
    assert np.max(np.abs(err) / np.maximum(np.abs(f["riv_q_average"] * dt), 1.0)) < 1e-6
    assert not np.isnan(f["total_storage"]).any()
    # single sub-domain walk (nthreads = 1 artefacts) gives the same numbers
    cfg1 = dict(cfg, nthreads=1)
    ora1 = parity.make_oracle(cfg1, dom, fields)
    for step in range(3):
        p, e, t = pkg.synthetic.make_forcing(13, step, dom["gid"], dt)
        ora1.f["precipitation"][:], ora1.f["potential_evaporation"][:], ora1.f["temperature"][:] = p, e, t
        ora1.update_model(dt)
    for k in ("riv_q", "olf_q", "ssf_q", "water_table_depth", "total_storage"):
        assert np.array_equal(ora1.f[k], f[k]), k


def test_edge_connectivity_bit_exact(pkg):
    """EdgeConnectivity of the land network (network.jl:136-153), host C++ against the oracle's
    restatement, on a masked raster."""
    from oracle import network as onw
    mask = np.ones((37, 53), dtype=bool)
    mask[:5, :7] = False
    mask[20:, 40:] = False
    mask[10:14, 22:30] = False
    cfg, dom, _ = pkg.synthetic.make_basin(37, 53, seed=4, mask=mask)
    art = pkg.build_network_artifacts(dict(cfg, land_routing=1, river_routing=1), dom)["land"]
    e = onw.edge_connectivity(dom["indices"], dom["d1"], dom["d2"])
    n = cfg["n"]
    for k in ("x_up", "x_down", "y_up", "y_down"):
        assert np.array_equal(art["edge_" + k], e["ind_" + k]), k
        assert int((e["ind_" + k] == n + 1).sum()) > 0
    # x_up and x_down are inverse to each other where both cells exist
    xu = e["ind_x_up"]
    has = xu <= n
    assert np.array_equal(e["ind_x_down"][xu[has] - 1], np.nonzero(has)[0] + 1)


def test_oracle_local_inertial_land_conserves_water(pkg):
    """2-D local-inertial overland flow coupled to the river (the oracle itself): what the cells
    hold changes by the runoff they receive, minus what leaves through the pits' ghost edges, plus
    what the negative-storage clip adds (li_land_error) -- every other flux is internal. The
    reference's data-free invariant (mass_balance.jl:61-62 pairs the two routing models)."""
    cfg, dom, fields = pkg.synthetic.make_basin(40, 56, seed=11, river_routing=1, land_routing=1)
    dt = cfg["dt"]
    ora = parity.make_oracle(cfg, dom, fields)
    down_r = np.asarray(parity.oracle_networks(cfg, dom)[1]["graph"].down)
    pits = down_r == 0
    for step in range(3):
        p, e, t = pkg.synthetic.make_forcing(11, step, dom["gid"], dt)
        ora.f["precipitation"][:], ora.f["potential_evaporation"][:], ora.f["temperature"][:] = p, e, t
        s0, err0 = ora.f["olf_storage"].sum(), ora.f["li_land_error"].sum()
        ora.update_model(dt)
        f = ora.f
        gained = f["li_land_runoff"].sum() * dt - f["riv_q_cumulative"][pits].sum() \
            + (f["li_land_error"].sum() - err0)
        assert abs((f["olf_storage"].sum() - s0) - gained) <= 1e-9 * max(abs(gained), f["olf_storage"].sum())
    assert ora.newton_stats()["substeps_river"] > 20
    assert (f["olf_h"] > 0).sum() > 50 and np.abs(f["li_land_qx_average"]).max() > 0
    # river cells: land h is the depth above bankfull, river storage never exceeds the cell's
    rl = dom["river_land_indices"] - 1
    over = f["olf_h"][rl] > 0
    assert np.all(f["riv_h"][over] >= f["li_bankfull_depth"][over])
    assert np.all(f["riv_h"][~over] <= f["li_bankfull_depth"][~over] * (1 + 1e-12))


def test_sharding_refuses_what_it_cannot_keep_exact(pkg):
    """Basin-aligned shards are independent only for the kinematic wave without reservoirs: the
    staggered (local-inertial) schemes take ONE sub-step length for the whole domain and the 2-D
    overland flow crosses basin divides -- shard_config refuses instead of changing the result."""
    cfg, dom, _ = pkg.synthetic.make_basin(24, 32, seed=3)
    sh = pkg.partition.partition_basins(dom, 2)[0]
    assert pkg.partition.shard_config(cfg, sh)["sharded"] is True
    for bad in (dict(river_routing=1), dict(river_routing=1, land_routing=1), dict(nres=2)):
        with pytest.raises(ValueError):
            pkg.partition.shard_config(dict(cfg, **bad), sh)


def test_create_rejects_bad_arguments_before_touching_the_device(pkg):
    """Argument validation of wflowb200_create happens before any CUDA call: status
    WFLOWB200_ERR_ARG (1) and a message, never a crash (the shim turns it into error(...))."""
    import ctypes as C
    L = pkg._lib.lib()
    cfg, dom, _ = pkg.synthetic.make_basin(6, 8, seed=1)
    idx = np.ascontiguousarray(dom["indices"], dtype=np.int64)
    ldd = np.ascontiguousarray(dom["ldd"], dtype=np.uint8)
    rli = np.ascontiguousarray(dom["river_land_indices"], dtype=np.int64)
    d = pkg._lib.Domain(dom["d1"], dom["d2"], idx.ctypes.data, ldd.ctypes.data, rli.ctypes.data)
    for bad in (dict(n=0), dict(n_layers=0), dict(n_layers=9), dict(kv_profile=7),
                dict(land_routing=1), dict(land_routing=2, river_routing=1)):
        c = pkg._lib.Config()
        c.n, c.nriv, c.n_layers = cfg["n"], cfg["nriv"], cfg["n_layers"]
        for k, v in bad.items():
            setattr(c, k, v)
        h = C.c_void_p()
        rc = L.wflowb200_create(C.byref(c), C.byref(d), C.byref(h))
        assert rc == 1 and not h.value, bad
        assert L.wflowb200_last_error(None).decode()
    assert L.wflowb200_create(None, None, None) == 1


def test_cut_basin_plan_is_consistent(pkg):
    """partition.cut_basin: ONE basin cut at confluences into parts that own whole upstream
    subtrees. Every cut edge is exactly one export of the upstream part and one import of the
    downstream part, at the position its source has among the destination's upstream sources in
    ascending GLOBAL id (the reference's summation order, utils.jl:472-477); the parts form a DAG
    (a spinning consumer can never be what its producer waits for);
    the host-side artefact builder accepts every part."""
    P = pkg.partition
    for network, parts in (("dendritic", 4), ("scheidegger", 3)):
        kw = dict(network="dendritic") if network == "dendritic" else {}
        cfg, dom, fields = pkg.synthetic.make_basin(40, 60, seed=5, **kw)
        down = P.downstream_ids(dom)
        n = len(down)
        owner = P.split_by_subtrees(down, parts)
        plans = P.cut_basin(dom, owner, parts)
        assert sorted(np.concatenate([pl["shard"].cells for pl in plans]).tolist()) == list(range(n))
        cut = [(u, down[u] - 1) for u in range(n) if down[u] > 0 and owner[u] != owner[down[u] - 1]]
        assert len(cut) >= parts - 1
        seen = 0
        part_edges = set()
        for p, pl in enumerate(plans):
            d, cells = pl["domain"], pl["shard"].cells
            assert np.all(d["down"][d["land_export_src"] - 1] == 0)
            assert np.all(d["ldd"][d["land_export_src"] - 1] == 5)
            for e, (q, k) in enumerate(pl["links"][0]):
                u = cells[d["land_export_src"][e] - 1]
                dq = plans[q]["domain"]
                v = plans[q]["shard"].cells[dq["land_import_dst"][k] - 1]
                assert down[u] - 1 == v and owner[v] == q and q != p
                ups = np.sort(np.nonzero(down == v + 1)[0])
                assert dq["land_import_pos"][k] == int(np.searchsorted(ups, u))
                part_edges.add((p, q))
                seen += 1
            lcfg = P.shard_config(cfg, pl["shard"])
            art = pkg.build_network_artifacts(lcfg, d)
            assert sorted(art["land"]["order"].tolist()) == list(range(1, lcfg["n"] + 1))
        assert seen == len(cut)
        # a part never waits (directly or not) for a part that waits for it
        left, edges = set(range(parts)), set(part_edges)
        while left:
            sinks = [p for p in left if not any(a == p and b in left for a, b in edges)]
            assert sinks, "cycle between the parts"
            left -= set(sinks)
        # river cut edges are the land cut edges between river cells
        rli = np.asarray(dom["river_land_indices"]) - 1
        is_riv = np.zeros(n, dtype=bool)
        is_riv[rli] = True
        n_riv_cut = sum(1 for u, v in cut if is_riv[u] and is_riv[v])
        assert sum(len(pl["links"][1]) for pl in plans) == n_riv_cut


@pytest.mark.parametrize("flags", [[], ["--local-inertial-land"], ["--hourly"]])
def test_bench_reference_arm_prints_the_contract_line(flags):
    """`bench.py --impl reference` (the CPU port on the host cores) prints ONE JSON line with the
    keys of the bench contract and the same `config` keys as the GPU arm -- also for the workload
    variants (2-D local-inertial overland flow, hourly step)."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference",
                          "--size", "40", "--steps", "2", "--warmup", "1"] + flags,
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
              "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert set(d["config"]) == {"workload", "cells_per_gpu", "river_cells_per_gpu", "parallelism"}
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
