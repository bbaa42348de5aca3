"""The N > 1 path on CPU: world_size-2 `gloo` processes shard ONE model domain by whole
drainage basins (wflow.jl_b200/partition.py), each rank advances its shard (here with the CPU
oracle standing in for the rank's GPU), and the gathered fields equal a single-process run of
the whole domain bit for bit -- the evidence that basin-aligned shards need no data-path
collective. Also: the shards are disjoint, cover the domain, cut no drainage
edge, and the product's host-side artefact builder accepts every shard."""
import os
import socket
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, d1, d2, steps, result_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from __graft_entry__ import load_pkg
    import parity
    pkg = load_pkg()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, dom, fields = pkg.synthetic.make_basin(d1, d2, seed=23)
    shards = pkg.partition.partition_basins(dom, world)
    sh = shards[rank]
    table = dict(pkg._lib.field_table())
    ldom = pkg.partition.shard_domain(dom, sh)
    lcfg = pkg.partition.shard_config(cfg, sh)
    lfields = pkg.partition.shard_fields(fields, table, sh)
    # the product's host-side artefact builder accepts the shard (no GPU needed)
    art = pkg.build_network_artifacts(lcfg, ldom)
    assert sorted(art["land"]["order"].tolist()) == list(range(1, lcfg["n"] + 1))
    # coverage / disjointness through the process group
    owned = torch.zeros(cfg["n"], dtype=torch.int64)
    owned[torch.from_numpy(sh.cells)] = 1
    dist.all_reduce(owned)
    assert bool((owned == 1).all())
    ora = parity.make_oracle(lcfg, ldom, lfields)
    dt = cfg["dt"]
    for step in range(steps):
        p, e, t = pkg.synthetic.make_forcing(23, step, ldom["gid"], dt)
        ora.f["precipitation"][:], ora.f["potential_evaporation"][:], ora.f["temperature"][:] = p, e, t
        ora.update_model(dt)
    # gather a land and a river field on every rank (dense global vectors, summed)
    out = {}
    for name, kind in (("recharge", 0), ("olf_q_average", 0), ("ssf_q", 0), ("total_storage", 0),
                       ("riv_q_average", 3), ("riv_h", 3)):
        size = cfg["nriv"] if kind == 3 else cfg["n"]
        g = torch.zeros(size, dtype=torch.float64)
        g[torch.from_numpy(sh.river_cells if kind == 3 else sh.cells)] = torch.from_numpy(
            np.ascontiguousarray(ora.f[name]))
        dist.all_reduce(g)
        out[name] = g.numpy()
    total = torch.tensor([float(lcfg["n"])], dtype=torch.float64)
    dist.all_reduce(total)
    assert int(total.item()) == cfg["n"]
    if rank == 0:
        np.savez(os.path.join(result_dir, "gathered.npz"), **out)
    dist.barrier()
    dist.destroy_process_group()


def test_basin_shards_need_no_exchange(pkg, tmp_path):
    import torch.multiprocessing as mp
    import parity
    d1, d2, steps, world = 36, 48, 3, 2
    mp.spawn(_worker, args=(world, _free_port(), d1, d2, steps, str(tmp_path)), nprocs=world,
             join=True)
    got = np.load(tmp_path / "gathered.npz")
    cfg, dom, fields = pkg.synthetic.make_basin(d1, d2, seed=23)
    ora = parity.make_oracle(cfg, dom, fields)
    dt = cfg["dt"]
    for step in range(steps):
        p, e, t = pkg.synthetic.make_forcing(23, step, dom["gid"], dt)
        ora.f["precipitation"][:], ora.f["potential_evaporation"][:], ora.f["temperature"][:] = p, e, t
        ora.update_model(dt)
    for name in got.files:
        assert np.array_equal(got[name], ora.f[name]), name


def test_partition_is_balanced_and_closed(pkg):
    cfg, dom, _ = pkg.synthetic.make_basin(80, 120, seed=4)
    for world in (2, 4, 8):
        shards = pkg.partition.partition_basins(dom, world)
        allc = np.concatenate([s.cells for s in shards])
        assert sorted(allc.tolist()) == list(range(cfg["n"]))
        allr = np.concatenate([s.river_cells for s in shards])
        assert sorted(allr.tolist()) == list(range(cfg["nriv"]))
        w = np.array([s.weight for s in shards])
        # greedy LPT: balanced unless one indivisible basin is heavier than a fair share
        basin = pkg.partition.basin_of_cells(dom["down"])
        wcell = np.ones(cfg["n"])
        wcell[dom["river_land_indices"] - 1] += 3.0
        heaviest = np.bincount(basin, weights=wcell).max()
        assert w.max() <= max(1.35 * w.mean(), heaviest)
        down = dom["down"]
        owner = np.empty(cfg["n"], dtype=np.int64)
        for s in shards:
            owner[s.cells] = s.rank
        has = down > 0
        assert np.all(owner[has] == owner[down[has] - 1])   # no drainage edge is cut
        # downstream_ids from the gridded LDD agrees with the generator's own table
        d2 = dict(dom)
        d2.pop("down")
        assert np.array_equal(pkg.partition.downstream_ids(d2), down)
