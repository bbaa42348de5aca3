"""ONE drainage basin cut across handles (GPUs): discharge crosses the cut on cut edges that the
producing GPU writes straight into the consuming GPU's inlet slots (wflowb200_exchange_*).
The cut basin must reproduce the single-handle run BIT FOR BIT: same sums in the same order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return max(1, torch.cuda.device_count())


@pytest.mark.parametrize("network,parts", [("dendritic", 2), ("dendritic", 4), ("scheidegger", 3)])
def test_cut_basin_equals_one_handle_bit_for_bit(pkg, network, parts):
    P = pkg.partition
    kw = dict(network="dendritic") if network == "dendritic" else {}
    cfg, dom, fields = pkg.synthetic.make_basin(60, 90, seed=31, **kw)
    dt = cfg["dt"]
    down = P.downstream_ids(dom)
    owner = P.split_by_subtrees(down, parts)
    plans = P.cut_basin(dom, owner, parts)
    n_cut = sum(len(pl["links"][0]) for pl in plans), sum(len(pl["links"][1]) for pl in plans)
    assert n_cut[0] >= parts - 1, "the partition cut nothing"
    ndev = _device_count()
    table = dict(pkg._lib.field_table())
    one = pkg.SbmModel(cfg, dom, fields)
    shards = []
    for p, pl in enumerate(plans):
        lcfg = P.shard_config(cfg, pl["shard"])
        lfields = P.shard_fields(fields, table, pl["shard"])
        lfields.pop("nlayers_kv", None)
        m = pkg.SbmModel(lcfg, pl["domain"], lfields, device=p % ndev)
        if ndev < parts:   # several spinning wavefronts on one GPU: all of them must be resident
            m.set_option("wave_grid_percent", 90 // parts)
        shards.append(m)
    with pytest.raises(pkg.WflowB200Error):   # cut edges without the exchange set up
        shards[-1].update_model(dt)
    group = pkg.ShardGroup(shards)
    P.connect_in_process(shards, plans, dt)
    for step in range(3):
        pr, e, t = pkg.synthetic.make_forcing(31, step, dom["gid"], dt)
        one.set_forcing(pr, e, t)
        one.update_model(dt)
        for m, pl in zip(shards, plans):
            c = pl["shard"].cells
            m.set_forcing(np.ascontiguousarray(pr[c]), np.ascontiguousarray(e[c]),
                          np.ascontiguousarray(t[c]))
        group.step_all(dt)
    worst = []
    for name in one.field_names():
        kind = table.get(name, 0)
        if kind in (4, 6):   # reservoirs; 2-D local-inertial overland flow (absent here)
            continue
        want = one.get(name)
        got = np.empty_like(want)
        for m, pl in zip(shards, plans):
            sh = pl["shard"]
            got[sh.river_cells if kind in (3, 5) else sh.cells] = m.get(name)
        if not np.array_equal(got, want, equal_nan=True):
            worst.append(name)
    print(f"{network}: {parts} parts on {min(ndev, parts)} GPU(s), cut edges land/river {n_cut}")
    assert not worst, worst
    group.close()
    for m in [one] + shards:
        m.close()
