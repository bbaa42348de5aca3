"""ONE basin cut across the GPUs of a node, one process per GPU (torchrun): cut edges over NVLink
peer memory (CUDA IPC), the per-step barrier over NCCL. Checks the gathered fields against a
single-handle run of the whole basin on rank 0 bit for bit and prints per-step times.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 scripts/cut_basin_multi_gpu.py --d1 600 --d2 600 --steps 12
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d1", type=int, default=300)
    ap.add_argument("--d2", type=int, default=300)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--network", default="dendritic")
    ap.add_argument("--seed", type=int, default=31)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    P = pkg.partition
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kw = dict(network="dendritic") if args.network == "dendritic" else {}
    cfg, dom, fields = pkg.synthetic.make_basin(args.d1, args.d2, seed=args.seed, **kw)
    dt = cfg["dt"]
    down = P.downstream_ids(dom)
    rli = np.asarray(dom["river_land_indices"]) - 1
    w = np.ones(len(down))
    w[rli] += 3.0
    owner = P.split_by_subtrees(down, world, w)
    plans = P.cut_basin(dom, owner, world)
    pl = plans[rank]
    table = dict(pkg._lib.field_table())
    lcfg = P.shard_config(cfg, pl["shard"])
    lfields = P.shard_fields(fields, table, pl["shard"])
    lfields.pop("nlayers_kv", None)
    m = pkg.SbmModel(lcfg, pl["domain"], lfields, device=local)
    uid = [pkg.SbmModel.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    m.comm_init_nccl(rank, world, uid[0])
    P.connect_distributed(m, pl, dt, dist)
    cells = pl["shard"].cells

    total = args.warmup + args.steps
    forcing = [pkg.synthetic.make_forcing(args.seed, s, dom["gid"], dt) for s in range(total)]
    forcing_cut = [tuple(np.ascontiguousarray(a[cells]) for a in f3) for f3 in forcing]

    def step(model, s, sel):   # (the forcing is generated outside the timed loops)
        model.set_forcing(*(forcing[s] if sel is None else forcing_cut[s]))
        model.update_model(dt)

    for s in range(args.warmup):
        step(m, s, cells)
    m.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for s in range(args.warmup, total):
        step(m, s, cells)
    m.synchronize()
    dist.barrier()
    t_cut = (time.perf_counter() - t0) / args.steps
    # gather a few fields on rank 0 and compare with the whole basin on one GPU
    names = ["riv_q_average", "riv_h", "olf_q_average", "ssf_q_average", "ssf_water_table_depth",
             "total_storage", "recharge", "olf_h", "riv_q", "ssf_q"]
    mine = {k: m.get(k) for k in names}
    gathered = [None] * world
    dist.gather_object(mine, gathered if rank == 0 else None, dst=0)
    report = None
    if rank == 0:
        one = pkg.SbmModel(cfg, dom, fields, device=local)
        for s in range(args.warmup):
            step(one, s, None)
        one.synchronize()
        t0 = time.perf_counter()
        for s in range(args.warmup, total):
            step(one, s, None)
        one.synchronize()
        t_one = (time.perf_counter() - t0) / args.steps
        bad = []
        for k in names:
            want = one.get(k)
            got = np.empty_like(want)
            for q, g in enumerate(gathered):
                sh = plans[q]["shard"]
                got[sh.river_cells if table.get(k, 0) == 3 else sh.cells] = g[k]
            if not np.array_equal(got, want, equal_nan=True):
                bad.append(k)
        report = dict(world=world, cells=int(cfg["n"]), river_cells=int(cfg["nriv"]),
                      cells_per_part=[int(len(p_["shard"].cells)) for p_ in plans],
                      cut_edges_land=sum(len(p_["links"][0]) for p_ in plans),
                      cut_edges_river=sum(len(p_["links"][1]) for p_ in plans),
                      levels_one_handle=int(one.stats()["wave_levels_land"]),
                      ms_per_step_cut=1e3 * t_cut, ms_per_step_one_gpu=1e3 * t_one,
                      fields_compared=names, fields_differing=bad, bit_identical=not bad)
        one.close()
        print(json.dumps(report))
    m.close()
    dist.destroy_process_group()
    if rank == 0 and report and not report["bit_identical"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
