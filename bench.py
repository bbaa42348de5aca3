#!/usr/bin/env python
"""bench.py -- cell-timesteps/s of the wflow_sbm hot path (SBM vertical + kinematic-wave
routing) on B200, with the HBM roofline of the vertical kernel and the CPU baseline beside it.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU implementation (oracle port)

One "step" = one model time step (update_model!, sbm_model.jl:60-92) of a synthetic D8 basin:
forcing -> fused SBM vertical kernel -> subsurface / overland / river kinematic wave (24 + 96
+ 1 internal sub-steps at the reference's default fixed internal time steps) -> storages.
N = 1 runs BASELINE.json configs[1] (synthetic 1000 x 1000 basin); N > 1 runs one such
sub-catchment tile per GPU (disjoint catchments, no data-path collective: weak scaling).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg  # noqa: E402

METRIC = "cell-timesteps/s (SBM vertical + kinwave)"
ADAPTIVE = False
V1_DRAM_BYTES_PER_CELL = 1477  # measured (ncu), see roofline.traffic_source
UNIT = "cell-timesteps/s"


# --------------------------------------------------------------------------------------------
# algorithmic bytes: every distinct input array read once + every
# reference-visible output array written once, Float64 (int32 counters 4 B)
# --------------------------------------------------------------------------------------------
def v1_bytes_per_cell(N: int, cfg: dict) -> int:
    reads = ["precipitation", "potential_evaporation", "temperature", "crop_coefficient",
             "river_fraction", "water_fraction", "olf_h", "waterdepth_river", "theta_s", "theta_r",
             "theta_fc", "soil_thickness", "soil_water_capacity", "saturated_water_depth",
             "compacted_soil_area_fraction", "infiltration_capacity_soil",
             "infiltration_capacity_compacted_soil", "kv_0",
             "hydraulic_conductivity_scale_parameter", "rooting_depth", "h1", "h2", "h4",
             "alpha_h1", "air_entry_pressure", "h3_high", "h3_low",
             "wet_root_distribution_parameter", "cap_hmax", "cap_n", "maximum_leakage"]
    writes = ["canopy_potevap", "throughfall", "interception_rate", "stemflow",
              "runoff_water_flux_surface", "waterdepth_land", "runoff_river", "runoff_land",
              "actual_open_water_evaporation_river", "actual_open_water_evaporation_land",
              "net_runoff_river", "soil_fraction", "potential_transpiration",
              "potential_soilevaporation", "soil_water_flux_surface", "water_table_depth",
              "total_soil_water_storage", "f_infiltration_reduction", "infiltration",
              "infiltration_excess", "transfer", "soil_evaporation_saturated_zone",
              "soil_evaporation", "h3", "actual_evaporation_unsaturated_store",
              "actual_evaporation_saturated_zone", "transpiration", "actual_infiltration",
              "saturation_excess_water", "actual_infiltration_soil",
              "actual_infiltration_compacted_soil", "excess_water_soil",
              "excess_water_compacted_soil", "unsaturated_store_depth",
              "unsaturated_store_capacity", "actual_capillary_flux", "actual_leakage", "recharge",
              "actual_evapotranspiration", "drainable_water_depth"]
    if cfg["has_lai"]:
        reads += ["leaf_area_index", "storage_specific_leaf", "storage_wood",
                  "light_extinction_coefficient"]
        writes += ["maximum_canopy_storage", "canopy_gap_fraction"]
        if cfg["gash"]:
            writes += ["evaporation_to_precipitation_ratio"]
    else:
        reads += ["maximum_canopy_storage", "canopy_gap_fraction"]
        if cfg["gash"]:
            reads += ["evaporation_to_precipitation_ratio"]
    if not cfg["gash"]:
        reads += ["canopy_storage"]
        writes += ["canopy_storage"]
    if cfg["snow"]:
        reads += ["temperature_interval_snowfall", "temperature_threshold_snowfall",
                  "snow_storage", "snow_water", "temperature_threshold_melt", "degree_day_factor",
                  "water_holding_capacity", "soil_surface_temperature", "w_soil"]
        writes += ["effective_precip", "snow_precip", "liquid_precip", "snow_water",
                   "snow_water_equivalent", "snow_melt", "snow_runoff", "snow_storage",
                   "soil_surface_temperature"]
    layered_r = 6 * N + 1  # uld, alt, cld(N+1), bc, kvfac, rootfraction
    layered_w = 2 * N      # ult, uld
    return 8 * (len(reads) + len(writes) + layered_r + layered_w) + 4 + 4  # + 2 int32 arrays


def v2_bytes_per_cell(N: int) -> int:
    return 8 * ((13 + 4 * N) + (12 + 2 * N)) + 8


# --------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows for k in range(4)
                          if len(r) >= 7 and r[3 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def build_tile(pkg, size: int, rank: int, seed: int):
    """One sub-catchment tile per rank: a `size` x `size` Scheidegger forest whose cell ids
    are offset so that every tile of the global raster is a different random forest."""
    return pkg.synthetic.make_basin(size, size, seed=seed, id_offset=rank * size * size,
                                    adaptive=ADAPTIVE)


# --------------------------------------------------------------------------------------------
def run_cpu(pkg, cfg, dom, fields, steps: int, warmup: int, seed: int, first_step: int = 0):
    """The CPU implementation of the path: the C oracle (port of the Julia algorithm, OpenMP,
    threaded over the reference's own sub-domain partition). Julia is not available here, so
    kind = "port". Returns (cell-timesteps/s, cores, seconds per step)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    ora = parity.make_oracle(cfg, dom, fields)
    dt = cfg["dt"]
    gid = dom["gid"]
    # all the host threads the process may use (torchrun exports OMP_NUM_THREADS=1)
    cores = ora._L.wfo_set_num_threads(len(os.sched_getaffinity(0)))

    def one(step):
        p, e, t = pkg.synthetic.make_forcing(seed, step, gid, dt)
        ora.f["precipitation"][:], ora.f["potential_evaporation"][:], ora.f["temperature"][:] = p, e, t
        t0 = time.perf_counter()
        ora.update_model(dt)
        return time.perf_counter() - t0

    for s in range(warmup):
        one(first_step + s)
    el = sum(one(first_step + warmup + s) for s in range(steps))
    return cfg["n"] * steps / el, cores, el / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=1000, help="raster side per GPU")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--cpu-steps", type=int, default=2, help="oracle steps of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--partition", type=int, default=0, metavar="G",
                    help="shard ONE G x G basin raster over the ranks by whole drainage basins "
                         "(strong scaling) instead of one --size tile per rank")
    ap.add_argument("--adaptive", action="store_true",
                    help="adaptive internal routing time steps instead of the fixed defaults")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=INT",
                    help="wflowb200_set_option (kernel organisation), e.g. vertical_graph=0 for ncu")
    ap.add_argument("--cfg", action="append", default=[], metavar="NAME=INT",
                    help="WflowB200Config tuning field, e.g. vertical_slices=1")
    args = ap.parse_args()
    rank, world, local = dist_env()
    global ADAPTIVE
    ADAPTIVE = args.adaptive
    pkg = load_pkg()
    workload = f"synthetic {args.size}x{args.size} D8 basin per GPU, wflow_sbm vertical + " \
               "kinematic-wave river/overland/subsurface, daily step, " + \
               ("adaptive internal steps" if args.adaptive else
                "fixed internal steps 3600/900/86400 s") + ", N=4 soil layers, snow on"

    # ------------------------------------------------------------------ reference arm ----
    if args.impl == "reference":
        if rank != 0:
            return
        cfg, dom, fields = build_tile(pkg, args.size, 0, args.seed)
        k = max(1, min(args.steps, 3))
        w = min(args.warmup, 1)
        value, cores, sps = run_cpu(pkg, cfg, dom, fields, k, w, args.seed)
        sample = (f"full {args.size}x{args.size} tile (n={cfg['n']}), {k} timed model steps after "
                  f"{w} warm-up (of the requested {args.steps}/{args.warmup}: bounded CPU sample)")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": k, "warmup": w, "ms_per_step": sps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload, "cells_per_gpu": cfg["n"], "river_cells": cfg["nriv"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------------ B200 arm ----
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    if args.partition:
        # every rank derives the same global domain and keeps its own basins
        gcfg, gdom, gfields = pkg.synthetic.make_basin(args.partition, args.partition,
                                                       seed=args.seed, adaptive=ADAPTIVE)
        shard = pkg.partition.partition_basins(gdom, world)[rank]
        cfg = pkg.partition.shard_config(gcfg, shard)
        dom = pkg.partition.shard_domain(gdom, shard)
        fields = pkg.partition.shard_fields(gfields, dict(pkg._lib.field_table()), shard)
        del gcfg, gdom, gfields
        workload = workload.replace(f"synthetic {args.size}x{args.size} D8 basin per GPU",
                                    f"ONE synthetic {args.partition}x{args.partition} D8 raster "
                                    f"sharded by whole drainage basins over {world} GPU(s)")
    else:
        cfg, dom, fields = build_tile(pkg, args.size, rank, args.seed)
    n, nriv, N, dt = cfg["n"], cfg["nriv"], cfg["N"], cfg["dt"]
    for kv in args.cfg:
        k, v = kv.split("=")
        cfg[k] = int(v)
    model = pkg.SbmModel(cfg, dom, fields, device=local)
    for kv in args.option:
        k, v = kv.split("=")
        model.set_option(k, int(v))
    gid = dom["gid"]
    if world > 1 or args.no_cpu_baseline:
        fields = None  # only the cpu_baseline leg needs the host copies again
    # the step's inputs wait in page-locked host memory (the contract's e2e leg copies them from
    # there): the library then copies them straight to the device, without its staging memcpy
    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    # a pool of distinct forcing fields, cycled (bounds the pinned memory of large tiles)
    n_forcing = min(args.warmup + args.steps, max(4, int(2.0e9 // (24 * max(len(gid), 1)))))
    forcing = [tuple(pin(a) for a in pkg.synthetic.make_forcing(args.seed, s, gid, dt))
               for s in range(n_forcing)]

    def barrier():
        model.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: spin the model up so that soil, overland and river stores are active
    for s in range(args.warmup):
        model.set_forcing(*forcing[s % len(forcing)])
        model.update_model(dt)
    model.synchronize()
    launches0 = model.stats()["kernel_launches"]

    # ---- leg 1: device-resident inputs (forcing of the last warm-up step stays in HBM) -----
    sampler = ClockSampler(local)
    sampler.start()
    model.set_timing(True)
    barrier()
    model.timer_start()
    for s in range(args.steps):
        model.update_model(dt)
    ms = model.timer_stop()
    barrier()
    st = model.stats()
    model.set_timing(False)
    launches = st["kernel_launches"] - launches0
    t_max = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_max = float(t_max.item())

    # ---- leg 2: end to end through the public API with HOST buffers ------------------------
    out = None
    barrier()
    # Double-buffered like a driver that reads step s + 1 while step s runs: the H2D copy of the
    # NEXT step's forcing is issued right after the step has been enqueued and overlaps its
    # kernels; every step's inputs are still copied inside the timed region (K copies for K steps).
    t0 = time.perf_counter()
    model.set_forcing(*forcing[(args.warmup) % len(forcing)])            # H2D, step 0
    for s in range(args.steps):
        model.update_model(dt)                           # asynchronous
        if s + 1 < args.steps:
            model.set_forcing(*forcing[(args.warmup + s + 1) % len(forcing)])  # H2D, step s + 1
        out = model.get("riv_q_average")                 # D2H of the step's result
    model.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e2e.item())
    assert out is not None and np.isfinite(out).all()

    total_cells = torch.tensor([float(n)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_cells, op=dist.ReduceOp.SUM)
    cells = float(total_cells.item())

    value = cells * args.steps / (ms_max * 1e-3)
    e2e_value = cells * args.steps / e2e_s
    peak, peak_src = measured_peak()
    k = max(st["timed_steps"], 1)
    v1_ms = st["ms_land_hydrology"] / k
    v1_bytes = v1_bytes_per_cell(N, cfg) * n
    achieved = v1_bytes / (v1_ms * 1e-3) / 1e9 if v1_ms > 0 else 0.0
    stage_ms = {kk[3:]: st[kk] / k for kk in st if kk.startswith("ms_")}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.partition else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload, "cells_per_gpu": n, "river_cells_per_gpu": nriv,
                   "parallelism": (f"{world} shards of whole drainage basins (greedy LPT), no "
                                   "data-path collective" if args.partition else
                                   f"{world} x disjoint sub-catchment tiles, no collective"),
                   "l2_policy": "working set (~1.9 kB/cell x 1e6 cells = 1.9 GB) exceeds the "
                                "126 MB L2; no explicit flush",
                   "wave_levels_land": st["wave_levels_land"],
                   "wave_levels_river": st["wave_levels_river"],
                   "substeps": [st["substeps_land"], st["substeps_river"], st["substeps_ssf"]]},
        "roofline": {"bound": "hbm", "kernel": "update_land_hydrology_model! = land_surface_kernel<4> + unsaturated-zone "
                               "loop engine + soil_column_kernel<4> (SBM vertical, V1)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_cell": v1_bytes_per_cell(N, cfg),
                     "ms_per_launch": v1_ms,
                     # DRAM bytes of the same kernels from the ncu --set full captures under
                     # profiles/ (r1e: land_surface 719 MB + loop engine 71 MB, r1c: soil_column
                     # 687 MB per 1e6 cells and step), over the live-measured duration
                     "traffic": (V1_DRAM_BYTES_PER_CELL * n / (v1_ms * 1e-3) / 1e9
                                 if v1_ms > 0 and N == 4 and cfg["gash"] and cfg["snow"] else None),
                     "traffic_bytes_per_cell": V1_DRAM_BYTES_PER_CELL,
                     "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, "
                                       "profiles/r1e_vertical_ncu.md + r1c_vertical_ncu.md"},
        "stage_ms_per_step": stage_ms,
        "stage_note": ("subsurface, soil_storage, overland and river run overlapped (subsurface "
                       "sweep + surface kernel on disjoint SMs, update_soil_water_storage! inside "
                       "the sweep): their time is booked under 'subsurface'"
                       if stage_ms.get("overland", 1.0) < 0.02 else
                       "overland and river run in one kernel, booked under 'overland'"
                       if stage_ms.get("river", 1.0) < 0.02 else "stages run back to back"),
        "routing": {"newton_calls": st["newton_calls_land"] + st["newton_calls_river"],
                    "newton_iters_booked": st["newton_iters_land"] + st["newton_iters_river"]},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 24 * n,
                "d2h_bytes_per_step": 8 * nriv, "ms_per_step": e2e_s / args.steps * 1e3},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # same regime as the GPU's timed steps: start the CPU model from the GPU's spun-up state
        warm = dict(fields)
        for name in model.field_names():
            warm[name] = model.get(name)
        cv, cores, sps = run_cpu(pkg, cfg, dom, warm, args.cpu_steps, 1, args.seed,
                                 first_step=args.warmup + args.steps)
        line["cpu_baseline"] = {
            "value": cv, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"full {args.size}x{args.size} tile, {args.cpu_steps} model steps after 1 "
                      f"warm-up, started from the GPU model's spun-up state ({sps:.2f} s/step); C/OpenMP port of the Julia algorithm, "
                      "threaded over the reference's sub-domain partition (Julia not installed)"}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
