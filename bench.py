#!/usr/bin/env python
"""bench.py -- cell-timesteps/s of the wflow_sbm hot path (SBM vertical + kinematic-wave
routing) on B200, with the HBM roofline of the vertical update and the CPU baseline beside it.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU implementation (oracle port)

One "step" = one model time step (update_model!, sbm_model.jl:60-92) of a synthetic D8 basin:
forcing -> SBM vertical update -> subsurface / overland / river kinematic wave (24 + 96 + 1
internal sub-steps at the reference's default fixed internal time steps) -> storages.
N = 1 runs BASELINE.json configs[1] (synthetic 1000 x 1000 basin); N > 1 runs one such
sub-catchment tile per GPU (disjoint catchments, no data-path collective: weak scaling) and, at
N = 8, additionally measures configs[2]: ONE 10000 x 10000 raster (10^8 cells) of sub-catchments
sharded over the GPUs, each rank generating only its own shard from (seed, global cell id).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg  # noqa: E402

METRIC = "cell-timesteps/s (SBM vertical + kinwave)"
UNIT = "cell-timesteps/s"
# DRAM bytes per cell of the vertical update (ncu dram__bytes_read.sum + dram__bytes_write.sum of
# its kernels, profiles/r2final_vertical_ncu.md: 448 + 314 + 258 + 13 + 512 + 177 MB per 10^6
# cells), N = 4, Gash + snow, daily step
V1_DRAM_BYTES_PER_CELL = 1722
V1_TRAFFIC_SOURCE = "ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r2final_vertical_ncu.md"
# a realistic per-step output set (write_output, io.jl:815-899; the Moselle TOML writes ~8 vectors)
OUTPUT_FIELDS = ["riv_q_average", "riv_h", "snow_storage", "saturated_water_depth",
                 "unsaturated_store_depth", "total_storage", "olf_q_average",
                 "ssf_water_table_depth"]


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


# --------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region, through NVML (a sample every few
    milliseconds; nvidia-smi takes longer per call than a timed step)."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.sm, self.mx, self.reasons = [], None, set()
        self._halt = threading.Event()
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            self.nv = None

    def run(self):
        nv = self.nv
        if nv is None:
            return
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)}
        try:
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            self.mx = None
        while not self._halt.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.004)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def workload_text(d1, d2, adaptive, local_inertial=False):
    if MOSELLE:
        return ("synthetic Moselle-shape basin per GPU (264x291 raster, 50 063 land / ~5 809 river cells, "
                "ONE outlet, all eight LDD codes), wflow_sbm vertical + kinematic-wave "
                "river/overland/subsurface with 2 reservoirs and lateral snow transport, daily step, "
                + ("adaptive internal steps (test/sbm_config.toml)" if adaptive else
                   "fixed internal steps 3600/900/86400 s") + ", N=4 soil layers, snow on")
    if HOURLY and not local_inertial:
        return (f"synthetic {d1}x{d2} D8 basin per GPU, wflow_sbm vertical (HOURLY step: modified "
                "Rutter interception) + kinematic-wave river/overland/subsurface, "
                + ("adaptive internal steps" if adaptive else "fixed internal steps 3600/900/86400 s "
                   "(1 / 4 / 1 sub-steps per hour)") + ", N=4 soil layers, snow on")
    if local_inertial == 2:
        return (f"synthetic {d1}x{d2} D8 basin per GPU, wflow_sbm vertical + kinematic-wave "
                "subsurface + 2-D LOCAL-INERTIAL overland flow coupled to the LOCAL-INERTIAL river "
                "(subgrid channel), daily step, sub-steps min(alpha L / sqrt(g h)) of land and "
                "river, N=4 soil layers, snow on")
    if local_inertial:
        return (f"synthetic {d1}x{d2} D8 basin per GPU, wflow_sbm vertical + kinematic-wave "
                "overland/subsurface + LOCAL-INERTIAL river with 1-D floodplain, daily step, fixed "
                "internal steps 3600/86400 s (river: adaptive alpha L / sqrt(g h)), N=4 soil "
                "layers, snow on")
    return (f"synthetic {d1}x{d2} D8 basin per GPU, wflow_sbm vertical + kinematic-wave "
            "river/overland/subsurface, daily step, "
            + ("adaptive internal steps" if adaptive else "fixed internal steps 3600/900/86400 s")
            + ", N=4 soil layers, snow on")


def config_block(workload, n, nriv, world):
    """The SAME keys in both arms (the driver compares them)."""
    return {"workload": workload, "cells_per_gpu": int(n), "river_cells_per_gpu": int(nriv),
            "parallelism": f"{world} x disjoint sub-catchment tiles, no collective"}


HOURLY = False   # --hourly: BASELINE configs[4] (hourly forcing => modified Rutter interception)
MOSELLE = False  # --moselle: BASELINE configs[0] shape and switches (test/sbm_config.toml:122-130)


def build_tile(pkg, d1, d2, rank, seed, adaptive, catchment_length=0, local_inertial=False):
    """One sub-catchment tile per rank: a d1 x d2 Scheidegger forest whose cell ids are offset so
    that every tile of the global raster is a different random forest. Only the rank's own
    cells are ever generated: everything is a pure function of (seed, global cell id)."""
    extra = dict(river_routing=1, floodplain=True) if local_inertial else {}
    if local_inertial == 2:
        extra = dict(river_routing=1, land_routing=1)
    if HOURLY:
        extra["dt"] = 3600.0
    if MOSELLE:   # one outlet, all eight LDD codes, 50 063 land / ~5 809 river cells (test/bmi.jl:108-115)
        return pkg.synthetic.make_basin(264, 291, seed=seed, id_offset=rank * 264 * 291, network="dendritic",
                                        n_active=50063, n_river=5809, adaptive=adaptive, reservoirs=2,
                                        snow_transport=True, **extra)
    return pkg.synthetic.make_basin(d1, d2, seed=seed, id_offset=rank * d1 * d2, adaptive=adaptive,
                                    catchment_length=catchment_length, **extra)


# --------------------------------------------------------------------------------------------
def make_cpu_model(cfg, dom, fields):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    ora = parity.make_oracle(cfg, dom, fields)
    # all the host threads the process may use (torchrun exports OMP_NUM_THREADS=1)
    cores = ora._L.wfo_set_num_threads(len(os.sched_getaffinity(0)))
    return ora, cores


def cpu_step(pkg, ora, cfg, gid, seed, step):
    p, e, t = pkg.synthetic.make_forcing(seed, step, gid, cfg["dt"])
    ora.f["precipitation"][:], ora.f["potential_evaporation"][:], ora.f["temperature"][:] = p, e, t
    t0 = time.perf_counter()
    ora.update_model(cfg["dt"])
    return time.perf_counter() - t0


# --------------------------------------------------------------------------------------------
def measure(pkg, torch, dist, model, cfg, dom, seed, steps, warmup, world, local, e2e=True,
            first_step=0):
    """Warm-up, the device-timed leg (forcing slabs resident in HBM, a different one every step)
    and the end-to-end leg (host buffers: H2D of every step's forcing and D2H of an output set
    inside the timed region). Returns a dict."""
    n, dt = cfg["n"], cfg["dt"]
    gid = dom["gid"]

    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    depth = int(max(2, min(steps, 24, 3.0e9 // (24 * max(n, 1)))))
    forcing = [tuple(pin(a) for a in pkg.synthetic.make_forcing(seed, first_step + s, gid, dt))
               for s in range(depth)]

    def barrier():
        model.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(warmup):   # spin the model up: soil, overland and river stores become active
        model.set_forcing(*forcing[s % len(forcing)])
        model.update_model(dt)
    model.synchronize()
    launches0 = model.stats()["kernel_launches"]

    # ---- leg 1: device-resident inputs, a different forcing slab every step ---------------------
    model.forcing_ring_create(depth)
    for k in range(depth):
        model.forcing_ring_put(k, *forcing[k % len(forcing)])
    sampler = ClockSampler(local)
    sampler.start()
    model.set_timing(True)
    barrier()
    model.timer_start()
    for s in range(steps):
        model.forcing_ring_use(s % depth)
        model.update_model(dt)
    ms = model.timer_stop()
    barrier()
    st = model.stats()
    model.set_timing(False)
    launches = st["kernel_launches"] - launches0
    t_max = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    out = {"ms": float(t_max.item()), "stats": st, "launches": int(launches)}

    # ---- leg 2: end to end through the public API with HOST buffers ------------------------------
    if e2e:
        # page-locked output buffers, two of them: the writer of step s reads one while the
        # copy of step s + 1 fills the other
        osize = model.output_size(OUTPUT_FIELDS)
        obuf = [torch.empty(osize, dtype=torch.float64).pin_memory().numpy() for _ in range(2)]
        checksum = 0.0
        barrier()
        t0 = time.perf_counter()
        model.forcing_ring_put(0, *forcing[0])                      # H2D, step 0
        for s in range(steps):
            model.forcing_ring_use(s % depth)
            model.update_model(dt)                                   # asynchronous
            if s + 1 < steps:                                        # H2D of step s + 1 overlaps step s
                model.forcing_ring_put((s + 1) % depth, *forcing[(s + 1) % len(forcing)])
            if s > 0:     # the outputs of step s - 1 land while step s computes; "write" them
                model.wait_outputs()
                checksum += float(obuf[(s - 1) & 1][::4096].sum())
            model.get_fields_async(OUTPUT_FIELDS, obuf[s & 1])       # D2H of the step's output set
        model.wait_outputs()
        checksum += float(obuf[(steps - 1) & 1][::4096].sum())
        model.synchronize()
        e2e_s = time.perf_counter() - t0
        t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        assert np.isfinite(checksum) and np.isfinite(obuf[0]).all() and np.isfinite(obuf[1]).all()
        out["e2e_s"] = float(t_e2e.item())
        out["d2h_bytes"] = int(osize * 8)
    out["clocks"] = sampler.stop()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=1000, help="raster side per GPU")
    ap.add_argument("--shape", default="", metavar="D1xD2", help="raster shape per GPU (overrides --size)")
    ap.add_argument("--catchment-length", type=int, default=0,
                    help="outlet lines every so many columns (a mosaic of catchments)")
    ap.add_argument("--local-inertial", action="store_true",
                    help="BASELINE configs[3]: local-inertial river + 1-D floodplain instead of the "
                         "kinematic-wave river")
    ap.add_argument("--local-inertial-land", action="store_true",
                    help="2-D local-inertial overland flow coupled to the local-inertial river "
                         "(land_routing = river_routing = local_inertial)")
    ap.add_argument("--moselle", action="store_true",
                    help="BASELINE configs[0]: Moselle-shape dendritic basin (50 063 cells) with its "
                         "reservoirs and lateral snow transport; add --adaptive for the TOML's time stepping")
    ap.add_argument("--hourly", action="store_true",
                    help="BASELINE configs[4]: hourly model step (modified Rutter interception)")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--cpu-steps", type=int, default=2, help="oracle steps of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config3", action="store_true",
                    help="skip the 10^8-cell raster block of the 8-GPU line")
    ap.add_argument("--adaptive", action="store_true",
                    help="adaptive internal routing time steps instead of the fixed defaults")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=INT",
                    help="wflowb200_set_option (kernel organisation), e.g. vertical_graph=0 for ncu")
    ap.add_argument("--cfg", action="append", default=[], metavar="NAME=INT",
                    help="WflowB200Config tuning field, e.g. vertical_slices=1")
    args = ap.parse_args()
    global HOURLY, MOSELLE
    HOURLY = bool(args.hourly)
    MOSELLE = bool(args.moselle)
    if args.local_inertial_land:
        args.local_inertial = 2
    rank, world, local = dist_env()
    pkg = load_pkg()
    d1, d2 = (int(x) for x in args.shape.split("x")) if args.shape else (args.size, args.size)
    if MOSELLE:
        d1, d2 = 264, 291
    workload = workload_text(d1, d2, args.adaptive, args.local_inertial)

    # ------------------------------------------------------------------ reference arm ----
    if args.impl == "reference":
        if rank != 0:
            return
        cfg, dom, fields = build_tile(pkg, d1, d2, 0, args.seed, args.adaptive, args.catchment_length, args.local_inertial)
        ora, cores = make_cpu_model(cfg, dom, fields)
        t_first = cpu_step(pkg, ora, cfg, dom["gid"], args.seed, 0)   # first warm-up step, timed
        # the requested W / K when they fit ~2 minutes of CPU work, else a bounded sample
        budget = 120.0
        w = max(1, min(args.warmup, int(budget / 4 / max(t_first, 1e-3))))
        k = max(1, min(args.steps, int(budget * 3 / 4 / max(t_first, 1e-3))))
        for s in range(1, w):
            cpu_step(pkg, ora, cfg, dom["gid"], args.seed, s)
        el = sum(cpu_step(pkg, ora, cfg, dom["gid"], args.seed, w + s) for s in range(k))
        value = cfg["n"] * k / el
        sample = (f"full {d1}x{d2} tile (n={cfg['n']}), {k} timed model steps after {w} warm-up "
                  f"(requested {args.steps}/{args.warmup}); C/OpenMP port of the Julia algorithm "
                  "threaded over the reference's sub-domain partition (Julia is not installed)")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": k, "warmup": w, "ms_per_step": el / k * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_block(workload, cfg["n"], cfg["nriv"], args.gpus),
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

    cfg, dom, fields = build_tile(pkg, d1, d2, rank, args.seed, args.adaptive, args.catchment_length, args.local_inertial)
    for kv in args.cfg:
        k, v = kv.split("=")
        cfg[k] = int(v)
    n, nriv, N, dt = cfg["n"], cfg["nriv"], cfg["N"], cfg["dt"]
    model = pkg.SbmModel(cfg, dom, fields, device=local)
    for kv in args.option:
        k, v = kv.split("=")
        model.set_option(k, int(v))
    do_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    if not do_cpu:
        fields = None  # only the cpu_baseline / parity leg needs the host copies again

    m = measure(pkg, torch, dist, model, cfg, dom, args.seed, args.steps, args.warmup, world, local)
    st = m["stats"]
    total_cells = torch.tensor([float(n)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_cells, op=dist.ReduceOp.SUM)
    cells = float(total_cells.item())
    value = cells * args.steps / (m["ms"] * 1e-3)
    e2e_value = cells * args.steps / m["e2e_s"]
    peak, peak_src = measured_peak()
    k = max(st["timed_steps"], 1)
    v1_ms = st["ms_land_hydrology"] / k
    v1_bytes = v1_bytes_per_cell(N, cfg) * n
    achieved = v1_bytes / (v1_ms * 1e-3) / 1e9 if v1_ms > 0 else 0.0
    stage_ms = {kk[3:]: st[kk] / k for kk in st if kk.startswith("ms_")}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["ms"] / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_block(workload, n, nriv, world),
        "details": {"l2_policy": "working set (~1.9 kB/cell x 1e6 cells = 1.9 GB) exceeds the 126 MB "
                                 "L2; no explicit flush; a different forcing slab every step",
                    "wave_levels_land": st["wave_levels_land"],
                    "wave_levels_river": st["wave_levels_river"],
                    "substeps": [st["substeps_land"], st["substeps_river"], st["substeps_ssf"]],
                    "e2e_outputs": OUTPUT_FIELDS},
        "roofline": {"bound": "hbm",
                     "kernel": "update_land_hydrology_model! = land_hydrology_kernel<4> + "
                               "unsat_engine_kernel<4> (loop engine) + soil_column_kernel<4> "
                               "(SBM vertical, V1)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_cell": v1_bytes_per_cell(N, cfg),
                     "ms_per_launch": v1_ms,
                     "traffic": (V1_DRAM_BYTES_PER_CELL * n / (v1_ms * 1e-3) / 1e9
                                 if v1_ms > 0 and N == 4 and cfg["gash"] and cfg["snow"] else None),
                     "traffic_bytes_per_cell": V1_DRAM_BYTES_PER_CELL,
                     "traffic_source": V1_TRAFFIC_SOURCE},
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
                "d2h_bytes_per_step": m["d2h_bytes"], "ms_per_step": m["e2e_s"] / args.steps * 1e3},
        "gpu_launches": m["launches"],
        "clocks": m["clocks"],
    }

    if do_cpu:
        # The CPU model starts from the GPU's spun-up state (same regime as the timed steps); the
        # GPU then takes the same steps with the same forcing and EVERY field is compared: parity
        # at the benchmarked size, in the benchmarked state.
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import parity
        first = args.warmup + 2 * args.steps
        warm = dict(fields)
        for name in model.field_names():
            warm[name] = model.get(name)
        ora, cores = make_cpu_model(cfg, dom, warm)
        cpu_step(pkg, ora, cfg, dom["gid"], args.seed, first)          # warm-up (first touch)
        el = sum(cpu_step(pkg, ora, cfg, dom["gid"], args.seed, first + 1 + s)
                 for s in range(args.cpu_steps))
        for s in range(1 + args.cpu_steps):
            model.set_forcing(*pkg.synthetic.make_forcing(args.seed, first + s, dom["gid"], dt))
            model.update_model(dt)
        try:
            # (the explicit local-inertial schemes turn last-bit noise into 1e-6: tests/test_tolerances.py)
            rep = parity.compare_models(model, ora, outliers=(1.0, 1e-5) if args.local_inertial else None)
            line["parity_checked"] = {"worst_rel": rep.worst_rel, "n_fields": len(rep),
                                      "cells": n, "steps": 1 + args.cpu_steps,
                                      "tolerance": ("elementwise 1e-10 relative + per-element atol "
                                                    "(tests/parity.py)" +
                                                    ("; local-inertial fields 1e-5" if args.local_inertial else "")),
                                      "ok": True}
        except AssertionError as e:
            line["parity_checked"] = {"ok": False, "error": str(e)[:300]}
        line["cpu_baseline"] = {
            "value": n * args.cpu_steps / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"full {d1}x{d2} tile, {args.cpu_steps} model steps after 1 warm-up, started "
                      f"from the GPU model's spun-up state ({el / args.cpu_steps:.2f} s/step); "
                      "C/OpenMP port of the Julia algorithm, threaded over the reference's "
                      "sub-domain partition (Julia not installed)"}
    elif rank == 0:
        line["cpu_baseline"] = None
    model.close()
    del model

    # ---- BASELINE configs[2]: ONE 10000 x 10000 multi-catchment raster (10^8 cells), sharded by
    # sub-catchment. The raster is a mosaic of 64 catchments of 1250 x 1250 cells (about the size
    # of configs[1]'s basin), 8 per GPU: a rank owns the strip of 1250 x 10000 cells that holds
    # its 8 catchments and generates it alone from (seed, global cell id). `config3_deep` is the
    # same raster WITHOUT the outlet lines -- every strip one forest whose drainage paths run the
    # whole 10000 cells: the latency-bound extreme (10000 dependent wavefront levels).
    if world == 8 and not args.no_config3 and not args.adaptive:
        g1, g2 = 10000 // world, 10000
        steps3 = max(20, min(args.steps, 30))
        for key, clen, what in (
                ("config3", 1250, "ONE synthetic 10000x10000 multi-catchment raster (1e8 cells): a "
                 "mosaic of 64 catchments of 1250x1250 cells, sharded by sub-catchment over 8 GPUs "
                 "(8 catchments = one strip of 1250x10000 cells per GPU)"),
                ("config3_deep", 0, "the same raster without the outlet lines: every strip of "
                 "1250x10000 cells is a forest whose drainage paths are 10000 cells long (10000 "
                 "dependent wavefront levels, 1250 cells per level: the latency-bound extreme)")):
            cfg3, dom3, fields3 = build_tile(pkg, g1, g2, rank, args.seed + 1, False, clen)
            model3 = pkg.SbmModel(cfg3, dom3, fields3, device=local)
            del fields3
            m3 = measure(pkg, torch, dist, model3, cfg3, dom3, args.seed + 1, steps3, 5, world,
                         local, e2e=False)
            tot3 = torch.tensor([float(cfg3["n"])], dtype=torch.float64, device="cuda")
            dist.all_reduce(tot3, op=dist.ReduceOp.SUM)
            st3 = m3["stats"]
            k3 = max(st3["timed_steps"], 1)
            line[key] = {
                "workload": what + ", fixed internal steps, no data-path collective",
                "cells": float(tot3.item()), "steps": steps3, "warmup": 5,
                "ms_per_step": m3["ms"] / steps3,
                "value": float(tot3.item()) * steps3 / (m3["ms"] * 1e-3), "unit": UNIT,
                "stage_ms_per_step": {kk[3:]: st3[kk] / k3 for kk in st3 if kk.startswith("ms_")},
                "wave_levels_land": st3["wave_levels_land"],
                # the vertical update at this tile size (rank 0's 12.5e6 cells)
                "roofline_v1_frac": (v1_bytes_per_cell(cfg3["n_layers"], cfg3) * cfg3["n"] /
                                     (st3["ms_land_hydrology"] / k3 * 1e-3) / 1e9 / peak
                                     if st3["ms_land_hydrology"] > 0 else None)}
            model3.close()
            del model3, cfg3, dom3
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
