"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs. Float64 fields ELEMENTWISE within 1e-10 relative plus a per-element
absolute tolerance derived from the operands of the reference expression (tests/parity.py;
calibrated by tests/test_tolerances.py), integer fields and indexing artefacts bit-exact."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity  # noqa: E402

pytestmark = pytest.mark.gpu


def _close(*models):
    for m in models:
        if hasattr(m, "close"):
            m.close()


@pytest.mark.parametrize("d1,d2,steps", [(48, 64, 3), (97, 131, 4)])
def test_update_model_matches_oracle(pkg, d1, d2, steps):
    gpu, ora, cfg = parity.run_pair(pkg, d1, d2, steps=steps, seed=42)
    rep = parity.compare_models(gpu, ora)
    st = gpu.stats()
    assert st["kernel_launches"] > 0
    assert st["substeps_land"] == 24 and st["substeps_river"] == 96 and st["substeps_ssf"] == 1
    print(rep.summary())
    _close(gpu)


def test_benchmark_size_1000x1000_all_defaults(pkg):
    """BASELINE configs[1], the raster bench.py times: 10^6 cells, 1000 wavefront levels, every
    default (overlapped subsurface / surface kernels, sliced vertical update with the tile order
    of the previous step, ~80 000 suspended Brooks-Corey loops per step once the soil is wet)."""
    gpu, ora, cfg = parity.run_pair(pkg, 1000, 1000, steps=4, seed=42, newton_trace=True)
    assert cfg["n"] == 1000000
    rep = parity.compare_models(gpu, ora)
    print(rep.summary())
    print("newton parity (nodes, nodes with a differing iteration total, gpu, oracle):",
          parity.newton_parity(gpu, ora))
    _close(gpu)


def test_wide_raster_takes_the_wide_domain_defaults(pkg):
    """2304 nodes per wavefront level: multi-piece land chunks (depth 6), no subsurface / surface
    overlap, update_soil_water_storage! as a kernel of its own -- chosen by the library, no
    option set."""
    gpu, ora, cfg = parity.run_pair(pkg, 2304, 64, steps=3, seed=17)
    a = gpu.artifacts("land")
    assert cfg["n"] // (len(a["wave_level_ptr"]) - 1) >= 2048
    rep = parity.compare_models(gpu, ora)
    print(rep.summary())
    _close(gpu)


@pytest.mark.parametrize("adaptive", [False, True])
def test_moselle_shape_dendritic_network(pkg, adaptive):
    """BASELINE configs[0] shape (test/bmi.jl:108-115: 50 063 land cells, ~5 809 river cells): ONE
    outlet, all eight LDD codes, node ids unrelated to the drainage order (steepest descent on a
    smooth random surface), fixed and adaptive internal time steps (test/sbm_config.toml:123)."""
    gpu, ora, cfg = parity.run_pair(pkg, 264, 291, steps=4, seed=7, network="dendritic",
                                    n_active=50063, n_river=5809, adaptive=adaptive,
                                    newton_trace=True)
    assert cfg["n"] == 50063 and abs(cfg["nriv"] - 5809) < 60
    ldd = gpu.artifacts("land")["ldd"]
    assert set(np.unique(ldd)) == set(range(1, 10)) and (ldd == 5).sum() == 1
    rep = parity.compare_models(gpu, ora)
    st, o = gpu.stats(), ora.newton_stats()
    for k in ("substeps_land", "substeps_river", "substeps_ssf"):
        assert st[k] == o[k], (k, st[k], o[k])
    print(rep.summary())
    print("sub-steps", {k: st[k] for k in st if k.startswith("substeps")},
          "newton parity:", parity.newton_parity(gpu, ora))
    _close(gpu)


@pytest.mark.parametrize("dt", [86400.0, 3600.0])
def test_adaptive_internal_time_steps_match_oracle(pkg, dt):
    """kinematic_wave__adaptive_time_step_flag = true: the sub-step lengths come from the
    type-7 quantile (surface, surface_kinwave.jl:674-704) / the minimum (subsurface,
    lateral_subsurface_flow.jl:314-344) of the per-node Courant steps, evaluated on the device
    after every sub-step; same number of sub-steps and same fields as the oracle."""
    gpu, ora, cfg = parity.run_pair(pkg, 40, 56, steps=5 if dt > 4000.0 else 3, seed=13,
                                    adaptive=True, dt=dt, snow=dt > 4000.0)
    parity.compare_models(gpu, ora)
    st, o = gpu.stats(), ora.newton_stats()
    for k in ("substeps_land", "substeps_river", "substeps_ssf"):
        assert st[k] == o[k], (k, st[k], o[k])
    print({k: st[k] for k in st if k.startswith("substeps")})
    if dt > 4000.0:  # the daily step needs several river / subsurface sub-steps
        assert st["substeps_land"] + st["substeps_river"] + st["substeps_ssf"] > 3
    _close(gpu)


def test_device_math_selftest(pkg):
    """device_math.cuh on the device: the loop engine's tracked-power trips within 1e-12 of the
    reference loop, pow(x, c) = exp(c log x) within 2e-13 relative of libdevice's pow, the guard-free
    division bit-identical to IEEE `/`, branch-free min/max identical to Julia's definition
    (NaN propagation), cld(x, 2e-4) identical to Julia's formula."""
    import ctypes as C
    out = (C.c_double * 6)()
    rc = pkg._lib.lib().wflowb200_selftest_math(0, 1 << 24, out)
    assert rc == 0
    w_exp, w_log, w_pow, w_div, w_mm, w_cld = list(out)
    print("selftest", list(out))
    assert w_exp <= 1e-12 and w_log <= 1e-12   # fast trips of the loop engine vs the reference loop
    assert w_pow <= 2e-13
    assert w_div == 0 and w_mm == 0 and w_cld == 0


@pytest.mark.parametrize("options,cfg_over", [
    ({"fuse_surface": 0}, {}),
    ({}, {"wave_piece_depth_land": 6}),
    ({}, {"wave_piece_depth_land": 2}),
    ({"fuse_surface": 0}, {"wave_piece_depth_land": 3}),
    ({"fuse_soil_storage": 0}, {"unsat_inline_iters": 8}),
    ({"overlap_subsurface": 0}, {"wave_piece_depth_land": -1}),
    ({"overlap_subsurface": 1, "overlap_subsurface_sms": 40}, {}),
    ({"vertical_graph": 0, "surface_river_period": 4, "surface_river_share": 2}, {}),
])
def test_alternative_kernel_paths_match_oracle(pkg, options, cfg_over):
    """Every kernel organisation behind the same entry point gives the same fields: overland
    and river as separate kernels (the default fuses them), multi-piece chunks, a long in-line
    loop limit in the vertical update (fewer suspended cells), the subsurface sweep with /
    without the fused soil-water storage and the overlapped surface kernel, the vertical update
    launch by launch instead of as a CUDA graph."""
    gpu, ora, cfg = parity.run_pair(pkg, 97, 131, steps=4, seed=29, cfg_over=cfg_over,
                                    options=options)
    parity.compare_models(gpu, ora)
    _close(gpu)


def test_fine_grained_entry_points_match_oracle(pkg):
    gpu, ora, cfg = parity.run_pair(pkg, 40, 50, steps=2, seed=3, fine_grained=True)
    parity.compare_models(gpu, ora)
    _close(gpu)


@pytest.mark.parametrize("root_each", [0, 1])
def test_newton_iteration_counts(pkg, root_each):
    """`kinematic_wave` (surface_process.jl:24-70): the number of solves is equal, and the Newton
    iteration totals are compared NODE BY NODE (exact integers, no tolerance on the totals). An
    iterate whose residual lies within a few ulp of the 1e-12 threshold may stop one iteration
    earlier or later under a different (faithful) libm -- the oracle against itself on a +-1 ulp
    noisy libm differs in ~2e-5 of the iterations (tests/test_tolerances.py) -- so a small number
    of nodes may differ; it is printed and bounded. root_each = 1: the fifth root of the previous
    discharge is evaluated before every solve like the reference does, instead of carried."""
    gpu, ora, cfg = parity.run_pair(pkg, 64, 96, steps=3, seed=11, newton_trace=True,
                                    options={"kinwave_root_each_substep": root_each})
    st, o = gpu.stats(), ora.newton_stats()
    assert st["newton_calls_land"] == o["newton_calls_land"]
    assert st["newton_calls_river"] == o["newton_calls_river"]
    np_ = parity.newton_parity(gpu, ora)
    print("newton parity (nodes, nodes with a differing iteration total, gpu total, oracle total):", np_)
    for dom, (nodes, differ, gsum, osum) in np_.items():
        assert differ <= max(2, nodes // 200), (dom, nodes, differ)
        assert abs(gsum - osum) <= max(8, differ * 8), (dom, gsum, osum)
    parity.compare_models(gpu, ora)
    _close(gpu)


def test_hourly_rutter_no_snow_variant(pkg):
    gpu, ora, cfg = parity.run_pair(pkg, 50, 70, steps=3, seed=5, dt=3600.0, snow=False)
    assert cfg["gash"] == 0
    parity.compare_models(gpu, ora)
    st = gpu.stats()
    assert st["substeps_land"] == 1 and st["substeps_river"] == 4
    _close(gpu)


def test_glacier_exponential_constant_infiltration_reduction(pkg):
    gpu, ora, cfg = parity.run_pair(pkg, 40, 60, steps=3, seed=8, glacier=True, kv_profile=1,
                                    soil_infiltration_reduction=True, external_inflow=True)
    parity.compare_models(gpu, ora)
    _close(gpu)


@pytest.mark.parametrize("mode", ["update_model", "fine_grained", "adaptive", "separate_kernels"])
def test_reservoirs_on_the_river(pkg, mode):
    """reservoir__flag = true (test/sbm_config.toml:126): reservoir outlets are dropped from the
    upstream lists of the land and river kinematic waves (domain.jl:96-122), the reservoir takes
    the river, overland and subsurface flow of its outlet cell and hands its outflow to the
    downstream river node (surface_kinwave.jl:441-489). Simple, modified-Puls, free-weir and
    observed-outflow reservoirs, with external abstraction / supply (reservoir.jl:389-634)."""
    kw = dict(reservoirs=6)
    opts = {}
    if mode == "adaptive":
        kw["adaptive"] = True
    if mode == "separate_kernels":
        opts = {"fuse_surface": 0}
    gpu, ora, cfg = parity.run_pair(pkg, 90, 120, steps=4, seed=31, fine_grained=mode == "fine_grained",
                                    options=opts, **kw)
    assert cfg["nres"] == 6
    rep = parity.compare_models(gpu, ora)
    assert rep["res_outflow"]["rel"] <= parity.RTOL and np.all(ora.f["res_outflow_average"] > 0)
    print(rep.summary(), "outflow", ora.f["res_outflow_average"])
    _close(gpu)


@pytest.mark.parametrize("kv_profile", [2, 3])
def test_layered_conductivity_profiles(pkg, kv_profile):
    """KvLayered / KvLayeredExponential (soil.jl:213-244, utils.jl:763-789) in the unsaturated
    zone, capillary rise and leakage; kh_layered_profile! (utils.jl:792-895) and
    kinematic_wave_ssf(::KhLayered) (subsurface_process.jl:47-51,183-228) in the subsurface flow."""
    gpu, ora, cfg = parity.run_pair(pkg, 60, 84, steps=4, seed=37, kv_profile=kv_profile)
    rep = parity.compare_models(gpu, ora)
    assert "ssf_kh" in rep and rep["ssf_kh"]["rel"] <= parity.RTOL
    print(rep.summary())
    _close(gpu)


@pytest.mark.parametrize("reservoirs", [0, 5])
def test_lateral_snow_transport(pkg, reservoirs):
    """snow_gravitational_transport__flag = true (test/sbm_config.toml:124): accucapacityflux of
    snow storage and snow water over the land network between the snow and the glacier model
    (sbm.jl:98-100, surface_process.jl:9-19, routing/utils.jl:82-168), also together with
    reservoirs (whose outlets stay in the transport graph but leave the upstream lists)."""
    gpu, ora, cfg = parity.run_pair(pkg, 70, 96, steps=4, seed=41, snow_transport=True,
                                    glacier=True, reservoirs=reservoirs)
    assert cfg["snow_transport"] == 1 and float(np.max(ora.f["snow_out"])) > 0.0
    rep = parity.compare_models(gpu, ora)
    print(rep.summary(), "max snow_out", float(np.max(ora.f["snow_out"])))
    _close(gpu)


@pytest.mark.parametrize("reservoirs", [0, 4])
def test_local_inertial_river_flow(pkg, reservoirs):
    """river_routing = "local_inertial" (BASELINE config #4, river part): adaptive sub-steps
    dt_s = alpha min(L / sqrt(g h)), edge flow, reservoirs as boundary conditions, node depth and
    storage (surface_staggered_scheme.jl:326-383,627-661,723-759,800-838,1004-1020); all sub-steps
    of a model step run inside one persistent kernel. Same number of sub-steps as the oracle."""
    gpu, ora, cfg = parity.run_pair(pkg, 70, 110, steps=4, seed=43, river_routing=1,
                                    reservoirs=reservoirs)
    # The explicit scheme with its wet/dry thresholds (h_thresh, h <= 0 limiters, Froude limit)
    # amplifies last-bit differences: the oracle against ITSELF on a +-1 ulp noisy libm differs by
    # up to 8e-3 while the dry channels fill and by ~1e-6 after four days, in nearly every river
    # element (tests/test_tolerances.py::test_local_inertial_amplifies_last_bit_noise). So river
    # fields are held to 1e-5 here; everything the river does not touch stays at 1e-10.
    rep = parity.compare_models(gpu, ora, outliers=(1.0, 1e-5))
    st, o = gpu.stats(), ora.newton_stats()
    assert abs(st["substeps_river"] - o["substeps_river"]) <= 1 and o["substeps_river"] > 20
    assert float(np.max(ora.f["riv_q_average"])) > 0.0
    print(rep.summary(), "river sub-steps", st["substeps_river"], o["substeps_river"])
    _close(gpu)


@pytest.mark.parametrize("reservoirs", [0, 3])
def test_local_inertial_river_with_floodplain(pkg, reservoirs):
    """floodplain_1d__flag with the local-inertial river (BASELINE config #4): floodplain flow at
    the edges over the six-level FloodPlainProfile tables, the opposite-direction rule, bankfull
    redistribution between channel and floodplain, reservoir inflow including the floodplain's
    (surface_staggered_scheme.jl:291-299,440-533,674-712,826-835; floodplain.jl:287-354) -- in the
    same two phases of the persistent kernel. Tolerances as for the river alone (chaotic scheme)."""
    gpu, ora, cfg = parity.run_pair(pkg, 70, 110, steps=4, seed=43, river_routing=1,
                                    floodplain=True, reservoirs=reservoirs)
    assert len(cfg["fp_depth"]) == 6
    rep = parity.compare_models(gpu, ora, outliers=(1.0, 1e-5))
    st, o = gpu.stats(), ora.newton_stats()
    assert abs(st["substeps_river"] - o["substeps_river"]) <= 1 and o["substeps_river"] > 20
    wet = int((ora.f["fp_h"] > 0.0).sum())
    assert wet > 10 and float(np.max(np.abs(ora.f["fp_q_average"]))) > 0.0
    assert np.array_equal(gpu.get("fp_h") > 0.0, ora.f["fp_h"] > 0.0)
    print(rep.summary(), "river sub-steps", st["substeps_river"], "nodes over bank", wet)
    _close(gpu)


@pytest.mark.parametrize("reservoirs,fine_grained", [(0, False), (3, False), (3, True)])
def test_local_inertial_land_and_river_flow(pkg, reservoirs, fine_grained):
    """land_routing = river_routing = "local_inertial" (BASELINE config #4): the 2-D local-inertial
    overland flow on the staggered grid coupled to the local-inertial river with its subgrid
    channel -- dt_s = min(stable_timestep(river), stable_timestep(land)), x / y edge flows
    (de Almeida et al. 2012), overland inflow of the reservoirs, river edge flow, reservoirs, water
    depth and storage of land and river cells with bankfull spill
    (surface_staggered_scheme.jl:1022-1043,1080-1097,1153-1546; surface_process.jl:123-159;
    surface_routing.jl:62-86) -- all sub-steps of a model step inside one persistent kernel. The
    EdgeConnectivity artefact is bit-exact; tolerances as for the river alone (an explicit scheme
    with wet/dry thresholds: the oracle against itself on a +-1 ulp libm needs them too)."""
    from oracle import network as onw
    gpu, ora, cfg = parity.run_pair(pkg, 70, 110, steps=4, seed=43, river_routing=1, land_routing=1,
                                    reservoirs=reservoirs, fine_grained=fine_grained)
    assert cfg["land_routing"] == 1 and cfg["li_land_theta"] == 0.9
    dom = pkg.synthetic.make_basin(70, 110, seed=43, river_routing=1, land_routing=1,
                                   reservoirs=reservoirs)[1]
    e = onw.edge_connectivity(dom["indices"], dom["d1"], dom["d2"])
    for k in ("x_up", "x_down", "y_up", "y_down"):
        assert np.array_equal(gpu.artifact("land", "edge_" + k), e["ind_" + k]), k
    rep = parity.compare_models(gpu, ora, outliers=(1.0, 1e-5))
    st, o = gpu.stats(), ora.newton_stats()
    assert abs(st["substeps_river"] - o["substeps_river"]) <= 1 and o["substeps_river"] > 20
    assert st["substeps_land"] == st["substeps_river"]
    wet = int((ora.f["olf_h"] > 0.0).sum())
    assert wet > 100 and float(np.max(np.abs(ora.f["li_land_qx_average"]))) > 0.0
    assert float(np.max(np.abs(ora.f["li_land_qy_average"]))) > 0.0
    over_bank = int((ora.f["olf_h"][dom["river_land_indices"] - 1] > 0.0).sum())
    assert over_bank > 0                    # the subgrid channels do spill onto their cells
    n_out = sum(v.get("outliers", 0) for v in rep.values())
    print(rep.summary(), "sub-steps", st["substeps_river"], o["substeps_river"], "wet cells", wet,
          "river cells over bank", over_bank, "threshold outliers", n_out)
    _close(gpu)


def test_local_inertial_land_edge_cases(pkg):
    """The 2-D local-inertial overland flow on a masked raster (cells whose neighbours are inactive
    or outside: EdgeConnectivity n + 1), on a domain without any river cell, on one-cell and
    two-cell domains, and together with adaptive internal time steps of the subsurface flow."""
    mask = np.ones((37, 53), dtype=bool)
    mask[:5, :7] = False
    mask[20:, 40:] = False
    mask[10:14, 22:30] = False
    cases = [dict(d=(37, 53), kw=dict(mask=mask)), dict(d=(24, 30), kw=dict(river_fraction_target=0.0)),
             dict(d=(1, 1), kw={}), dict(d=(2, 1), kw={}), dict(d=(1, 40), kw={}),
             dict(d=(30, 44), kw=dict(adaptive=True))]
    for case in cases:
        gpu, ora, cfg = parity.run_pair(pkg, *case["d"], steps=3, seed=5, river_routing=1,
                                        land_routing=1, **case["kw"])
        rep = parity.compare_models(gpu, ora, outliers=(1.0, 1e-5))
        st, o = gpu.stats(), ora.newton_stats()
        assert abs(st["substeps_river"] - o["substeps_river"]) <= 1, case["d"]
        print(case["d"], "n", cfg["n"], "nriv", cfg["nriv"], rep.summary(), "sub-steps", st["substeps_river"])
        _close(gpu)


@pytest.mark.parametrize("adaptive,reservoirs", [(False, 0), (True, 0), (False, 3), (True, 3)])
def test_kinematic_wave_river_with_floodplain(pkg, adaptive, reservoirs):
    """floodplain_1d__flag with the kinematic-wave river: per sub-step the channel-floodplain
    exchange (bankfull redistribution, lateral inflow of the wave), the Manning flow capacity of
    the floodplain from its profile and accucapacityflux! of the floodplain storage
    (surface_kinwave.jl:387-432,567-601,650-659) -- inside the river's skewed wavefront, three
    published values per node and sub-step."""
    gpu, ora, cfg = parity.run_pair(pkg, 70, 110, steps=4, seed=43, floodplain=True,
                                    adaptive=adaptive, reservoirs=reservoirs)
    assert len(cfg["fp_depth"]) == 6 and cfg["river_routing"] == 0 and cfg["nres"] == reservoirs
    rep = parity.compare_models(gpu, ora)
    st, o = gpu.stats(), ora.newton_stats()
    assert st["substeps_river"] == o["substeps_river"]
    wet = int((ora.f["fp_h"] > 0.0).sum())
    assert wet > 10 and float(np.max(ora.f["fp_q_average"])) > 0.0
    assert np.array_equal(gpu.get("fp_h") > 0.0, ora.f["fp_h"] > 0.0)
    print(rep.summary(), "river sub-steps", st["substeps_river"], "nodes over bank", wet)
    _close(gpu)


def test_five_soil_layers(pkg):
    gpu, ora, cfg = parity.run_pair(pkg, 32, 48, steps=2, seed=2,
                                    soil_layer_thickness_mm=(50, 50, 300, 800))
    assert cfg["N"] == 5
    parity.compare_models(gpu, ora)
    _close(gpu)


def test_masked_raster_and_single_column(pkg):
    mask = np.ones((37, 53), dtype=bool)
    mask[:5, :7] = False
    mask[20:, 40:] = False
    gpu, ora, cfg = parity.run_pair(pkg, 37, 53, steps=2, seed=4, mask=mask)
    parity.compare_models(gpu, ora)
    _close(gpu)
    gpu, ora, cfg = parity.run_pair(pkg, 1, 40, steps=2, seed=4)  # one chain, 40 levels
    parity.compare_models(gpu, ora)
    _close(gpu)


def test_degenerate_domains(pkg):
    """No river cells at all (nriv = 0: the river kernels and the fused surface kernel are
    skipped), a one-cell domain and a two-cell column."""
    gpu, ora, cfg = parity.run_pair(pkg, 12, 16, steps=3, seed=3, river_fraction_target=0.0)
    assert cfg["nriv"] == 0
    parity.compare_models(gpu, ora)
    _close(gpu)
    for d1, d2 in ((1, 1), (2, 1)):
        gpu, ora, cfg = parity.run_pair(pkg, d1, d2, steps=2, seed=3)
        parity.compare_models(gpu, ora)
        _close(gpu)


def test_artifacts_bit_exact_on_device_handle(pkg):
    cfg, dom, fields = pkg.synthetic.make_basin(120, 200, seed=1)
    cfg["land_streamorder_min"], cfg["river_streamorder_min"] = 3, 3
    gpu = pkg.SbmModel(cfg, dom, fields)
    land, river = parity.oracle_networks(cfg, dom)
    from oracle.oracle import _csr
    for name, o in (("land", land), ("river", river)):
        a = gpu.artifacts(name)
        assert np.array_equal(a["order"], o["order"])
        assert np.array_equal(a["upstream_ptr"], o["up_ptr"])
        assert np.array_equal(a["upstream_idx"], o["up_idx"])
        sp, so = _csr(o["order_subdomain"])
        assert np.array_equal(a["subdomain_ptr"], sp) and np.array_equal(a["subdomain_order"], so)
    _close(gpu)


def test_state_roundtrip_and_set_value(pkg):
    """BMI-style set/get (bmi.jl:208-248): what is written is read back bit-exactly, for a
    scalar, a layered and a river field, through the slot permutation."""
    cfg, dom, fields = pkg.synthetic.make_basin(31, 45, seed=6)
    gpu = pkg.SbmModel(cfg, dom, fields)
    rng = np.random.default_rng(0)
    for name in ("snow_storage", "unsaturated_layer_depth", "cumulative_layer_depth", "riv_q",
                 "olf_q", "ssf_q"):
        shape = gpu._shape(name)
        a = rng.random(shape)
        gpu.set(name, a)
        assert np.array_equal(gpu.get(name), a), name
    nl = rng.integers(1, 5, size=cfg["n"])
    gpu.set("number_of_layers", nl)
    assert np.array_equal(gpu.get("number_of_layers"), nl)
    _close(gpu)


@pytest.mark.parametrize("adaptive", [False, True])
def test_two_shards_equal_one_handle_bit_for_bit(pkg, adaptive):
    """The N > 1 PRODUCT path: one domain cut into two shards of whole drainage basins
    (partition.py), each shard its own handle, against the single-handle run -- bit for bit, with
    fixed internal time steps (no exchange at all) and with adaptive ones, where the type-7
    quantile and the minimum of the Courant steps are reduced over the shards (histograms of the
    radix select, counts, minima: the same reductions run over NCCL with one process per GPU)."""
    cfg, dom, fields = pkg.synthetic.make_basin(60, 90, seed=27, adaptive=adaptive)
    dt = cfg["dt"]
    one = pkg.SbmModel(cfg, dom, fields)
    shards = pkg.partition.partition_basins(dom, 2)
    table = dict(pkg._lib.field_table())
    parts = []
    for sh in shards:
        ldom = pkg.partition.shard_domain(dom, sh)
        lcfg = pkg.partition.shard_config(cfg, sh)
        lfields = pkg.partition.shard_fields(fields, table, sh)
        lfields.pop("nlayers_kv", None)
        parts.append(pkg.SbmModel(lcfg, ldom, lfields))
    if adaptive:   # a shard refuses to run alone
        with pytest.raises(pkg.WflowB200Error):
            parts[0].update_model(dt)
    group = pkg.ShardGroup(parts)
    for step in range(3):
        p, e, t = pkg.synthetic.make_forcing(27, step, dom["gid"], dt)
        one.set_forcing(p, e, t)
        one.update_model(dt)
        for m, sh in zip(parts, shards):
            m.set_forcing(np.ascontiguousarray(p[sh.cells]), np.ascontiguousarray(e[sh.cells]),
                          np.ascontiguousarray(t[sh.cells]))
        group.step_all(dt)
    st = one.stats()
    for m in parts:
        sm = m.stats()
        for k in ("substeps_land", "substeps_river", "substeps_ssf"):
            assert sm[k] == st[k], (k, sm[k], st[k])
    for name in one.field_names():
        kind = table.get(name, 0)
        if kind in (4, 6):   # reservoirs; 2-D local-inertial overland flow (absent here)
            continue
        want = one.get(name)
        got = np.empty_like(want)
        for m, sh in zip(parts, shards):
            got[sh.river_cells if kind in (3, 5) else sh.cells] = m.get(name)
        assert np.array_equal(got, want, equal_nan=True), name
    print("sub-steps", {k: st[k] for k in st if k.startswith("substeps")})
    group.close()
    _close(one, *parts)


def test_forcing_ring_cyclic_lai_and_output_gather(pkg):
    """The steps either side of the path: forcing slabs staged ahead in an HBM ring,
    cyclic LAI slabs staged once and switched on the device, several output vectors fetched with
    one copy -- same results as set_forcing / set / get."""
    cfg, dom, fields = pkg.synthetic.make_basin(40, 56, seed=9)
    dt, n = cfg["dt"], cfg["n"]
    a = pkg.SbmModel(cfg, dom, fields)
    b = pkg.SbmModel(cfg, dom, fields)
    rng = np.random.default_rng(1)
    lai = np.stack([fields["leaf_area_index"] * (0.6 + 0.1 * m) for m in range(12)])
    b.set_cyclic_lai(lai)
    b.forcing_ring_create(3)
    forc = [tuple(np.ascontiguousarray(x) for x in pkg.synthetic.make_forcing(9, s, dom["gid"], dt))
            for s in range(5)]
    for s in range(min(3, len(forc))):
        b.forcing_ring_put(s % 3, *forc[s])
    names = ["riv_q_average", "riv_h", "snow_storage", "saturated_water_depth",
             "unsaturated_layer_depth", "total_storage", "olf_q_average", "ssf_water_table_depth"]
    for s in range(5):
        month = (s * 5) % 12
        a.set("leaf_area_index", lai[month])
        a.set_forcing(*forc[s])
        a.update_model(dt)
        b.use_cyclic_lai(month)
        b.forcing_ring_use(s % 3)
        b.update_model(dt)
        if s + 3 < len(forc):
            b.forcing_ring_put(s % 3, *forc[s + 3])   # refill the slot just consumed
        got = b.get_fields(names)
        for nm in names:
            assert np.array_equal(got[nm], a.get(nm), equal_nan=True), (s, nm)
    _close(a, b)


def test_water_balance_closure(pkg):
    """The reference's data-free invariant (test/run_sbm.jl:1238-1264): the storage change of
    every component closes against its fluxes. Checked on the GPU fields alone."""
    cfg, dom, fields = pkg.synthetic.make_basin(64, 80, seed=13)
    dt = cfg["dt"]
    gpu = pkg.SbmModel(cfg, dom, fields)
    for step in range(3):
        gpu.set_forcing(*pkg.synthetic.make_forcing(13, step, dom["gid"], dt))
        gpu.update_model(dt)
    s0 = gpu.get("olf_storage")
    r0 = gpu.get("riv_storage")
    gpu.set_forcing(*pkg.synthetic.make_forcing(13, 3, dom["gid"], dt))
    gpu.update_model(dt)
    # overland: dS = (qin_avg + inwater - q_avg) * dt ... to_river leaves through the receiving
    # node's accounting, so only the balance of river reaches is exact per node:
    r1 = gpu.get("riv_storage")
    qin, q, inw = gpu.get("riv_qin_average"), gpu.get("riv_q_average"), gpu.get("riv_inwater")
    err = (r1 - r0) - (qin + inw - q) * dt
    scale = np.maximum(np.abs(q * dt), 1.0)
    assert np.max(np.abs(err) / scale) < 1e-6
    _close(gpu)
