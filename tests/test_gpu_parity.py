"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs. Float64 fields within 1e-10 relative (tests/parity.py), integer fields
and indexing artefacts bit-exact."""
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
    worst = parity.compare_models(gpu, ora)
    st = gpu.stats()
    assert st["kernel_launches"] > 0
    assert st["substeps_land"] == 24 and st["substeps_river"] == 96 and st["substeps_ssf"] == 1
    print(f"worst scaled rel diff {worst:.3e}")
    _close(gpu)


def test_band_kernel_matches_oracle(pkg, monkeypatch):
    """The experimental single-sub-step subsurface kernel over bands (WFB_SSF_BANDS=1, read at
    create) gives the same fields as the chunk walk and the oracle."""
    monkeypatch.setenv("WFB_SSF_BANDS", "1")
    gpu, ora, cfg = parity.run_pair(pkg, 97, 131, steps=4, seed=21)
    parity.compare_models(gpu, ora)
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
    """device_math.cuh on the device: table-driven exp / log within 2 ulp of libdevice's (each is
    < 1 ulp from the correctly rounded value), pow(x, c) within 2e-13 relative, the guard-free
    division bit-identical to IEEE `/`, branch-free min/max identical to Julia's definition
    (NaN propagation), cld(x, 2e-4) identical to Julia's formula."""
    import ctypes as C
    out = (C.c_double * 6)()
    rc = pkg._lib.lib().wflowb200_selftest_math(0, 1 << 24, out)
    assert rc == 0
    w_exp, w_log, w_pow, w_div, w_mm, w_cld = list(out)
    print("selftest", list(out))
    assert w_exp <= 2 and w_log <= 2
    assert w_pow <= 2e-13
    assert w_div == 0 and w_mm == 0 and w_cld == 0


@pytest.mark.parametrize("env", [{"WFB_FUSE_SURFACE": "0"}, {"WFB_FUSE_ROUTING": "1"},
                                 {"WFB_SSF_S1": "1"}, {"WFB_PIECE_LAND": "6", "WFB_V_SLICES": "4"},
                                 {"WFB_FUSE_SOIL_STORAGE": "0"}, {"WFB_OVERLAP_SSF": "0"},
                                 {"WFB_OVERLAP_SSF": "1", "WFB_SSF_OVERLAP_SMS": "40"}])
def test_alternative_kernel_paths_match_oracle(pkg, monkeypatch, env):
    """Every kernel organisation behind the same entry point gives the same fields: overland
    and river as separate kernels (the default fuses them), subsurface + soil storage + overland
    + river in one kernel, the slim single-sub-step subsurface node, multi-piece chunks with a
    sliced vertical update (all read at create)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    gpu, ora, cfg = parity.run_pair(pkg, 97, 131, steps=4, seed=29)
    parity.compare_models(gpu, ora)
    _close(gpu)


def test_fine_grained_entry_points_match_oracle(pkg):
    gpu, ora, cfg = parity.run_pair(pkg, 40, 50, steps=2, seed=3, fine_grained=True)
    parity.compare_models(gpu, ora)
    _close(gpu)


def test_newton_iteration_counts_match(pkg):
    gpu, ora, cfg = parity.run_pair(pkg, 64, 96, steps=3, seed=11)
    st = gpu.stats()
    L = ora._L
    import ctypes as C
    # oracle counters live in the C struct; exposed through the model's python mirror
    o = ora.newton_stats()
    assert st["newton_calls_land"] == o["newton_calls_land"]
    assert st["newton_calls_river"] == o["newton_calls_river"]
    # same number of Newton iterations (north_star); libm last-bit differences may move an
    # iterate across the 1e-12 residual threshold for a handful of nodes
    for k in ("land", "river"):
        a, b = st[f"newton_iters_{k}"], o[f"newton_iters_{k}"]
        assert abs(a - b) <= 1e-4 * max(b, 1), (k, a, b)
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
