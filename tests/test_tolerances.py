"""The per-field tolerances of tests/parity.py, calibrated on the CPU: the oracle against ITSELF
on a libm that is noisy by +-1 ulp in exp / log / cbrt (oracle/wfo_math.h: WFO_ALT_LIBM) brackets
what two faithful math libraries (glibc here, libdevice on the GPU, Julia's in the reference) may
do to the results. It must pass the elementwise test; a relative error of 1e-8 planted in one
element of any field must fail it."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity  # noqa: E402


def _pair(pkg, d1, d2, steps, seed=42, **kw):
    cfg, dom, fields = pkg.synthetic.make_basin(d1, d2, seed=seed, **kw)
    nets = parity.oracle_networks(cfg, dom)
    a = parity.make_oracle(cfg, dom, fields, nets)
    b = parity.make_oracle(cfg, dom, fields, nets, variant="alt")
    parity.step_models(pkg, (a, b), dom, cfg, seed, steps)
    return a, b, cfg


@pytest.mark.parametrize("d1,d2,steps,kw", [
    (160, 240, 6, {}),
    (120, 150, 4, dict(dt=3600.0, snow=False)),
    (150, 170, 4, dict(network="dendritic", n_active=20000, n_river=2300)),
    (150, 170, 3, dict(network="dendritic", n_active=20000, n_river=2300, adaptive=True)),
])
def test_noisy_libm_oracle_passes_elementwise(pkg, d1, d2, steps, kw):
    a, b, cfg = _pair(pkg, d1, d2, steps, **kw)
    rep = parity.compare_models(b, a)
    print(rep.summary())
    assert rep.worst_rel <= parity.RTOL
    # the Newton iteration totals of two faithful libms are close, not equal
    sa, sb = a.newton_stats(), b.newton_stats()
    for k in ("newton_calls_land", "newton_calls_river", "substeps_land", "substeps_river"):
        assert sa[k] == sb[k]


def test_local_inertial_amplifies_last_bit_noise(pkg):
    """Why the local-inertial river flow is not held to 1e-10: the explicit scheme with its
    wet/dry thresholds (surface_staggered_scheme.jl:359-380) turns the +-1 ulp noise of a faithful
    libm into visible differences -- the oracle against itself on the noisy libm, same sub-step
    counts, differs by more than 1e-8 in the river depths after the first day and still passes at
    1e-5 after four."""
    a, b, cfg = _pair(pkg, 70, 110, 1, seed=43, river_routing=1, reservoirs=4)
    h0, h1 = a.f["riv_h"], b.f["riv_h"]
    m = h0 > 1e-6
    assert float(np.max(np.abs(h1[m] - h0[m]) / h0[m])) > 1e-8
    assert a.newton_stats()["substeps_river"] == b.newton_stats()["substeps_river"] > 100
    a, b, cfg = _pair(pkg, 70, 110, 4, seed=43, river_routing=1, reservoirs=4)
    rep = parity.compare_models(b, a, outliers=(1.0, 1e-5))
    print(rep.summary())


@pytest.mark.parametrize("reservoirs", [0, 3])
def test_local_inertial_land_tolerance_is_calibrated(pkg, reservoirs):
    """The tolerance of the GPU test of the 2-D local-inertial overland flow (same basin, same four
    steps): the oracle against itself on the +-1 ulp libm passes it, with some elements between
    1e-10 and 1e-5 (the scheme's thresholds at work) -- and an error of 1e-3 planted in one wet
    cell's depth does not pass."""
    a, b, cfg = _pair(pkg, 70, 110, 4, seed=43, river_routing=1, land_routing=1, reservoirs=reservoirs)
    assert a.newton_stats()["substeps_river"] == b.newton_stats()["substeps_river"] > 100
    rep = parity.compare_models(b, a, outliers=(1.0, 1e-5))
    assert sum(v.get("outliers", 0) for v in rep.values()) > 0
    k = int(np.argmax(a.f["olf_h"]))
    b.f["olf_h"][k] *= 1.0 + 1e-3
    with pytest.raises(AssertionError, match="olf_h"):
        parity.compare_models(b, a, outliers=(1.0, 1e-5))


def test_planted_error_is_caught(pkg):
    """A relative error of 1e-8 in ONE element fails the comparison, whatever the field's largest
    magnitude is: for every flux / storage / discharge field, at its smallest element that lies
    above the field's absolute tolerance."""
    a, b, cfg = _pair(pkg, 48, 64, 3)
    rli = a.f["river_land_indices"]
    caught = skipped = 0
    for name in list(a.f):
        o = a.f[name]
        if o.dtype.kind != "f" or name in ("precipitation", "potential_evaporation", "temperature"):
            continue
        c = dict(a.cfg)
        st = a.newton_stats()
        c["S_land"], c["S_river"] = st["substeps_land"], st["substeps_river"]
        atol = np.broadcast_to(parity.atol_of(name, lambda n: a.f[n], c, o.shape, parity.RTOL, rli), o.shape)
        ok = np.isfinite(o) & (np.abs(o) * 1e-8 > 4.0 * (atol + parity.RTOL * np.abs(o)))
        if not ok.any():
            skipped += 1
            continue
        k = np.unravel_index(np.argmin(np.where(ok, np.abs(o), np.inf)), o.shape)
        saved = b.f[name][k]
        b.f[name][k] = o[k] * (1.0 + 1e-8)
        with pytest.raises(AssertionError):
            parity.compare_models(b, a, names=[name])
        b.f[name][k] = saved
        caught += 1
    print(f"planted errors caught in {caught} fields ({skipped} fields hold no value above their floor)")
    assert caught > 120
