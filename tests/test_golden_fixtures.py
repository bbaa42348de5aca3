"""Committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the
CPU oracle): the oracle must keep reproducing them (CPU), and the CUDA path, called through the
C ABI, must match them within the north-star tolerance (GPU)."""
import os
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import parity  # noqa: E402
import make_golden  # noqa: E402

CASES = sorted(make_golden.CASES)


def _fixture(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden_fixture(pkg, name):
    want = _fixture(name)
    cfg, dom, fields, ora = make_golden.run_case(pkg, name)
    assert len(want) >= 150
    for k, w in want.items():
        g = np.asarray(ora.f[k])
        assert g.shape == w.shape, k
        if g.dtype.kind == "i":
            assert np.array_equal(g, w), k
        else:
            assert np.allclose(g, w, rtol=1e-12, atol=0.0, equal_nan=True), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_matches_golden_fixture(pkg, name):
    d1, d2, steps, seed, kw = make_golden.CASES[name]
    cfg, dom, fields = pkg.synthetic.make_basin(d1, d2, seed=seed, **kw)
    gpu = pkg.SbmModel(cfg, dom, fields)
    dt = cfg["dt"]
    for step in range(steps):
        gpu.set_forcing(*pkg.synthetic.make_forcing(seed, step, dom["gid"], dt))
        gpu.update_model(dt)
    gpu.synchronize()
    fx = _fixture(name)
    fx["river_land_indices"] = dom["river_land_indices"] - 1
    st = gpu.stats()
    golden = types.SimpleNamespace(f=fx, cfg=cfg, newton_stats=lambda: st)
    rep = parity.compare_models(gpu, golden, rtol=parity.RTOL,
                                names=[n for n in gpu.field_names() if n in fx])
    assert len(rep) >= 140
    print(f"{name}: {rep.summary()}")
    gpu.close()
