"""Generates tests/golden/sbm_*.npz: every field of the CPU oracle (oracle/, the restatement of
the reference algorithm pinned by tests/test_oracle_golden.py) after a few model steps of a small
seeded synthetic basin. The reference itself (Julia) cannot run in this image, so these are
ORACLE outputs, committed so that (a) a change of the oracle shows up as a diff of a tracked
fixture and (b) the CUDA path is also compared with bytes that do not depend on the oracle
being rebuilt on the GPU box.

    python tests/golden/make_golden.py      # rewrites the fixtures
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (d1, d2, steps, seed, make_basin keyword arguments)
    "sbm_daily_gash_snow_24x32": (24, 32, 3, 7, {}),
    "sbm_hourly_rutter_20x28": (20, 28, 4, 11, dict(dt=3600.0, snow=False)),
    "sbm_adaptive_18x30": (18, 30, 3, 19, dict(adaptive=True)),
}


def run_case(pkg, name):
    import parity
    d1, d2, steps, seed, kw = CASES[name]
    cfg, dom, fields = pkg.synthetic.make_basin(d1, d2, seed=seed, **kw)
    ora = parity.make_oracle(cfg, dom, fields)
    dt = cfg["dt"]
    for step in range(steps):
        p, e, t = pkg.synthetic.make_forcing(seed, step, dom["gid"], dt)
        ora.f["precipitation"][:], ora.f["potential_evaporation"][:], ora.f["temperature"][:] = p, e, t
        ora.update_model(dt)
    return cfg, dom, fields, ora


def main():
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    names = [n for n, _ in pkg._lib.field_table()] + ["number_of_layers", "n_unsatlayers"]
    for name in CASES:
        cfg, dom, fields, ora = run_case(pkg, name)
        out = {k: np.asarray(ora.f[k]) for k in names if k in ora.f}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, len(out), "fields,", cfg["n"], "cells")


if __name__ == "__main__":
    main()
