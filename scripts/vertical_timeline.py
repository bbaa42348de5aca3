"""Developer aid: completion times of the kernels of the vertical update (B200)."""
import sys
sys.path.insert(0, '/root/repo')
from __graft_entry__ import load_pkg
pkg = load_pkg()
cfg, dom, fields = pkg.synthetic.make_basin(1000, 1000, seed=42)
opts = [a for a in sys.argv[1:] if not a.startswith("--")]
if "--cfg" in sys.argv:   # e.g. --cfg unsat_inline_iters=4
    k, v = sys.argv[sys.argv.index("--cfg") + 1].split("=")
    cfg[k] = int(v)
    opts = [a for a in opts if a != f"{k}={v}"]
m = pkg.SbmModel(cfg, dom, fields)
m.set_option("vertical_graph", 0)
m.set_option("vertical_timeline", 1)
for kv in opts:   # option=value
    k, v = kv.split("=")
    m.set_option(k, int(v))
dt = cfg["dt"]
for s in range(12):
    m.set_forcing(*pkg.synthetic.make_forcing(42, s, dom["gid"], dt))
    m.update_model(dt)
for rep in range(3):
    m.update_model(dt)
    m.synchronize()
    print("land_hydrology, unsat_engine, soil_column done at [us]:", [round(1000 * x) for x in m.vertical_timeline()[1:]])
print("suspended cells by trip count <=4, (4,8], (8,16], ... ; later loops:", m.unsat_buckets())
m.close()
