import sys
sys.path.insert(0,'/root/repo')
from __graft_entry__ import load_pkg
pkg=load_pkg()
cfg,dom,fields=pkg.synthetic.make_basin(1000,1000,seed=42)
m=pkg.SbmModel(cfg,dom,fields)
dt=cfg['dt']
for s in range(14):
    m.set_forcing(*pkg.synthetic.make_forcing(42,s,dom['gid'],dt))
    m.update_model(dt)
    print('step',s,end=' ',flush=True); m.stats()
