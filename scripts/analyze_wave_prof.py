"""Developer aid: summarise gpurun_out/wave_prof.csv (WFB_WAVE_PROF=1 overland, 2 river, 3 subsurface)."""
import sys
import numpy as np
d = np.genfromtxt(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/wave_prof.csv', delimiter=',', names=True)
span = (d['end_ns'].max() - d['start_ns'].min()) / 1e6
print(len(d), 'chunks; kernel span ms', span)
dur = (d['end_ns'] - d['loaded_ns']) / 1e3
load = (d['loaded_ns'] - d['start_ns']) / 1e3
us_stage = dur / d['stages']
print('chunk load us: med %.2f p90 %.2f;  walk us: med %.1f p90 %.1f max %.1f' % (np.median(load), np.percentile(load, 90), np.median(dur), np.percentile(dur, 90), dur.max()))
print('us/stage: p10 %.2f med %.2f p90 %.2f p99 %.2f' % tuple(np.percentile(us_stage, [10, 50, 90, 99])))
noin = d['inlets'] == 0
print('chunks without inlets: %d, us/stage med %.2f p90 %.2f' % (noin.sum(), np.median(us_stage[noin]), np.percentile(us_stage[noin], 90)))
win = ~noin
if win.sum():
    print('chunks with inlets: %d, us/stage med %.2f p90 %.2f; inlet-wait+sync cycles/stage med %.0f p90 %.0f' % (
        win.sum(), np.median(us_stage[win]), np.percentile(us_stage[win], 90), np.median(d['wait_cyc'][win] / d['stages'][win]),
        np.percentile(d['wait_cyc'][win] / d['stages'][win], 90)))
print('mean concurrency (sum of chunk lifetimes / span): %.1f warps' % (((d['end_ns'] - d['start_ns']).sum() / 1e6) / span))
print('start times ms pct 50/90/99/max', np.percentile(d['start_ns'] / 1e6, [50, 90, 99, 100]))
idx = np.argsort(d['end_ns'])[-4:]
for i in idx:
    print(int(d['chunk'][i]), 'start %.3f end %.3f ms' % (d['start_ns'][i] / 1e6, d['end_ns'][i] / 1e6), 'stages', int(d['stages'][i]),
          'nodes', int(d['nodes'][i]), 'l0', int(d['l0'][i]), 'l1', int(d['l1'][i]), 'inlets', int(d['inlets'][i]), 'us/stage %.2f' % us_stage[i])
for L in range(0, int(d['l1'].max()) + 1, 100):
    m = (d['l1'] >= L) & (d['l1'] < L + 100)
    if m.sum():
        print('  outlet levels %4d..%4d: chunks %5d  start med %.3f ms  end med %.3f ms' % (L, L + 99, m.sum(), np.median(d['start_ns'][m]) / 1e6, np.median(d['end_ns'][m]) / 1e6))
