"""Developer aid: summarise gpurun_out/wave_prof.csv (WFB_WAVE_PROF=1 overland, 2 river, 3 subsurface)."""
import sys
import numpy as np
d = np.genfromtxt(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/wave_prof.csv', delimiter=',', names=True)
print(len(d), 'chunks; kernel span ms', (d['end_ns'].max() - d['start_ns'].min()) / 1e6)
dur = (d['end_ns'] - d['start_ns']) / 1e3
us_stage = dur / d['stages']
print('chunk duration us: med %.1f p90 %.1f max %.1f' % (np.median(dur), np.percentile(dur, 90), dur.max()))
print('us/stage: p10 %.2f med %.2f p90 %.2f' % tuple(np.percentile(us_stage, [10, 50, 90])))
noin = d['inlets'] == 0
print('chunks without inlets: %d, us/stage med %.2f p90 %.2f' % (noin.sum(), np.median(us_stage[noin]), np.percentile(us_stage[noin], 90)))
big = noin & (d['nodes'] > 128)
if big.sum():
    print('  of those with >128 nodes: %d, us/stage med %.2f; barrier-wait cycles/stage of thread 0 med %.0f' % (
        big.sum(), np.median(us_stage[big]), np.median(d['bar_cyc'][big] / d['stages'][big])))
win = ~noin
if win.sum():
    print('chunks with inlets: %d, us/stage med %.2f p90 %.2f; fetch busy cycles/stage med %.0f' % (
        win.sum(), np.median(us_stage[win]), np.percentile(us_stage[win], 90), np.median(d['fetch_cyc'][win] / d['stages'][win])))
print('sum of chunk durations / span = mean concurrency %.1f' % (dur.sum() / 1e3 / ((d['end_ns'].max() - d['start_ns'].min()) / 1e6)))
order = np.argsort(d['start_ns'])
print('start times ms pct 50/90/99/max', np.percentile(d['start_ns'] / 1e6, [50, 90, 99, 100]))
idx = np.argsort(d['end_ns'])[-5:]
for i in idx:
    print(int(d['chunk'][i]), 'start %.3f end %.3f ms' % (d['start_ns'][i] / 1e6, d['end_ns'][i] / 1e6), 'stages', int(d['stages'][i]),
          'nodes', int(d['nodes'][i]), 'l0', int(d['l0'][i]), 'l1', int(d['l1'][i]), 'inlets', int(d['inlets'][i]), 'us/stage %.2f' % us_stage[i])
# global pace: when does the chunk holding level L finish, per 100 levels
for L in range(0, int(d['l1'].max()) + 1, 100):
    m = (d['l1'] >= L) & (d['l1'] < L + 100)
    if m.sum():
        print('  outlet levels %4d..%4d: chunks %4d  start med %.3f ms  end med %.3f ms' % (L, L + 99, m.sum(), np.median(d['start_ns'][m]) / 1e6, np.median(d['end_ns'][m]) / 1e6))
