import sys
import numpy as np
d = np.genfromtxt(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/wave_prof.csv', delimiter=',', names=True)
S = int(sys.argv[2]) if len(sys.argv) > 2 else 24
print(len(d), 'chunks; kernel span ms', (d['end_ns'].max() - d['start_ns'].min()) / 1e6)
dur = (d['end_ns'] - d['start_ns']) / 1e3
big = d['nodes'] > 300
print('nodes: min/med/max', d['nodes'].min(), np.median(d['nodes']), d['nodes'].max(), ' big chunks', big.sum())
print('proc cycles/stage (big) med %.0f  wait cycles/stage med %.0f' % (
    np.median(d['proc_cyc'][big] / d['stages'][big]), np.median(d['wait_cyc'][big] / d['stages'][big])))
print('work/stage (big) med %.0f node-substeps' % np.median(d['nodes'][big] * S / d['stages'][big]))
noin = big & (d['wait_cyc'] / d['stages'] < 3000)
print('chunks that barely wait:', noin.sum(), 'their us/stage med %.2f' % np.median(dur[noin] / d['stages'][noin]) if noin.sum() else '')
idx = np.argsort(d['end_ns'])[-5:]
for i in idx:
    print(int(d['chunk'][i]), 'start %.2f end %.2f ms' % (d['start_ns'][i] / 1e6, d['end_ns'][i] / 1e6), 'stages', int(d['stages'][i]),
          'nodes', int(d['nodes'][i]), 'l0', int(d['l0'][i]), 'l1', int(d['l1'][i]),
          'wait/stage %.0f proc/stage %.0f' % (d['wait_cyc'][i] / d['stages'][i], d['proc_cyc'][i] / d['stages'][i]))
print('start times ms pct 50/90/99/max', np.percentile(d['start_ns'] / 1e6, [50, 90, 99, 100]))
