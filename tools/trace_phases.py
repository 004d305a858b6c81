#!/usr/bin/env python
"""Development aid: phase boundaries of one graph-replayed GAN step from a chrome trace (tools/step_timeline.py with
S2AG_TRACE=...): start/end of every recurrent kernel, the loss and Adam kernels, so the serial chain
pass #1 -> D step -> D(out) -> loss -> BPTT -> tail can be read off.  usage: python tools/trace_phases.py trace.json"""
import json
import sys

tr = json.load(open(sys.argv[1]))
evs = [e for e in tr['traceEvents'] if e.get('cat') == 'kernel']
t0 = min(e['ts'] for e in evs)
row = []
for e in sorted(evs, key=lambda e: e['ts']):
    n = e['name']
    tag = None
    if 'gru_persist_fwd' in n: tag = 'PF'
    elif 'gru_persist_bwd' in n: tag = 'PB'
    elif 'gru_cluster_fwd' in n: tag = 'cf'
    elif 'gru_cluster_bwd' in n: tag = 'CB40' if '<40>' in n else 'cb32'
    elif 'adam' in n: tag = 'ADAM'
    elif 'dis_loss' in n or 'gen_loss' in n: tag = 'LOSS'
    if tag:
        row.append('%s@%.2f-%.2f(s%s)' % (tag, (e['ts'] - t0) / 1000, (e['ts'] + e['dur'] - t0) / 1000, e['args'].get('stream')))
print(' '.join(row))
