#!/usr/bin/env python
"""Development aid: approximate critical path of one graph-replayed step from a chrome trace written by
tools/step_timeline.py (S2AG_TRACE=...).  Walking back from the last kernel, the predecessor of a kernel is the kernel
(on any stream) that ended last before it started -- the one it most plausibly waited for.  Prints the chain's time by
kernel name, the idle gaps, and the chain itself in 0.5 ms buckets.  usage: python tools/critical_path.py trace.json"""
import collections
import json
import sys

tr = json.load(open(sys.argv[1]))
evs = [(e['ts'], e['ts'] + e['dur'], e['name']) for e in tr['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy')]
evs.sort()
t0 = evs[0][0]
cur = max(evs, key=lambda e: e[1])
chain = [cur]
while True:
    s = cur[0]
    prev = [e for e in evs if e[1] <= s + 0.5 and e is not cur and e[0] < s]
    if not prev:
        break
    cur = max(prev, key=lambda e: e[1])
    chain.append(cur)
chain.reverse()
by = collections.Counter()
cnt = collections.Counter()
gap = 0.0
for a, b in zip(chain[:-1], chain[1:]):
    gap += max(0.0, b[0] - a[1])
for s, e, n in chain:
    key = n.split('<')[0].replace('void ', '')[:60]
    by[key] += e - s
    cnt[key] += 1
tot = chain[-1][1] - chain[0][0]
print("chain: %d kernels, span %.0f us, kernel time %.0f us, gaps %.0f us" % (len(chain), tot, sum(by.values()), gap))
for k, v in by.most_common(30):
    print("%8.0f us %5.1f%% %4d  %s" % (v, 100 * v / tot, cnt[k], k))
print()
bucket = collections.defaultdict(collections.Counter)
for s, e, n in chain:
    bucket[int((s - t0) // 500)][n.split('<')[0].replace('void ', '').split('::')[-1][:28]] += e - s
for b in sorted(bucket):
    print("%5.1f ms  %s" % (b * 0.5, ", ".join("%s %.0f" % kv for kv in bucket[b].most_common(4))))
