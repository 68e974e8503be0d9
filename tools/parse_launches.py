"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel totals.
    python tools/parse_launches.py launches.csv [top_n] [--last-step]
--last-step: only the launches of the last complete training step (from the launch after the previous step's weight
re-pack to this step's re-pack), with the stream each kernel ran on."""
import collections
import csv
import sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
last_step = "--last-step" in sys.argv
rows = list(csv.reader(open(args[0])))
hdr = None
L = []
for r in rows:
    if 'Kernel Name' in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(d['Metric Value'].replace(',', ''))
    u = d['Metric Unit']
    v = v / 1e3 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1e3)
    L.append((d['Kernel Name'][:84], v, d.get('Stream', '?')))
if last_step:
    pk = [i for i, (n, _, _) in enumerate(L) if 'pack_weights_multi' in n]
    groups = []
    for i in pk:
        if groups and i - groups[-1][-1] <= 3:
            groups[-1].append(i)
        else:
            groups.append([i])
    assert len(groups) >= 2, "need two complete steps in the capture"
    L = L[groups[-2][-1] + 1:groups[-1][-1] + 1]
acc = collections.defaultdict(lambda: [0.0, 0, set()])
for n, v, s in L:
    acc[n][0] += v
    acc[n][1] += 1
    acc[n][2].add(s)
tot = sum(v[0] for v in acc.values())
top = int(args[1]) if len(args) > 1 else 30
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%10.1f us %5d  %5.1f%%  avg %8.1f  streams %-8s %s" % (v[0], v[1], 100 * v[0] / tot, v[0] / v[1], ",".join(sorted(v[2])), k))
print("total %.1f us over %d launches" % (tot, sum(v[1] for v in acc.values())))
per_stream = collections.defaultdict(float)
for n, v, s in L:
    per_stream[s] += v
print("per stream:", {k: round(v, 1) for k, v in per_stream.items()})
