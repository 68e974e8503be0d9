import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; acc = collections.defaultdict(lambda: [0.0, 0])
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    v = float(d['Metric Value'].replace(',', '')); u = d['Metric Unit']
    v = v / 1e3 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1e3)
    name = d['Kernel Name'][:80]
    acc[name][0] += v; acc[name][1] += 1
tot = sum(v[0] for v in acc.values())
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%10.1f us %5d  %5.1f%%  avg %8.1f  %s" % (v[0], v[1], 100 * v[0] / tot, v[0] / v[1], k))
print("total %.1f us over %d launches" % (tot, sum(v[1] for v in acc.values())))
