import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if hdr is None:
        if "Kernel Name" in r: hdr = r
        continue
    if len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(d["Metric Value"].replace(",", ""))
    if d.get("Metric Unit") == "ns": v /= 1000.0
    elif d.get("Metric Unit") == "ms": v *= 1000.0
    name = d["Kernel Name"].split("(")[0][:60]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{k:62s} n={v[0]:5d} avg={v[1]/v[0]:8.2f} us share={v[1]/tot:6.3f}")
