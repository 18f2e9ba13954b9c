import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h = {n:i for i,n in enumerate(rows[hi])}
agg = collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r) <= h['Metric Value']: continue
    name = r[h['Kernel Name']].split('(')[0][-60:]
    v = float(r[h['Metric Value']].replace(',',''))
    unit = r[h['Metric Unit']]
    if unit == 'ns': v /= 1000
    elif unit == 'ms': v *= 1000
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += v; a[2] = max(a[2], v)
for k,(n,t,m) in agg.items():
    print(f"{k:62s} n={n:4d} total {t:9.1f} us avg {t/n:8.2f} max {m:8.1f}")
