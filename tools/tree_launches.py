import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
data = [(r[ki], float(r[vi].replace(',', '')), r[ui]) for r in rows[hdr + 1:] if len(r) > vi]
idx = [i for i, d in enumerate(data) if 'k_msm_digits' in d[0]]
seg = data[idx[-1]:]
agg = {}
tot = 0
for name, v, u in seg:
    ms = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v
    tot += ms
    k = name.split('(')[0].split('<')[0].replace('void ', '').replace('g16::', '')
    agg.setdefault(k, []).append(ms)
for k, v in agg.items():
    print("%-40s n=%2d total %7.3f ms  [%s]" % (k[:40], len(v), sum(v), " ".join("%.3f" % x for x in v[:8])))
print("total", tot)
