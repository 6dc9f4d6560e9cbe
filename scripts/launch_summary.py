"""Summarise an ncu launch list (gpu__time_duration.sum csv): per-kernel totals and the in-order list."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == 'ID':
        hdr, start = r, i + 1
        break
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[start:] if len(r) >= len(hdr)]
agg = collections.OrderedDict()
for r in data:
    k = r[idx['Kernel Name']][:70] + ' g=' + r[idx['Grid Size']] + ' b=' + r[idx['Block Size']]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[idx['Metric Value']])
tot = sum(a[1] for a in agg.values())
print(f"{len(data)} launches, {tot/1e3:.1f} us total (serialised, cold-cache)")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{a[0]:6d} {a[1]/1e3:10.1f} us  {a[1]/a[0]/1e3:8.2f} us avg {100*a[1]/tot:5.1f}%  {k}")
if len(sys.argv) > 2:
    lo, hi = int(sys.argv[2]), int(sys.argv[3])
    for r in data[lo:hi]:
        print(f"{float(r[idx['Metric Value']])/1e3:8.2f} us  {r[idx['Kernel Name']][:80]} g={r[idx['Grid Size']]}")
