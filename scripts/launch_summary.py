"""Per-kernel totals of an ncu launch list (csv with gpu__time_duration.sum [+ dram bytes])."""
import csv
import sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
ki = H.index('Kernel Name'); vi = H.index('Metric Value'); mi = H.index('Metric Name')
tot = {}
for r in data:
    k = r[ki].split('(')[0].split('::')[-1][:40]; t = tot.setdefault(k, [0, 0.0, 0.0, 0.0])
    v = float(r[vi].replace(',', ''))
    if r[mi] == 'gpu__time_duration.sum': t[0] += 1; t[1] += v / 1000
    elif r[mi] == 'dram__bytes_read.sum': t[2] += v
    elif r[mi] == 'dram__bytes_write.sum': t[3] += v
all_us = sum(v[1] for v in tot.values())
print('total %.1f us' % all_us)
for k, v in sorted(tot.items(), key=lambda x: -x[1][1])[:top]:
    print('%-42s %5d %10.1f us %5.1f%%  rd %12.0f wr %12.0f' % (k, v[0], v[1], 100 * v[1] / all_us, v[2], v[3]))
