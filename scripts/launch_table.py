"""Print the last epoch's kernels of an ncu launch list (csv from --metrics gpu__time_duration.sum): name, microseconds."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]
ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
data = [(r[ik].split('(')[0][:70], float(r[iv].replace(',', ''))) for r in rows[hi + 1:] if len(r) > iv]
names = [k for k, _ in data]
last = max(i for i, k in enumerate(names) if 'k_adagrad' in k)
prev = max(i for i, k in enumerate(names[:last]) if 'k_adagrad' in k)
tot = 0.0
for k, v in data[prev + 1:last + 1]:
    print(f"{v / 1000:9.1f} us  {k}")
    tot += v
print(f"{tot / 1000:9.1f} us  sum over {last - prev} kernels (cold cache, serialised)")
