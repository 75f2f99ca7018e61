"""Read an `ncu --page source --csv` dump: top SASS instructions by stall samples (offline analysis helper)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = []
for r in rows[hi + 1:]:
    if len(r) > isamp and r[0] != 'Address':
        try:
            int(r[isamp] or 0)
            data.append(r)
        except ValueError:
            pass
tot = sum(int(r[isamp] or 0) for r in data)
print('total samples', tot, 'instructions', len(data), 'warp-instr executed', sum(int(r[iex] or 0) for r in data))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:N]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{int(r[isamp]):6d} {100*int(r[isamp])/max(tot,1):5.1f}% ex={r[iex]:>8s} {r[0][-5:]} {r[isrc][:90]:90s} {st}")
