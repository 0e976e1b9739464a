"""Top stall-sample instructions from `ncu --page source --csv` (SASS view)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
si, ns, ie = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[hi + 1:] if len(r) > ns and r[ns].isdigit()]
tot = sum(int(r[ns]) for r in body) or 1
print(f"total samples {tot}, instructions {len(body)}")
top = sorted(body, key=lambda r: -int(r[ns]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    st = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{100 * int(r[ns]) / tot:5.1f}%  exec={r[ie]:>8}  {r[si][:90]:90s} {st}")
