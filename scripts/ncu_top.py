"""Top stall-sample SASS lines of kernel block N in an `ncu --page source --csv` dump: ncu_top.py file.csv [block] [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = starts[blk]
e = starts[blk + 1] if blk + 1 < len(starts) else len(rows)
print(rows[s][1][:100])
hdr = rows[s + 1]
body = rows[s + 2:e]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[ismp] or 0) for r in body)
print("lines", len(body), "samples", tot)
order = sorted(range(len(body)), key=lambda i: -int(body[i][ismp] or 0))[:top]
for i in sorted(order):
    r = body[i]
    print(f"{i:5d} {int(r[ismp]):6d} {100*int(r[ismp])/max(tot,1):5.1f}% ex={r[iex]:>8s}  {r[isrc].strip()[:110]}")
