"""Attribute executed instructions / stall samples of k_amp_tc to its warp roles (regions between USETMAXREG)."""
import csv, subprocess, sys
from collections import Counter
rep=sys.argv[1]; tiles=float(sys.argv[2]) if len(sys.argv)>2 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h2 = rows[hi[0]]
data = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))] if len(r) == len(h2)]
isrc, isamp, iex = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
marks=[i for i,x in enumerate(data) if 'USETMAXREG' in x[isrc]]
regions={'prologue':(0,marks[0]),'activation':(marks[0],marks[1]),'producers+mma':(marks[1],marks[2]),'epilogue+wait loops':(marks[2],len(data))} if len(marks)>=3 else {'all':(0,len(data))}
allex=sum(int(x[iex] or 0) for x in data)
print(f"total warp-instructions {allex}  per tile {allex/tiles:.0f}")
for n,(a,b) in regions.items():
    ex=sum(int(x[iex] or 0) for x in data[a:b]); sm=sum(int(x[isamp] or 0) for x in data[a:b])
    print(f'{n:14s} static {b-a:5d}  executed/tile {ex/tiles:9.1f} ({100*ex/allex:4.1f}%)  stall samples {sm}')
    ops=Counter()
    for x in data[a:b]:
        t=x[isrc].split()
        if not t: continue
        op=(t[1] if t[0].startswith('@') else t[0]).split('.')[0]; ops[op]+=int(x[iex] or 0)
    print('      ', ' '.join(f'{o}:{c/tiles:.0f}' for o,c in ops.most_common(16)))
