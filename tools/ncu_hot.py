"""Top stalled SASS instructions of one role region of a k_amp_tc capture (regions between USETMAXREG, like ncu_roles.py).
Usage: python tools/ncu_hot.py report.ncu-rep [region index 0..3 = prologue, activation, producers+mma, epilogue] [top N]"""
import csv, subprocess, sys
rep = sys.argv[1]; reg = int(sys.argv[2]) if len(sys.argv) > 2 else 1; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h2 = rows[hi[0]]
data = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))] if len(r) == len(h2)]
isrc, isamp, iex = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
stall_cols = [(i, n) for i, n in enumerate(h2) if n.startswith("stall_")]
marks = [i for i, x in enumerate(data) if 'USETMAXREG' in x[isrc]]
bounds = [0] + marks[:3] + [len(data)]
a, b = bounds[reg], bounds[reg + 1]
tot = sum(int(x[isamp] or 0) for x in data[a:b])
print(f"region {reg}: SASS lines {a}..{b}, {tot} stall samples; columns: line, samples, executed, instruction, top stall reasons")
order = sorted(range(a, b), key=lambda i: -int(data[i][isamp] or 0))[:topn]
for i in sorted(order):
    x = data[i]
    st = sorted(((int(x[j] or 0), n[6:]) for j, n in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {int(x[isamp] or 0):5d} {int(x[iex] or 0):9d}  {x[isrc][:70]:70s} " + " ".join(f"{n}:{c}" for c, n in st if c))
