"""Summarise an .ncu-rep (raw + source pages) into the handful of numbers DESIGN.md / profiles cite."""
import csv, subprocess, sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, vals = r[0], r[-1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg"]
units = r[1] if len(r) > 2 else [""] * len(hdr)
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h} = {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]          # one block per captured kernel: take the first
h2 = rows[hi[0]]
data = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))] if len(r) == len(h2)]
isrc, isamp, iex = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
tot = sum(int(x[isamp] or 0) for x in data)
byop, exe = Counter(), Counter()
for x in data:
    t = x[isrc].split()
    op = (t[1] if t and t[0].startswith("@") else (t[0] if t else "?"))
    byop[op] += int(x[isamp] or 0); exe[op] += int(x[iex] or 0)
print("total warp-instructions executed", sum(exe.values()), "samples", tot)
for op, c in byop.most_common(14):
    print(f"  {op:32s} samples {100*c/max(tot,1):5.1f}%  executed {exe[op]}")
st = Counter()
for n in h2:
    if n.startswith("stall_") and "Not Issued" not in n:
        i = h2.index(n); st[n] = sum(int(x[i] or 0) for x in data)
print("stalls:", [(k, v) for k, v in st.most_common(8)])
