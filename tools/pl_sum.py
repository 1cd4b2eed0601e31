"""Sum a per-launch dump (BVG_PROF_DUMP lines) by class.  Usage: python tools/pl_sum.py file [file ...]"""
import re, sys
for f in sys.argv[1:]:
    us = {0: 0.0, 1: 0.0, 2: 0.0, 3: 0.0}
    for l in open(f):
        m = re.match(r"bvg_prof (\d+) cls (\d) us ([\d.]+)", l)
        if m: us[int(m.group(2))] += float(m.group(3))
    print(f, "serialised ms (wide, narrow, up, other):", {k: round(v / 1000, 2) for k, v in us.items()}, "sum", round(sum(us.values()) / 1000, 2))
