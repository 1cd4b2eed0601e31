#!/bin/bash
# quick A/B of a kernel change: per-launch table, decode time, op + generator parity subset.  Usage: bash tools/gpu_call4.sh <tag>
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
TAG=${1:-r2x}
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tcgen05_vs_oracle or full_generator_bf16_snr or ragged_batch_bf16 or cluster" 2>&1 | tail -4
BVG_PROF_DUMP=1 timeout -s KILL 200 python tools/per_launch.py 2> gpurun_out/${TAG}_per_launch.txt | tail -3
python - <<PY
import re
us={0:0,1:0,2:0,3:0}
for l in open("gpurun_out/${TAG}_per_launch.txt"):
    m=re.match(r"bvg_prof (\d+) cls (\d) us ([\d.]+)",l)
    if m: us[int(m.group(2))]+=float(m.group(3))
print("serialised ms by class (wide, narrow, up, other):",{k:round(v/1000,2) for k,v in us.items()}, "sum", round(sum(us.values())/1000,2))
PY
timeout -s KILL 200 python tools/latency_ab.py 16 234 2>&1 | tail -1; timeout -s KILL 200 python tools/latency_ab.py 1 157 2>&1 | tail -1
