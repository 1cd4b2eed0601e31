#!/bin/bash
# same-box A/B of two builds of the library: index_tts_lora_b200/libbvg_prev.so (built from the previous commit) against
# libbvg.so, alternating, per-launch class sums + decode time.  Usage: bash tools/ab_libs.sh <tag>
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
TAG=${1:-ab}
for rep in 1 2; do
  for v in prev new; do
    if [ $v = prev ]; then export BVG_LIB=$PWD/index_tts_lora_b200/libbvg_prev.so; else unset BVG_LIB; fi
    BVG_PROF_DUMP=1 timeout -s KILL 200 python tools/per_launch.py 2> gpurun_out/${TAG}_pl_${v}${rep}.txt > /dev/null
    echo "$v$rep: $(timeout -s KILL 200 python tools/latency_ab.py 16 234 2>&1 | tail -1)"
  done
done
python tools/pl_sum.py gpurun_out/${TAG}_pl_*.txt
